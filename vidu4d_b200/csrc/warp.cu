// warp.cu -- fused "bob" skinning warp of the canonical surfels into every frame's camera space, forward and backward
// (the step right BEFORE the rasterizer in a Stage-3 iteration).
//
// Replaces (behaviour, not code) of the reference's PyTorch chain
//   lab4d/nnutils/deformable_gaussian.py:1395-1434   forward_warp
//   lab4d/nnutils/deformable_gaussian.py:1033-1046   apply_qt_to_gaussian           x' = q x q* + t,  r' = q (x) r
//   lab4d/nnutils/warping.py:378-444                 SkinningWarp.forward (forward direction, return_qt)
//   lab4d/nnutils/skinning.py:89-142                 Gaussian skinning logits        -(|x_bone / gauss|^2 + delta)
//   lab4d/utils/geom_utils.py:48-92                  dual_quaternion_skinning        sign-aligned blend, normalise
//   lab4d/utils/quat_transform.py                    quaternion algebra (real part first); the 3-vector products are what
//   lab4d/third_party/quaternion/src/quaternion.cu:27-60   the reference's own helper kernels do (row N3)
//   lab4d/utils/loss_utils.py:21-42                  cross_entropy_skin_loss
// which materialises two (M, P, B, 4) dual-quaternion tensors (~240 MB each at M = 2, P = 300 K, B = 25), a (M, P, B)
// sign tensor and several (M, P, B, 3) temporaries per step, and recomputes the FRAME-INDEPENDENT skinning weights for
// every frame.  Here one thread owns one surfel: its B skinning weights live in registers, are computed once, and are
// reused for all M frames; the per-bone / per-frame transforms sit in shared memory.  HBM traffic is the algorithmic
// minimum: read 28 B (+ 4B of delta) per surfel, write 28 B per surfel and frame.
//
// Inputs that the kernel treats as small tables are prepared by the caller with ordinary (differentiable) torch ops:
//   o2b_q (B,4), o2b_t (B,3)   object -> bone rigid transform of the REST pose (inverse of rest_articulation)
//   inv_gauss (B,3)            exp(-log_gauss)
//   se3_r, se3_d (M,B,4)       per frame and bone: t_articulation o rest_articulation^-1 as a dual quaternion
//   cam_q (M,4), cam_t (M,3)   field2cam of each frame
// The backward recomputes the weights (cheaper than storing (P,B) floats) and returns gradients for every input; table
// gradients are reduced warp -> block (shared memory) -> global (one atomic per value and block).
#include "common.cuh"

namespace {

struct Q4 { float w, x, y, z; };
__device__ __forceinline__ Q4 qmul(const Q4& a, const Q4& b) {
    return {a.w * b.w - a.x * b.x - a.y * b.y - a.z * b.z, a.w * b.x + a.x * b.w + a.y * b.z - a.z * b.y,
            a.w * b.y - a.x * b.z + a.y * b.w + a.z * b.x, a.w * b.z + a.x * b.y - a.y * b.x + a.z * b.w};
}
__device__ __forceinline__ Q4 qconj(const Q4& q) { return {q.w, -q.x, -q.y, -q.z}; }
__device__ __forceinline__ Q4 pure(float x, float y, float z) { return {0.f, x, y, z}; }
__device__ __forceinline__ Q4 qadd(const Q4& a, const Q4& b) { return {a.w + b.w, a.x + b.x, a.y + b.y, a.z + b.z}; }
__device__ __forceinline__ Q4 qscale(const Q4& a, float s) { return {a.w * s, a.x * s, a.y * s, a.z * s}; }
__device__ __forceinline__ float qdot(const Q4& a, const Q4& b) { return a.w * b.w + a.x * b.x + a.y * b.y + a.z * b.z; }
// v' = q v q*  (general q, as quat_transform.py:259-276)
__device__ __forceinline__ void qrot(const Q4& q, float vx, float vy, float vz, float& ox, float& oy, float& oz) {
    const Q4 t = qmul(qmul(q, pure(vx, vy, vz)), qconj(q));
    ox = t.x; oy = t.y; oz = t.z;
}
// vjp of v' = q v q*:  g_v = (q* g q)[1:],  g_q = -2 g (x) q (x) v
__device__ __forceinline__ void qrot_vjp(const Q4& q, float vx, float vy, float vz, float gx, float gy, float gz, Q4& gq,
                                         float& gvx, float& gvy, float& gvz) {
    const Q4 g = pure(gx, gy, gz);
    const Q4 t = qmul(qmul(qconj(q), g), q);
    gvx = t.x; gvy = t.y; gvz = t.z;
    gq = qscale(qmul(qmul(g, q), pure(vx, vy, vz)), -2.0f);
}
__device__ __forceinline__ Q4 ldq(const float* p) { return {p[0], p[1], p[2], p[3]}; }

struct WarpDims { int P, B, M; };

// shared-memory tables: o2b_q[B*4] o2b_t[B*3] ig[B*3] se3_r[M*B*4] se3_d[M*B*4] cam_q[M*4] cam_t[M*3]
struct Tables {
    const float *oq, *ot, *ig, *sr, *sd, *cq, *ct;
    __device__ __forceinline__ static int floats(int B, int M) { return B * 10 + M * B * 8 + M * 7; }
    __device__ __forceinline__ void bind(float* s, int B, int M) {
        oq = s; ot = oq + B * 4; ig = ot + B * 3; sr = ig + B * 3; sd = sr + M * B * 4; cq = sd + M * B * 4; ct = cq + M * 4;
    }
};

__device__ __forceinline__ void load_tables(float* s, const WarpDims d, const float* __restrict__ o2b_q, const float* __restrict__ o2b_t,
                                            const float* __restrict__ inv_gauss, const float* __restrict__ se3_r,
                                            const float* __restrict__ se3_d, const float* __restrict__ cam_q,
                                            const float* __restrict__ cam_t) {
    const int B = d.B, M = d.M;
    float* p = s;
    for (int i = threadIdx.x; i < B * 4; i += blockDim.x) p[i] = __ldg(o2b_q + i);
    p += B * 4;
    for (int i = threadIdx.x; i < B * 3; i += blockDim.x) p[i] = __ldg(o2b_t + i);
    p += B * 3;
    for (int i = threadIdx.x; i < B * 3; i += blockDim.x) p[i] = __ldg(inv_gauss + i);
    p += B * 3;
    for (int i = threadIdx.x; i < M * B * 4; i += blockDim.x) p[i] = __ldg(se3_r + i);
    p += M * B * 4;
    for (int i = threadIdx.x; i < M * B * 4; i += blockDim.x) p[i] = __ldg(se3_d + i);
    p += M * B * 4;
    for (int i = threadIdx.x; i < M * 4; i += blockDim.x) p[i] = __ldg(cam_q + i);
    p += M * 4;
    for (int i = threadIdx.x; i < M * 3; i += blockDim.x) p[i] = __ldg(cam_t + i);
}

// Skinning weights of one surfel: softmax over bones of -(|x_bone * inv_gauss|^2 + delta); also the arg-max bone and
// log-sum-exp minus max (= the cross entropy against the arg-max one-hot, loss_utils.py:21-42).
template <int BMAX>
__device__ __forceinline__ void skin_weights(const Tables& T, int B, float x, float y, float z, const float* __restrict__ delta_row,
                                             float (&w)[BMAX], int& anchor, float& entropy) {
    float mx = -3.4e38f;
#pragma unroll
    for (int b = 0; b < BMAX; b++) {
        w[b] = -3.4e38f;
        if (b < B) {
            float bx, by, bz;
            qrot(ldq(T.oq + 4 * b), x, y, z, bx, by, bz);
            bx = (bx + T.ot[3 * b]) * T.ig[3 * b]; by = (by + T.ot[3 * b + 1]) * T.ig[3 * b + 1]; bz = (bz + T.ot[3 * b + 2]) * T.ig[3 * b + 2];
            float l = -(bx * bx + by * by + bz * bz);
            if (delta_row) l -= __ldg(delta_row + b);
            w[b] = l;
            mx = fmaxf(mx, l);
        }
    }
    float Z = 0.f;
    anchor = 0;
    bool found = false;
#pragma unroll
    for (int b = 0; b < BMAX; b++) {
        if (b < B) {
            if (!found && w[b] == mx) { anchor = b; found = true; }      // first arg-max, like torch.argmax
            w[b] = expf(w[b] - mx);
            Z += w[b];
        } else w[b] = 0.f;
    }
    const float inv = 1.0f / Z;
#pragma unroll
    for (int b = 0; b < BMAX; b++) w[b] *= inv;
    entropy = logf(Z);
}

// sign-aligned dual-quaternion blend of frame m (geom_utils.py:66-83): returns the un-normalised QR, QD
template <int BMAX>
__device__ __forceinline__ void blend(const Tables& T, int B, int m, const float (&w)[BMAX], int anchor, Q4& QR, Q4& QD) {
    const float* sr = T.sr + (size_t)m * B * 4;
    const float* sd = T.sd + (size_t)m * B * 4;
    const Q4 qa = ldq(sr + 4 * anchor);
    QR = {0.f, 0.f, 0.f, 0.f}; QD = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int b = 0; b < BMAX; b++) {
        if (b < B) {
            const Q4 r = ldq(sr + 4 * b), dq = ldq(sd + 4 * b);
            const float s = (qdot(qa, r) > 0.f ? 1.f : -1.f) * w[b];
            QR = qadd(QR, qscale(r, s)); QD = qadd(QD, qscale(dq, s));
        }
    }
}

template <int BMAX>
__global__ void __launch_bounds__(256)
bob_warp_fwd_kernel(const WarpDims d, const float* __restrict__ xyz, const float* __restrict__ rot,
                    const float* __restrict__ o2b_q, const float* __restrict__ o2b_t, const float* __restrict__ inv_gauss,
                    const float* __restrict__ delta, const float* __restrict__ se3_r, const float* __restrict__ se3_d,
                    const float* __restrict__ cam_q, const float* __restrict__ cam_t, float* __restrict__ xyz_cam,
                    float* __restrict__ rot_cam, float* __restrict__ skin_entropy) {
    extern __shared__ float smem[];
    load_tables(smem, d, o2b_q, o2b_t, inv_gauss, se3_r, se3_d, cam_q, cam_t);
    __syncthreads();
    Tables T; T.bind(smem, d.B, d.M);
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= d.P) return;
    const float x = xyz[3 * (size_t)p], y = xyz[3 * (size_t)p + 1], z = xyz[3 * (size_t)p + 2];
    const Q4 r = ldq(rot + 4 * (size_t)p);
    float w[BMAX];
    int anchor; float ent;
    skin_weights<BMAX>(T, d.B, x, y, z, delta ? delta + (size_t)p * d.B : nullptr, w, anchor, ent);
    if (skin_entropy) skin_entropy[p] = ent;
    for (int m = 0; m < d.M; m++) {
        Q4 QR, QD;
        blend<BMAX>(T, d.B, m, w, anchor, QR, QD);
        const float inv = rsqrtf(qdot(QR, QR));
        const Q4 q = qscale(QR, inv), qd = qscale(QD, inv);
        const Q4 t4 = qmul(qd, qconj(q));                           // t = 2 (qd (x) q*)[1:]
        float tx, ty, tz;
        qrot(q, x, y, z, tx, ty, tz);
        tx += 2.f * t4.x; ty += 2.f * t4.y; tz += 2.f * t4.z;
        const Q4 rt = qmul(q, r);
        const Q4 qc = ldq(T.cq + 4 * m);
        float cx, cy, cz;
        qrot(qc, tx, ty, tz, cx, cy, cz);
        float* ox = xyz_cam + ((size_t)m * d.P + p) * 3;
        ox[0] = cx + T.ct[3 * m]; ox[1] = cy + T.ct[3 * m + 1]; ox[2] = cz + T.ct[3 * m + 2];
        const Q4 rc = qmul(qc, rt);
        reinterpret_cast<float4*>(rot_cam)[(size_t)m * d.P + p] = make_float4(rc.w, rc.x, rc.y, rc.z);
    }
}

// Recursive-halving warp reduction of 32 values: lane L ends with the sum over the warp of v[L].  31 shuffles for 32
// sums (a plain shuffle tree needs 5 per value) and the 32 results land one per lane, so the shared-memory accumulation
// that follows is ONE conflict-free atomic instruction instead of 32 single-lane ones.
__device__ __forceinline__ float butterfly32(float (&v)[32], int lane) {
#pragma unroll
    for (int off = 16, n = 16; off >= 1; off >>= 1, n >>= 1) {
        const bool hi = lane & off;
#pragma unroll
        for (int i = 0; i < 16; i++) {
            if (i < n) {
                const float keep = hi ? v[i + n] : v[i], send = hi ? v[i] : v[i + n];
                v[i] = keep + __shfl_xor_sync(0xffffffffu, send, off);
            }
        }
    }
    return v[0];
}

// warp sum, then one shared-memory add per warp (only the 8 warps of a block contend on an address)
__device__ __forceinline__ void acc_table(float* slot, float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if ((threadIdx.x & 31) == 0 && v != 0.f) atomicAdd(slot, v);
}
__device__ __forceinline__ void acc_q(float* slot, const Q4& q) {
    acc_table(slot, q.w); acc_table(slot + 1, q.x); acc_table(slot + 2, q.y); acc_table(slot + 3, q.z);
}

template <int BMAX>
__global__ void __launch_bounds__(256)
bob_warp_bwd_kernel(const WarpDims d, const float* __restrict__ xyz, const float* __restrict__ rot,
                    const float* __restrict__ o2b_q, const float* __restrict__ o2b_t, const float* __restrict__ inv_gauss,
                    const float* __restrict__ delta, const float* __restrict__ se3_r, const float* __restrict__ se3_d,
                    const float* __restrict__ cam_q, const float* __restrict__ cam_t, const float* __restrict__ g_xyz_cam,
                    const float* __restrict__ g_rot_cam, const float* __restrict__ g_entropy, float* __restrict__ g_xyz,
                    float* __restrict__ g_rot, float* __restrict__ g_delta, float* __restrict__ g_tables) {
    extern __shared__ float smem[];
    const int B = d.B, M = d.M, nt = Tables::floats(B, M);
    load_tables(smem, d, o2b_q, o2b_t, inv_gauss, se3_r, se3_d, cam_q, cam_t);
    float* gacc = smem + nt;                                        // gradient of every table entry, same layout
    for (int i = threadIdx.x; i < nt; i += blockDim.x) gacc[i] = 0.f;
    __syncthreads();
    Tables T; T.bind(smem, B, M);
    float* g_oq = gacc; float* g_ot = g_oq + B * 4; float* g_ig = g_ot + B * 3; float* g_sr = g_ig + B * 3;
    float* g_sd = g_sr + M * B * 4; float* g_cq = g_sd + M * B * 4; float* g_ct = g_cq + M * 4;
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    const bool live = p < d.P;                                       // dead lanes still take part in the warp sums (with zeros)
    const size_t ps = live ? (size_t)p : 0;
    const float x = xyz[3 * ps], y = xyz[3 * ps + 1], z = xyz[3 * ps + 2];
    const Q4 r = ldq(rot + 4 * ps);
    float w[BMAX], gw[BMAX];
    int anchor; float ent;
    skin_weights<BMAX>(T, B, x, y, z, delta ? delta + ps * B : nullptr, w, anchor, ent);
#pragma unroll
    for (int b = 0; b < BMAX; b++) gw[b] = 0.f;
    float gxx = 0.f, gxy = 0.f, gxz = 0.f;
    Q4 grot = {0.f, 0.f, 0.f, 0.f};
    const float lv = live ? 1.f : 0.f;
    for (int m = 0; m < M; m++) {
        Q4 QR, QD;
        blend<BMAX>(T, B, m, w, anchor, QR, QD);
        const float inv = rsqrtf(qdot(QR, QR));
        const Q4 q = qscale(QR, inv), qd = qscale(QD, inv);
        const Q4 t4 = qmul(qd, qconj(q));
        float tx, ty, tz;
        qrot(q, x, y, z, tx, ty, tz);
        tx += 2.f * t4.x; ty += 2.f * t4.y; tz += 2.f * t4.z;
        const Q4 rt = qmul(q, r);
        const Q4 qc = ldq(T.cq + 4 * m);
        const float* gx = g_xyz_cam + ((size_t)m * d.P + ps) * 3;
        const float gcx = gx[0] * lv, gcy = gx[1] * lv, gcz = gx[2] * lv;
        const float4 grc4 = reinterpret_cast<const float4*>(g_rot_cam)[(size_t)m * d.P + ps];
        const Q4 grc = {grc4.x * lv, grc4.y * lv, grc4.z * lv, grc4.w * lv};
        // ---- camera: x_c = qc x_t qc* + tc ; r_c = qc (x) r_t
        Q4 gqc; float gtx, gty, gtz;
        qrot_vjp(qc, tx, ty, tz, gcx, gcy, gcz, gqc, gtx, gty, gtz);
        gqc = qadd(gqc, qmul(grc, qconj(rt)));
        const Q4 grt = qmul(qconj(qc), grc);
        acc_q(g_cq + 4 * m, gqc);
        acc_table(g_ct + 3 * m, gcx); acc_table(g_ct + 3 * m + 1, gcy); acc_table(g_ct + 3 * m + 2, gcz);
        // ---- x_t = q x q* + t ; r_t = q (x) rot
        Q4 gq; float ax, ay, az;
        qrot_vjp(q, x, y, z, gtx, gty, gtz, gq, ax, ay, az);
        gxx += ax; gxy += ay; gxz += az;
        gq = qadd(gq, qmul(grt, qconj(r)));
        grot = qadd(grot, qmul(qconj(q), grt));
        // ---- t = 2 (qd (x) q*)[1:]
        const Q4 gc = pure(2.f * gtx, 2.f * gty, 2.f * gtz);
        const Q4 gqd = qmul(gc, q);
        gq = qadd(gq, qconj(qmul(qconj(qd), gc)));
        // ---- normalisation q = QR / |QR|, qd = QD / |QR|
        const Q4 gQD = qscale(gqd, inv);
        const Q4 gQR = qscale(qadd(gq, qscale(q, -(qdot(q, gq) + qdot(qd, gqd)))), inv);
        // ---- blend
        const float* sr = T.sr + (size_t)m * B * 4;
        const float* sd = T.sd + (size_t)m * B * 4;
        const Q4 qa = ldq(sr + 4 * anchor);
        const int lane = threadIdx.x & 31;
#pragma unroll
        for (int c4 = 0; c4 < BMAX / 4; c4++) {
            if (4 * c4 < B) {                                        // four bones x (4 + 4) components = 32 sums per butterfly
                float vals[32];
#pragma unroll
                for (int j = 0; j < 4; j++) {
                    const int b = 4 * c4 + j;
                    float ws = 0.f;
                    if (b < B) {
                        const Q4 rb = ldq(sr + 4 * b), db = ldq(sd + 4 * b);
                        const float s = qdot(qa, rb) > 0.f ? 1.f : -1.f;
                        gw[b] += s * (qdot(gQR, rb) + qdot(gQD, db));
                        ws = w[b] * s;
                    }
                    vals[8 * j] = gQR.w * ws; vals[8 * j + 1] = gQR.x * ws; vals[8 * j + 2] = gQR.y * ws; vals[8 * j + 3] = gQR.z * ws;
                    vals[8 * j + 4] = gQD.w * ws; vals[8 * j + 5] = gQD.x * ws; vals[8 * j + 6] = gQD.y * ws; vals[8 * j + 7] = gQD.z * ws;
                }
                const float tot = butterfly32(vals, lane);
                const int b = 4 * c4 + (lane >> 3), comp = lane & 7;
                if (b < B && tot != 0.f)
                    atomicAdd((comp < 4 ? g_sr : g_sd) + ((size_t)m * B + b) * 4 + (comp & 3), tot);
            }
        }
    }
    // ---- softmax (+ entropy) backward, then the Gaussian logits
    float dotwg = 0.f;
#pragma unroll
    for (int b = 0; b < BMAX; b++) dotwg += w[b] * gw[b];
    const float ge = (g_entropy && live) ? g_entropy[ps] : 0.f;
    const int lane = threadIdx.x & 31;
#pragma unroll
    for (int c3 = 0; c3 < (BMAX + 2) / 3; c3++) {
        if (3 * c3 < B) {                                            // three bones x (4 + 3 + 3) components per butterfly
            float vals[32];
            vals[30] = 0.f; vals[31] = 0.f;
#pragma unroll
            for (int j = 0; j < 3; j++) {
                const int b = 3 * c3 + j;
#pragma unroll
                for (int i = 0; i < 10; i++) vals[10 * j + i] = 0.f;
                if (b < B && b < BMAX) {
                    const float gl = (w[b] * (gw[b] - dotwg) + ge * (w[b] - (b == anchor ? 1.f : 0.f))) * lv;
                    if (g_delta && live) g_delta[ps * B + b] = -gl;
                    const Q4 qb = ldq(T.oq + 4 * b);
                    float bx, by, bz;
                    qrot(qb, x, y, z, bx, by, bz);
                    bx += T.ot[3 * b]; by += T.ot[3 * b + 1]; bz += T.ot[3 * b + 2];
                    const float i0 = T.ig[3 * b], i1 = T.ig[3 * b + 1], i2 = T.ig[3 * b + 2];
                    // logit = -(|xb * ig|^2 + delta):  g_s = -2 s gl
                    const float gs0 = -2.f * bx * i0 * gl, gs1 = -2.f * by * i1 * gl, gs2 = -2.f * bz * i2 * gl;
                    const float gb0 = gs0 * i0, gb1 = gs1 * i1, gb2 = gs2 * i2;
                    Q4 gqb; float ax, ay, az;
                    qrot_vjp(qb, x, y, z, gb0, gb1, gb2, gqb, ax, ay, az);
                    gxx += ax; gxy += ay; gxz += az;
                    vals[10 * j] = gqb.w; vals[10 * j + 1] = gqb.x; vals[10 * j + 2] = gqb.y; vals[10 * j + 3] = gqb.z;
                    vals[10 * j + 4] = gb0; vals[10 * j + 5] = gb1; vals[10 * j + 6] = gb2;
                    vals[10 * j + 7] = gs0 * bx; vals[10 * j + 8] = gs1 * by; vals[10 * j + 9] = gs2 * bz;
                }
            }
            const float tot = butterfly32(vals, lane);
            const int j = lane / 10, i = lane - 10 * j, b = 3 * c3 + j;
            if (lane < 30 && b < B && tot != 0.f) {
                float* dst = i < 4 ? g_oq + 4 * b + i : (i < 7 ? g_ot + 3 * b + (i - 4) : g_ig + 3 * b + (i - 7));
                atomicAdd(dst, tot);
            }
        }
    }
    if (live) {
        g_xyz[3 * ps] = gxx; g_xyz[3 * ps + 1] = gxy; g_xyz[3 * ps + 2] = gxz;
        reinterpret_cast<float4*>(g_rot)[ps] = make_float4(grot.w, grot.x, grot.y, grot.z);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < nt; i += blockDim.x)
        if (gacc[i] != 0.f) atomicAdd(g_tables + i, gacc[i]);
}

}  // namespace

extern "C" {

SR_API size_t sr_bob_warp_table_floats(int32_t B, int32_t M) { return (size_t)B * 10 + (size_t)M * B * 8 + (size_t)M * 7; }

SR_API int sr_bob_warp_forward(int32_t P, int32_t B, int32_t M, const float* xyz, const float* rot, const float* o2b_q,
                               const float* o2b_t, const float* inv_gauss, const float* delta, const float* se3_r,
                               const float* se3_d, const float* cam_q, const float* cam_t, float* xyz_cam, float* rot_cam,
                               float* skin_entropy, void* stream_) {
    if (P < 0 || B < 1 || B > 64 || M < 1) return SR_EINVAL;
    if (P == 0) return 0;
    if (!xyz || !rot || !o2b_q || !o2b_t || !inv_gauss || !se3_r || !se3_d || !cam_q || !cam_t || !xyz_cam || !rot_cam) return SR_EINVAL;
    if (((uintptr_t)rot_cam & 15)) return SR_EINVAL;
    const WarpDims d{P, B, M};
    const size_t smem = sr_bob_warp_table_floats(B, M) * sizeof(float);
    if (smem > 200 * 1024) return SR_EINVAL;          // split the frames over several calls
    cudaStream_t s = (cudaStream_t)stream_;
    ProfileScope ps("bob_warp_fwd", s);
    auto launch = [&](auto kern) -> int {
        if (smem > 48 * 1024 && cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) return SR_ECUDA;
        kern<<<(P + 255) / 256, 256, smem, s>>>(d, xyz, rot, o2b_q, o2b_t, inv_gauss, delta, se3_r, se3_d, cam_q, cam_t, xyz_cam,
                                                 rot_cam, skin_entropy);
        return 0;
    };
    const int rc = B <= 32 ? launch(bob_warp_fwd_kernel<32>) : launch(bob_warp_fwd_kernel<64>);
    if (rc) return rc;
    sr_count_launch();
    return cudaGetLastError() == cudaSuccess ? 0 : SR_ECUDA;
}

/* g_tables: float[sr_bob_warp_table_floats(B, M)], laid out o2b_q[B*4] o2b_t[B*3] inv_gauss[B*3] se3_r[M*B*4] se3_d[M*B*4]
 * cam_q[M*4] cam_t[M*3]; zeroed here. */
SR_API int sr_bob_warp_backward(int32_t P, int32_t B, int32_t M, const float* xyz, const float* rot, const float* o2b_q,
                                const float* o2b_t, const float* inv_gauss, const float* delta, const float* se3_r,
                                const float* se3_d, const float* cam_q, const float* cam_t, const float* g_xyz_cam,
                                const float* g_rot_cam, const float* g_entropy, float* g_xyz, float* g_rot, float* g_delta,
                                float* g_tables, void* stream_) {
    if (P < 0 || B < 1 || B > 64 || M < 1) return SR_EINVAL;
    if (!g_tables) return SR_EINVAL;
    cudaStream_t s = (cudaStream_t)stream_;
    const size_t nt = sr_bob_warp_table_floats(B, M);
    if (cudaMemsetAsync(g_tables, 0, nt * sizeof(float), s) != cudaSuccess) return SR_ECUDA;
    if (P == 0) return 0;
    if (!xyz || !rot || !o2b_q || !o2b_t || !inv_gauss || !se3_r || !se3_d || !cam_q || !cam_t || !g_xyz_cam || !g_rot_cam ||
        !g_xyz || !g_rot)
        return SR_EINVAL;
    if (((uintptr_t)g_rot_cam & 15) || ((uintptr_t)g_rot & 15)) return SR_EINVAL;
    const WarpDims d{P, B, M};
    const size_t smem = 2 * nt * sizeof(float);
    if (smem > 200 * 1024) return SR_EINVAL;
    ProfileScope ps("bob_warp_bwd", s);
    auto launch = [&](auto kern) -> int {
        if (smem > 48 * 1024 && cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) return SR_ECUDA;
        // 128-thread blocks: the kernel holds ~200 registers per thread (B weights + B weight gradients)
        kern<<<(P + 127) / 128, 128, smem, s>>>(d, xyz, rot, o2b_q, o2b_t, inv_gauss, delta, se3_r, se3_d, cam_q, cam_t, g_xyz_cam,
                                                 g_rot_cam, g_entropy, g_xyz, g_rot, g_delta, g_tables);
        return 0;
    };
    const int rc = B <= 32 ? launch(bob_warp_bwd_kernel<32>) : launch(bob_warp_bwd_kernel<64>);
    if (rc) return rc;
    sr_count_launch();
    return cudaGetLastError() == cudaSuccess ? 0 : SR_ECUDA;
}

}  // extern "C"
