// loss.cu -- the Stage-3 image losses fused with the render() post-processing, forward AND backward in one pass
// (second half).
//
// Replaces, per frame, ~35 + ~40 small PyTorch kernels forward and ~100 backward:
//   gs/gaussian_renderer/__init__.py:121-145   alpha / rotated normal / nan_to_num'd expected & median depth / surf_depth
//   gs/utils/point_utils.py:9-37               depth_to_normal (central differences, cross, normalise) * alpha.detach()
//   lab4d/nnutils/deformable_gaussian.py:188-190  render += (1 - acc) * learnable_bkgd           (optional)
//   lab4d/engine/model.py:674-692              masked L1 on rgb where vis2d > 0, mean over all elements
//   lab4d/engine/model.py:649-653              (acc - mask)^2 * mask_balance_wt, mean
//   lab4d/engine/model.py:817-842              lambda_n * mean(1 - <rend_normal, surf_normal>), lambda_d * mean(rend_dist)
//
// The losses are scalars whose gradient with respect to the rasterizer's outputs is analytic, so ONE kernel computes the
// four loss terms (block reduction + 4 atomics per block) and writes dL/dcolor and dL/dallmap directly; the rasterizer
// backward consumes them (times the upstream scalar, sr_backward_batch's grad_scale).  The surf_normal stencil is
// handled as a gather: a pixel re-derives its four neighbours' stencils, so there are no atomics on image data and the
// result is deterministic.  Batched: blockIdx.z = frame.
#include "post_common.cuh"

namespace {
using namespace post;

struct LossCfg {
    int W, H;
    float tanx, tany, depth_ratio;
    float w_rgb, w_mask, lambda_normal, lambda_dist;
};

__global__ void __launch_bounds__(256)
loss_depth_kernel(const LossCfg g, const float* __restrict__ allmap, float* __restrict__ surf_depth) {
    const int x = blockIdx.x * 32 + (threadIdx.x & 31), y = blockIdx.y * 8 + (threadIdx.x >> 5), f = blockIdx.z;
    if (x >= g.W || y >= g.H) return;
    const size_t N = (size_t)g.W * g.H, p = (size_t)y * g.W + x;
    const float* am = allmap + (size_t)f * 8 * N;
    const float med = nan_to_num00(am[5 * N + p]);
    const float ex = nan_to_num00(am[p] / am[N + p]);
    surf_depth[(size_t)f * N + p] = ex * (1.f - g.depth_ratio) + g.depth_ratio * med;
}

// un-normalised stencil normal at (x, y) and the two difference vectors; false outside the interior
__device__ __forceinline__ bool stencil(const PostCam& c, int W, int H, int x, int y, const float* __restrict__ sd,
                                        float* dx, float* dy, float* n) {
    if (!(x >= 1 && x < W - 1 && y >= 1 && y < H - 1)) return false;
    const size_t p = (size_t)y * W + x;
    float ax, ay, az, bx, by, bz, ex, ey, ez, fx, fy, fz;
    ray_dir(c, x, y + 1, ax, ay, az); ray_dir(c, x, y - 1, bx, by, bz);
    ray_dir(c, x + 1, y, ex, ey, ez); ray_dir(c, x - 1, y, fx, fy, fz);
    const float dD = sd[p + W], dU = sd[p - W], dR = sd[p + 1], dL = sd[p - 1];
    dx[0] = dD * ax - dU * bx; dx[1] = dD * ay - dU * by; dx[2] = dD * az - dU * bz;
    dy[0] = dR * ex - dL * fx; dy[1] = dR * ey - dL * fy; dy[2] = dR * ez - dL * fz;
    n[0] = dx[1] * dy[2] - dx[2] * dy[1]; n[1] = dx[2] * dy[0] - dx[0] * dy[2]; n[2] = dx[0] * dy[1] - dx[1] * dy[0];
    return true;
}

// gradient of lambda_n * mean(1 - <rend_normal, surf_normal>) through the stencil centred at (x, y), with respect to its
// two difference vectors (alpha detached): g_dx, g_dy
__device__ __forceinline__ bool stencil_loss_vjp(const PostCam& c, const LossCfg& g, int x, int y, const float* __restrict__ sd,
                                                 const float* __restrict__ am, size_t N, float scale, float* gdx, float* gdy) {
    float dx[3], dy[3], n[3];
    if (!stencil(c, g.W, g.H, x, y, sd, dx, dy, n)) return false;
    const size_t p = (size_t)y * g.W + x;
    const float a = am[N + p];
    const float n0 = am[2 * N + p], n1 = am[3 * N + p], n2 = am[4 * N + p];
    // g_sn = -lambda_n / N * rend_normal ; times alpha (the detached factor)
    const float s = scale * a;
    const float g0 = s * (c.R[0] * n0 + c.R[1] * n1 + c.R[2] * n2), g1 = s * (c.R[3] * n0 + c.R[4] * n1 + c.R[5] * n2),
                g2 = s * (c.R[6] * n0 + c.R[7] * n1 + c.R[8] * n2);
    const float len = sqrtf(n[0] * n[0] + n[1] * n[1] + n[2] * n[2]);
    float gn0, gn1, gn2;
    if (len > 1e-12f) {
        const float inv = 1.0f / len;
        const float u0 = n[0] * inv, u1 = n[1] * inv, u2 = n[2] * inv;
        const float d = u0 * g0 + u1 * g1 + u2 * g2;
        gn0 = (g0 - u0 * d) * inv; gn1 = (g1 - u1 * d) * inv; gn2 = (g2 - u2 * d) * inv;
    } else {                                             // v / eps branch of F.normalize
        gn0 = g0 * 1e12f; gn1 = g1 * 1e12f; gn2 = g2 * 1e12f;
    }
    gdx[0] = dy[1] * gn2 - dy[2] * gn1; gdx[1] = dy[2] * gn0 - dy[0] * gn2; gdx[2] = dy[0] * gn1 - dy[1] * gn0;
    gdy[0] = gn1 * dx[2] - gn2 * dx[1]; gdy[1] = gn2 * dx[0] - gn0 * dx[2]; gdy[2] = gn0 * dx[1] - gn1 * dx[0];
    return true;
}

__global__ void __launch_bounds__(256)
loss_grad_kernel(const LossCfg g, const float* __restrict__ color, const float* __restrict__ allmap,
                 const float* __restrict__ wvt, const float* __restrict__ target, const float* __restrict__ vis2d,
                 const float* __restrict__ mask_gt, const float* __restrict__ mask_wt, const float* __restrict__ bkgd,
                 const float* __restrict__ surf_depth, float* __restrict__ loss_terms, float* __restrict__ g_color,
                 float* __restrict__ g_allmap, float* __restrict__ g_bkgd) {
    const int x = blockIdx.x * 32 + (threadIdx.x & 31), y = blockIdx.y * 8 + (threadIdx.x >> 5), f = blockIdx.z;
    const int W = g.W, H = g.H;
    const size_t N = (size_t)W * H;
    const float invN = 1.0f / (float)N;
    float part[7] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};   // rgb, mask, normal, dist loss terms; dL/dbkgd
    if (x < W && y < H) {
        const size_t p = (size_t)y * W + x;
        const float* am = allmap + (size_t)f * 8 * N;
        const float* sd = surf_depth + (size_t)f * N;
        const PostCam c = load_cam(wvt + (size_t)f * 16, W, H, g.tanx, g.tany);
        const float a = am[N + p], d0 = am[p], m5 = am[5 * N + p];
        const float n0 = am[2 * N + p], n1 = am[3 * N + p], n2 = am[4 * N + p];
        float g_acc = 0.f;
        // ---- rgb: masked L1 (model.py:674-692), after the optional learnable-background composite
        const float v = vis2d ? (vis2d[(size_t)f * N + p] > 0.f ? 1.f : 0.f) : 1.f;
        const float s_rgb = g.w_rgb * invN * (1.0f / 3.0f);
#pragma unroll
        for (int ch = 0; ch < 3; ch++) {
            float col = color[((size_t)f * 3 + ch) * N + p];
            const float bk = bkgd ? __ldg(bkgd + ch) : 0.f;
            col += (1.f - a) * bk;
            const float d = col - target[((size_t)f * 3 + ch) * N + p];
            part[0] += fabsf(d) * v * s_rgb;
            const float gc = (d > 0.f ? 1.f : (d < 0.f ? -1.f : 0.f)) * v * s_rgb;
            g_color[((size_t)f * 3 + ch) * N + p] = gc;
            part[4 + ch] = gc * (1.f - a);
            g_acc -= gc * bk;
        }
        // ---- mask: (acc - gt)^2 * balance weight (model.py:649-653)
        if (mask_gt) {
            const float wt = mask_wt ? mask_wt[(size_t)f * N + p] : 1.f;
            const float d = a - mask_gt[(size_t)f * N + p];
            part[1] = g.w_mask * d * d * wt * invN;
            g_acc += g.w_mask * 2.f * d * wt * invN;
        }
        // ---- normal consistency: lambda_n * mean(1 - <rend_normal, surf_normal>)   (model.py:817-842)
        const float r0 = c.R[0] * n0 + c.R[1] * n1 + c.R[2] * n2, r1 = c.R[3] * n0 + c.R[4] * n1 + c.R[5] * n2,
                    r2 = c.R[6] * n0 + c.R[7] * n1 + c.R[8] * n2;
        float s0 = 0.f, s1 = 0.f, s2 = 0.f;           // surf_normal at this pixel
        {
            float dx[3], dy[3], n[3];
            if (stencil(c, W, H, x, y, sd, dx, dy, n)) {
                const float inv = a / fmaxf(sqrtf(n[0] * n[0] + n[1] * n[1] + n[2] * n[2]), 1e-12f);
                s0 = n[0] * inv; s1 = n[1] * inv; s2 = n[2] * inv;
            }
        }
        const float sN = g.lambda_normal * invN;
        part[2] = sN * (1.f - (r0 * s0 + r1 * s1 + r2 * s2));
        // d/d rend_normal = -sN * surf_normal  ->  allmap[2..4] through W^T
        const float q0 = -sN * s0, q1 = -sN * s1, q2 = -sN * s2;
        float* ga = g_allmap + (size_t)f * 8 * N;
        ga[2 * N + p] = c.R[0] * q0 + c.R[3] * q1 + c.R[6] * q2;
        ga[3 * N + p] = c.R[1] * q0 + c.R[4] * q1 + c.R[7] * q2;
        ga[4 * N + p] = c.R[2] * q0 + c.R[5] * q1 + c.R[8] * q2;
        // d/d surf_depth[p]: gather from the four stencils this pixel takes part in
        float gp0 = 0.f, gp1 = 0.f, gp2 = 0.f, gdx[3], gdy[3];
        if (g.lambda_normal != 0.f) {
            if (stencil_loss_vjp(c, g, x, y - 1, sd, am, N, -sN, gdx, gdy)) { gp0 += gdx[0]; gp1 += gdx[1]; gp2 += gdx[2]; }
            if (stencil_loss_vjp(c, g, x, y + 1, sd, am, N, -sN, gdx, gdy)) { gp0 -= gdx[0]; gp1 -= gdx[1]; gp2 -= gdx[2]; }
            if (stencil_loss_vjp(c, g, x - 1, y, sd, am, N, -sN, gdx, gdy)) { gp0 += gdy[0]; gp1 += gdy[1]; gp2 += gdy[2]; }
            if (stencil_loss_vjp(c, g, x + 1, y, sd, am, N, -sN, gdx, gdy)) { gp0 -= gdy[0]; gp1 -= gdy[1]; gp2 -= gdy[2]; }
        }
        float rx, ry, rz;
        ray_dir(c, x, y, rx, ry, rz);
        const float gsd = gp0 * rx + gp1 * ry + gp2 * rz;
        const float gex = gsd * (1.f - g.depth_ratio), gmed = gsd * g.depth_ratio;
        const float q = d0 / a;
        const bool qfin = isfinite(q);                  // nan_to_num passes the gradient only where its input is finite
        ga[p] = qfin ? gex / a : 0.f;
        ga[N + p] = g_acc + (qfin ? -gex * d0 / (a * a) : 0.f);
        ga[5 * N + p] = isfinite(m5) ? gmed : 0.f;
        // ---- distortion: lambda_d * mean(rend_dist)
        part[3] = g.lambda_dist * am[6 * N + p] * invN;
        ga[6 * N + p] = g.lambda_dist * invN;
        ga[7 * N + p] = 0.f;
    }
    // ---- block reduction of the loss terms (and dL/dbkgd): 7 atomics per block
    __shared__ float red[8][7];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int i = 0; i < 7; i++) {
        float vsum = part[i];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) vsum += __shfl_xor_sync(0xffffffffu, vsum, o);
        if (lane == 0) red[warp][i] = vsum;
    }
    __syncthreads();
    if (threadIdx.x < 7) {
        float t = 0.f;
#pragma unroll
        for (int w = 0; w < 8; w++) t += red[w][threadIdx.x];
        if (threadIdx.x < 4) atomicAdd(loss_terms + (size_t)f * 4 + threadIdx.x, t);
        else if (g_bkgd) atomicAdd(g_bkgd + (size_t)f * 3 + (threadIdx.x - 4), t);
    }
}

}  // namespace

extern "C" {

SR_API int sr_render_loss_batch(int32_t M, int32_t W, int32_t H, float tanx, float tany, float depth_ratio,
                                const float* color, const float* allmap, const float* world_view_transform,
                                const float* target_rgb, const float* vis2d, const float* mask_gt, const float* mask_wt,
                                const float* learnable_bkgd, float w_rgb, float w_mask, float lambda_normal, float lambda_dist,
                                float* loss_terms, float* dL_dcolor, float* dL_dallmap, float* dL_dbkgd, float* surf_depth_scratch,
                                void* stream_) {
    if (M < 1 || W <= 0 || H <= 0 || !color || !allmap || !world_view_transform || !target_rgb || !loss_terms || !dL_dcolor ||
        !dL_dallmap || !surf_depth_scratch)
        return SR_EINVAL;
    cudaStream_t s = (cudaStream_t)stream_;
    LossCfg g{W, H, tanx, tany, depth_ratio, w_rgb, w_mask, lambda_normal, lambda_dist};
    dim3 grid((W + 31) / 32, (H + 7) / 8, M);
    if (cudaMemsetAsync(loss_terms, 0, (size_t)M * 4 * sizeof(float), s) != cudaSuccess) return SR_ECUDA;
    if (dL_dbkgd && cudaMemsetAsync(dL_dbkgd, 0, (size_t)M * 3 * sizeof(float), s) != cudaSuccess) return SR_ECUDA;
    {
        ProfileScope ps("loss_depth", s);
        loss_depth_kernel<<<grid, 256, 0, s>>>(g, allmap, surf_depth_scratch);
    }
    {
        ProfileScope ps("loss_grad", s);
        loss_grad_kernel<<<grid, 256, 0, s>>>(g, color, allmap, world_view_transform, target_rgb, vis2d, mask_gt, mask_wt,
                                             learnable_bkgd, surf_depth_scratch, loss_terms, dL_dcolor, dL_dallmap, dL_dbkgd);
    }
    sr_count_launch(2);
    return cudaGetLastError() == cudaSuccess ? 0 : SR_ECUDA;
}

}  // extern "C"
