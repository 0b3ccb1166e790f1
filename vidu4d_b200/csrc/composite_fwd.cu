// composite_fwd.cu -- per-tile front-to-back alpha composite of colour / depth / normal / distortion.
//
// Replaces (behaviour, not code) of the reference's forward renderCUDA, RAST/cuda_rasterizer/forward.cu:265-463.
//
// Blackwell design (DESIGN.md "composite"):
//   * one CTA per 16x16 tile (same tiles as the reference so `ranges`/`n_contrib` mean the same thing),
//     8 warps, each warp owns an 8x4-pixel sub-tile;
//   * the tile's sorted instance records (80 B each, contiguous) stream into shared memory with ONE
//     cp.async.bulk (TMA 1-D) per 128-instance chunk, double-buffered on mbarriers;
//   * every warp tests 32 instances at a time against its sub-tile with the conservative per-instance
//     cull rectangle (lane = instance), ballots, and then only evaluates the survivors (lane = pixel).
//     Culling never changes a result: a culled (pixel, instance) pair provably has alpha < 1/255.
//   * per-pair arithmetic is written with explicit-rounding intrinsics in the contraction pattern of the
//     reference's sm_100a SASS, so colour/depth/alpha/normal/median planes are bit-identical to it.
#include "common.cuh"

namespace {

constexpr int CHUNK = 128;
constexpr int STAGES = 2;

__device__ __forceinline__ float fm(float a, float b) { return __fmul_rn(a, b); }
__device__ __forceinline__ float fa(float a, float b) { return __fadd_rn(a, b); }
__device__ __forceinline__ float ff(float a, float b, float c) { return __fmaf_rn(a, b, c); }

__global__ void __launch_bounds__(256)
composite_fwd_kernel(const uint2* __restrict__ ranges, const float4* __restrict__ irec, int W, int H,
                     const float* __restrict__ bg, float* __restrict__ final_T, uint32_t* __restrict__ n_contrib,
                     float* __restrict__ out_color, float* __restrict__ out_others, uint32_t* __restrict__ tile_last) {
    __shared__ __align__(128) float4 stage[STAGES][CHUNK * 5];
    __shared__ __align__(8) uint64_t full_bar[STAGES];
    __shared__ uint32_t tile_max_s;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int tile = blockIdx.y * gridDim.x + blockIdx.x;
    const uint2 range = ranges[tile];
    const int len = (int)(range.y - range.x);
    const int nchunks = (len + CHUNK - 1) / CHUNK;

    // warp -> 8x4 sub-tile, lane -> pixel
    const int sx0 = (warp & 1) * 8, sy0 = (warp >> 1) * 4;
    const int lx = sx0 + (lane & 7), ly = sy0 + (lane >> 3);
    const int pix_x = blockIdx.x * SR_TILE + lx, pix_y = blockIdx.y * SR_TILE + ly;
    const bool inside = pix_x < W && pix_y < H;
    const float pixx = (float)pix_x + 0.5f, pixy = (float)pix_y + 0.5f;

    if (tid == 0) {
        mbar_init(&full_bar[0], 1);
        mbar_init(&full_bar[1], 1);
        fence_mbar_init();
        tile_max_s = 0;
    }
    __syncthreads();
    const float4* src = irec + (size_t)range.x * 5;
    if (tid == 0) {
        for (int c = 0; c < STAGES && c < nchunks; c++) {
            const uint32_t bytes = (uint32_t)min(CHUNK, len - c * CHUNK) * 80u;
            mbar_expect_tx(&full_bar[c], bytes);
            bulk_g2s(stage[c], src + (size_t)c * CHUNK * 5, bytes, &full_bar[c]);
        }
    }

    float T = 1.0f, C0 = 0.f, C1 = 0.f, C2 = 0.f, D = 0.f, N0 = 0.f, N1 = 0.f, N2 = 0.f;
    float dist1 = 0.f, dist2 = 0.f, distortion = 0.f, median_depth = 0.f, median_weight = 0.f;
    uint32_t median_contributor = 0, last_contributor = 0;
    bool done = !inside;

    for (int c = 0; c < nchunks; c++) {
        const int s = c & 1;
        mbar_wait(&full_bar[s], (uint32_t)((c >> 1) & 1));
        const int cnt = min(CHUNK, len - c * CHUNK);
        const float4* S = stage[s];
        if (!__all_sync(0xffffffffu, done)) {
            for (int b = 0; b < cnt; b += 32) {
                const int j = b + lane;
                uint32_t cull = 0;
                if (j < cnt) cull = __float_as_uint(S[j * 5 + 4].w);
                const int cx0 = cull & 15, cx1 = (cull >> 4) & 15, cy0 = (cull >> 8) & 15, cy1 = (cull >> 12) & 15;
                const bool hit = (cull >> 16) && cx0 <= sx0 + 7 && cx1 >= sx0 && cy0 <= sy0 + 3 && cy1 >= sy0;
                uint32_t m = __ballot_sync(0xffffffffu, hit);
                while (m) {
                    const int jj = b + __ffs(m) - 1;
                    m &= m - 1;
                    if (done) continue;
                    const float4 r0 = S[jj * 5], r1 = S[jj * 5 + 1], r2 = S[jj * 5 + 2];
                    // T rows: Tu=(r0.x,r0.y,r0.z) Tv=(r0.w,r1.x,r1.y) Tw=(r1.z,r1.w,r2.x); xy=(r2.y,r2.z); opac=r2.w
                    const float kx = ff(pixx, r1.z, -r0.x), ky = ff(pixx, r1.w, -r0.y), kz = ff(pixx, r2.x, -r0.z);
                    const float lx_ = ff(pixy, r1.z, -r0.w), ly_ = ff(pixy, r1.w, -r1.x), lz_ = ff(pixy, r2.x, -r1.y);
                    const float pz = ff(kx, ly_, -fm(ky, lx_));
                    if (pz == 0.0f) continue;
                    const float ppx = ff(ky, lz_, -fm(kz, ly_));
                    const float ppy = ff(kz, lx_, -fm(kx, lz_));
                    const float sx = __fdiv_rn(ppx, pz), sy = __fdiv_rn(ppy, pz);
                    const float rho3d = ff(sx, sx, fm(sy, sy));
                    const float dx = fa(r2.y, -pixx), dy = fa(r2.z, -pixy);
                    const float q2 = ff(dx, dx, fm(dy, dy));
                    const float rho2d = fa(q2, q2);
                    const float rho = fminf(rho3d, rho2d);
                    const float depth = (rho3d <= rho2d) ? fa(r2.x, ff(r1.z, sx, fm(r1.w, sy))) : r2.x;
                    if (depth < 0.2f) continue;
                    const float power = fm(rho, -0.5f);
                    if (power > 0.0f) continue;
                    const float alpha = fminf(0.99f, fm(r2.w, expf(power)));
                    if (alpha < 1.0f / 255.0f) continue;
                    const float test_T = fm(T, fa(1.0f, -alpha));
                    if (test_T < 0.0001f) { done = true; continue; }
                    const float4 r3 = S[jj * 5 + 3], r4 = S[jj * 5 + 4];
                    const uint32_t contributor = (uint32_t)(c * CHUNK + jj + 1);
                    const float A = fa(1.0f, -T);
                    // forward.cu:412 evaluates the depth mapping in double (FAR/NEAR_PLANE are double literals)
                    const double dd = (double)depth;
                    const float mdep = (float)(fma(dd, 100.0, -20.0) / (dd * 99.8));
                    const float mm = fm(mdep, mdep);
                    const float error = ff(-dist1, fa(mdep, mdep), ff(A, mm, dist2));
                    distortion = ff(T, fm(alpha, error), distortion);
                    if (T > 0.5f) { median_depth = depth; median_weight = fm(T, alpha); median_contributor = contributor; }
                    N0 = ff(T, fm(r3.x, alpha), N0);
                    N1 = ff(T, fm(r3.y, alpha), N1);
                    N2 = ff(T, fm(r3.z, alpha), N2);
                    D = ff(T, fm(depth, alpha), D);
                    dist1 = ff(T, fm(alpha, mdep), dist1);
                    dist2 = ff(T, fm(alpha, mm), dist2);
                    C0 = ff(T, fm(alpha, r3.w), C0);
                    C1 = ff(T, fm(alpha, r4.x), C1);
                    C2 = ff(T, fm(alpha, r4.y), C2);
                    T = test_T;
                    last_contributor = contributor;
                }
                if (__all_sync(0xffffffffu, done)) break;
            }
        }
        const int active = __syncthreads_count(!done);
        if (active == 0) {
            // drain the one copy that may still be in flight before the CTA (and its smem) goes away
            if (c + 1 < nchunks) mbar_wait(&full_bar[(c + 1) & 1], (uint32_t)(((c + 1) >> 1) & 1));
            break;
        }
        if (tid == 0 && c + STAGES < nchunks) {
            const int cn = c + STAGES;
            const uint32_t bytes = (uint32_t)min(CHUNK, len - cn * CHUNK) * 80u;
            mbar_expect_tx(&full_bar[s], bytes);
            bulk_g2s(stage[s], src + (size_t)cn * CHUNK * 5, bytes, &full_bar[s]);
        }
    }

    // per-tile maximum of last_contributor: lets the backward start its reverse walk there
    uint32_t wmax = last_contributor;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) wmax = max(wmax, __shfl_xor_sync(0xffffffffu, wmax, o));
    if (lane == 0 && wmax) atomicMax(&tile_max_s, wmax);
    __syncthreads();
    if (tid == 0) tile_last[tile] = tile_max_s;

    if (inside) {
        const size_t N = (size_t)W * H, pid = (size_t)W * pix_y + pix_x;
        final_T[pid] = T;
        final_T[pid + N] = dist1;
        final_T[pid + 2 * N] = dist2;
        n_contrib[pid] = last_contributor;
        n_contrib[pid + N] = median_contributor;
        out_color[pid] = ff(T, __ldg(bg), C0);
        out_color[pid + N] = ff(T, __ldg(bg + 1), C1);
        out_color[pid + 2 * N] = ff(T, __ldg(bg + 2), C2);
        out_others[pid] = D;
        out_others[pid + N] = fa(1.0f, -T);
        out_others[pid + 2 * N] = N0;
        out_others[pid + 3 * N] = N1;
        out_others[pid + 4 * N] = N2;
        out_others[pid + 5 * N] = median_depth;
        out_others[pid + 6 * N] = distortion;
        out_others[pid + 7 * N] = median_weight;
    }
}

}  // namespace

cudaError_t launch_composite_fwd(const FwdArgs& a) {
    dim3 grid(a.il.tiles_x, a.il.tiles_y, 1);
    composite_fwd_kernel<<<grid, 256, 0, a.stream>>>(
        (const uint2*)(a.img + a.il.ranges), (const float4*)(a.bin + a.bl.inst_rec), a.cam.W, a.cam.H, a.cam.bg,
        (float*)(a.img + a.il.final_T), (uint32_t*)(a.img + a.il.n_contrib), a.out_color, a.out_others,
        (uint32_t*)(a.img + a.il.tile_last));
    sr_count_launch();
    return cudaGetLastError();
}
