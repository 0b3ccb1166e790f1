// composite_bwd.cu -- per-tile reverse walk: dL/d{T, mean2D, opacity, colour, normal} per surfel.
//
// Replaces (behaviour, not code) of the reference's backward renderCUDA, RAST/cuda_rasterizer/backward.cu:143-449.
//
// The reference issues up to 16 global float atomics per contributing (pixel, surfel) pair
// (backward.cu:345-446).  Here (DESIGN.md "composite backward"):
//   * same tiling and warp-autonomous TMA-fed instance stream as the forward, walked back to front, starting at the
//     sub-tile's deepest contributor recorded by the forward (no work on the occluded tail);
//   * the survivors of a stage are NOT re-derived: the forward left, per stage and 2x2 group, the mask of instances
//     that contributed to a pixel of the group -- exactly the pairs worth evaluating;
//   * the 16 per-pair gradient components are reduced across the pixels of a group with a recursive-halving
//     butterfly and leave the SM as ONE 16-byte vector reduction per lane (REDG.ADD.F32x4) into the surfel's
//     accumulator, whose first 16 floats are in butterfly order -- instead of 16 scalar atomics per contributing pixel.
#include "composite_common.cuh"

namespace {
using namespace comp;

// Recursive-halving sum of v[0..15] over each group of GL lanes (GL = 32, 16, 8, 4 or 2).  Every step exchanges half
// of the still-live values with the lane `off` away; afterwards a lane holds NV = max(1, 32/GL ... ) totals:
//   GL=32 or 16: 1 value,  GL=8: 2,  GL=4: 4,  GL=2: 8 values, of components  base(lane) + i,  i < NV  (see comp_base).
template <int GL>
__device__ __forceinline__ void butterfly16(float (&v)[16], int lane) {
    int n = 16;
#pragma unroll
    for (int off = GL / 2; off >= 1; off >>= 1) {
        if (n > 1) {
            n >>= 1;
            const bool hi = lane & off;
#pragma unroll
            for (int i = 0; i < 8; i++) {
                if (i < n) {
                    const float keep = hi ? v[i + n] : v[i], send = hi ? v[i] : v[i + n];
                    v[i] = keep + __shfl_xor_sync(0xffffffffu, send, off);
                }
            }
        } else {
            v[0] += __shfl_xor_sync(0xffffffffu, v[0], off);
        }
    }
}
// values a lane holds after butterfly16<GL>, and the component index of its first one
template <int GL> struct BflyOut { static constexpr int NV = GL >= 16 ? 1 : 16 / GL; };
template <int GL>
__device__ __forceinline__ int comp_base(int lane) {
    int c = 0, n = 16;
#pragma unroll
    for (int off = GL / 2; off >= 1; off >>= 1) {
        if (n > 1) { n >>= 1; if (lane & off) c += n; }
    }
    return c;
}

template <int G, int MINB = 20, int ABL = 0>
__global__ void __launch_bounds__(32 * WPC, MINB / WPC)
composite_bwd_kernel(const FrameStrides fs, const float* __restrict__ grad_scale, const uint2* __restrict__ ranges,
                     const uint32_t* __restrict__ tile_order, int n_items, int tiles_x,
                     const float4* __restrict__ irec, int W, int H,
                     const float* __restrict__ bg, const float* __restrict__ final_Ts,
                     const uint32_t* __restrict__ n_contrib, const uint32_t* __restrict__ sub_last,
                     const float* __restrict__ dL_dpixels, const float* __restrict__ dL_dothers,
                     const uint32_t* __restrict__ contrib_masks, float* __restrict__ sgrad) {
    // WPC independent warps per CTA (the SM holds at most 32 CTAs): work item = (tile, 8x4 sub-tile), tiles in
    // longest-list-first order; the warps of a CTA share nothing
    __shared__ __align__(128) float4 st_all[WPC][NST][WB * REC4];
    __shared__ __align__(8) uint64_t bar_all[WPC][NST];
    float4 (*st)[WB * REC4] = st_all[threadIdx.x >> 5];
    uint64_t* bar = bar_all[threadIdx.x >> 5];

    const int lane = threadIdx.x & 31;
    const int f = (int)(blockIdx.x % (unsigned)fs.frames);
    const int item = (int)(blockIdx.x / (unsigned)fs.frames) * WPC + (threadIdx.x >> 5);
    if (item >= n_items) return;
    ranges = fr(ranges, fs.img, f); tile_order = fr(tile_order, fs.img, f); irec = fr(irec, fs.bin, f);
    final_Ts = fr(final_Ts, fs.img, f); n_contrib = fr(n_contrib, fs.img, f); sub_last = fr(sub_last, fs.img, f);
    dL_dpixels = fr(dL_dpixels, fs.dcolor, f); dL_dothers = fr(dL_dothers, fs.dothers, f);
    contrib_masks = fr(contrib_masks, fs.bin, f); sgrad = fr(sgrad, fs.geom, f);
    const int warp = item & 7;
    const int tile = (int)tile_order[item >> 3];
    const int tile_x = tile % tiles_x, tile_y = tile / tiles_x;
    const uint2 range = ranges[tile];
    // nothing beyond this sub-tile's deepest contributor matters
    const int len = min((int)(range.y - range.x), (int)sub_last[tile * 8 + warp]);
    const int nb = (len + WB - 1) / WB;
    if (nb == 0) return;

    using GS = GroupShape<G>;
    const int sx0 = (warp & 1) * 8, sy0 = (warp >> 1) * 4;
    const int g = lane / GS::GL, l = lane % GS::GL;
    const int lx = GS::block_x(g) + l % GS::BW, ly = GS::block_y(g) + l / GS::BW;      // pixel inside the sub-tile

    const float4* src = irec + (size_t)range.x * REC4;
    // the k-th consumed batch (k = 0,1,..) is b = nb-1-k; it lives in stage k % NST
    auto issue = [&](int k) {   // lane 0 only
        const int b = nb - 1 - k, s = k % NST;
        const uint32_t bytes = (uint32_t)min(WB, len - b * WB) * 80u;
        mbar_expect_tx(&bar[s], bytes);
        bulk_g2s(st[s], src + (size_t)b * WB * REC4, bytes, &bar[s]);
    };
    if (lane == 0) {
#pragma unroll
        for (int s = 0; s < NST; s++) mbar_init(&bar[s], 1);
        fence_mbar_init();
        for (int k = 0; k < NST && k < nb; k++) issue(k);
    }
    __syncwarp();

    BwdPixel px;
    px.load(tile_x * SR_TILE + sx0 + lx, tile_y * SR_TILE + sy0 + ly, W, H, bg, final_Ts, n_contrib, dL_dpixels, dL_dothers, grad_scale);

    // The forward recorded, per stage and pixel, which instances contributed (common.cuh bin_layout: contrib).  A group
    // walks the union of its pixels' masks -- exactly the instances worth visiting: no cull test, no trimming, and a
    // lane knows beforehand whether ITS pixel takes part, so the alpha / depth tests of the forward are not repeated.
    const uint32_t* cm_in = contrib_masks + (((size_t)(range.x >> 5) + tile) * 8 + warp) * 32 + (ly * 8 + lx);
    uint32_t next_mask = __ldg(cm_in + (size_t)(nb - 1) * SR_CONTRIB_STAGE_WORDS);
    constexpr uint32_t GMASK = GS::GL == 32 ? 0xffffffffu : ((1u << GS::GL) - 1u);

    for (int k = 0; k < nb; k++) {
        const int b = nb - 1 - k, s = k % NST;
        const uint32_t own = next_mask;                       // instances of this stage that reached MY pixel
        if (b > 0) next_mask = __ldg(cm_in + (size_t)(b - 1) * SR_CONTRIB_STAGE_WORDS);   // lands during this stage
        uint32_t mym = own;                                   // ... that reached any pixel of my group
#pragma unroll
        for (int o = GS::GL / 2; o > 0; o >>= 1) mym |= __shfl_xor_sync(0xffffffffu, mym, o);
        mbar_wait(&bar[s], (uint32_t)((k / NST) & 1));
        const float4* S = st[s];
        while (__any_sync(0xffffffffu, mym != 0u)) {
            const bool act = mym != 0u;                       // my group still has an instance to visit
            const int jj = act ? 31 - __clz(mym) : 0;
            mym &= ~(1u << jj) | (act ? 0u : 0xffffffffu);
            const int pos = b * WB + jj;                      // == `contributor` after the decrement
            const bool contrib = act && ((own >> jj) & 1u);   // ... and it reached MY pixel
            float v[16];
#pragma unroll
            for (int i = 0; i < 16; i++) v[i] = 0.f;
            float m2x = 0.f, m2y = 0.f;
            bool lowpass = false;
            // ABL (diagnostic builds only, SURFEL_BWD_ABL): 1 = no reductions, 2 = no reductions + conflict-free
            // (broadcast) record loads, 3 = one of the four vector reductions only
            const float4* Rj = (ABL == 2) ? S + (jj & 1) * REC4 : S + jj * REC4;
            if (contrib) px.pair(Rj, pos, v, m2x, m2y, lowpass);
            if constexpr (GS::GL == 1) {
                if constexpr (ABL == 1 || ABL == 2) {
                    if (contrib && v[0] + v[5] + v[9] + v[12] + v[15] + m2x == 1.2345e-30f) red_pixel(sgrad, S + jj * REC4, v, m2x, m2y, lowpass);
                } else if constexpr (ABL == 3) {
                    if (contrib) {
                        float* gp = sgrad + (size_t)__float_as_uint(S[jj * REC4 + 4].z) * SR_GRAD_FLOATS;
                        const float w4[16] = {v[0] + v[4] + v[8] + v[12], v[1] + v[5] + v[9] + v[13], v[2] + v[6] + v[10] + v[14], v[3] + v[7] + v[11] + v[15] + m2x + m2y};
                        red_add<4>(gp, w4);
                    }
                } else
                if (contrib) red_pixel(sgrad, S + jj * REC4, v, m2x, m2y, lowpass);
            } else {
                const uint32_t cb = __ballot_sync(0xffffffffu, contrib);
                if (cb) {
                    // this lane's group has a contributor?  (its REDs are skipped otherwise; the shuffles are warp-wide)
                    const bool gact = (cb >> (g * GS::GL)) & GMASK;
                    const uint32_t id = __float_as_uint(S[jj * REC4 + 4].z);
                    float* gp = sgrad + (size_t)id * SR_GRAD_FLOATS;
                    butterfly16<GS::GL>(v, lane);
                    constexpr int NV = BflyOut<GS::GL>::NV;
                    const int c0 = comp_base<GS::GL>(lane);
                    if (gact && (GS::GL < 32 || (lane & 1) == 0)) red_add<NV>(gp + c0, v);
                    if (__any_sync(0xffffffffu, lowpass)) {
                        // 2-value halving butterfly inside the group: its lane 0 ends with sum(m2x), lane GL/2 with sum(m2y)
                        const bool hi = lane & (GS::GL / 2);
                        float w = (hi ? m2y : m2x) + __shfl_xor_sync(0xffffffffu, hi ? m2x : m2y, GS::GL / 2);
#pragma unroll
                        for (int o = GS::GL / 4; o > 0; o >>= 1) w += __shfl_xor_sync(0xffffffffu, w, o);
                        if (gact && (l & (GS::GL / 2 - 1)) == 0) atomicAdd(gp + SR_G_M2D + (l >= GS::GL / 2 ? 1 : 0), w);
                    }
                }
            }
        }
        __syncwarp();
        if (lane == 0 && k + NST < nb) issue(k + NST);
    }
}

}  // namespace

cudaError_t launch_composite_bwd(const BwdArgs& a) {
    cudaError_t e = sr_memset_frames(a.geom + a.gl.sgrad, (size_t)a.fs.geom, (size_t)(a.cam.P > 0 ? a.cam.P : 1) * SR_GRAD_FLOATS * 4,
                                     a.fs.frames, a.stream);
    if (e != cudaSuccess) return e;
    ProfileScope ps("composite_bwd", a.stream);
    auto launch = [&](auto kern) {
        kern<<<((a.il.tiles * 8 + comp::WPC - 1) / comp::WPC) * a.fs.frames, 32 * comp::WPC, 0, a.stream>>>(
            a.fs, a.grad_scale, (const uint2*)(a.img + a.il.ranges), (const uint32_t*)(a.img + a.il.tile_order), a.il.tiles * 8, a.il.tiles_x,
            (const float4*)(a.bin + a.bl.inst_rec), a.cam.W, a.cam.H, a.cam.bg,
            (const float*)(a.img + a.il.final_T), (const uint32_t*)(a.img + a.il.n_contrib),
            (const uint32_t*)(a.img + a.il.tile_last), a.dL_dcolor, a.dL_dothers,
            (const uint32_t*)(a.bin + a.bl.contrib), (float*)(a.geom + a.gl.sgrad));
    };
    // the forward stores per-PIXEL contribution masks, so the backward's grouping is independent of the forward's
    const int G = comp::bwd_groups_from_env();
    switch (G) {
        case 1: launch(composite_bwd_kernel<1>); break;
        case 2: launch(composite_bwd_kernel<2>); break;
        case 4: launch(composite_bwd_kernel<4>); break;
        case 16: launch(composite_bwd_kernel<16>); break;
        case 32: {
            // register target: 20 CTAs/SM -> 78 registers, 28 -> 71 (no spills, 0.186 -> 0.181 ms), 32 -> 64 with spills (0.201)
            static const int occ = [] { const char* e = getenv("SURFEL_BWD_OCC"); return e ? atoi(e) : 28; }();
            if (occ == 20) launch(composite_bwd_kernel<32, 20>);
            else if (occ == 32) launch(composite_bwd_kernel<32, 32>);
            else {
                static const int abl = [] { const char* e = getenv("SURFEL_BWD_ABL"); return e ? atoi(e) : 0; }();
                if (abl == 1) launch(composite_bwd_kernel<32, 28, 1>);
                else if (abl == 2) launch(composite_bwd_kernel<32, 28, 2>);
                else if (abl == 3) launch(composite_bwd_kernel<32, 28, 3>);
                else launch(composite_bwd_kernel<32, 28>);
            }
            break;
        }
        default: {
            // CTAs/SM the register allocation targets: 24 -> 80 registers with spilled values reloaded on every
            // contributing iteration; 20 -> 96 registers, no spills: 0.372 -> 0.337 ms on the headline frame (16: 0.342)
            static const int occ = [] { const char* e = getenv("SURFEL_BWD_OCC"); return e ? atoi(e) : 20; }();
            if (occ == 24) launch(composite_bwd_kernel<8, 24>);
            else if (occ == 16) launch(composite_bwd_kernel<8, 16>);
            else launch(composite_bwd_kernel<8, 20>);
            break;
        }
    }
    sr_count_launch();
    return cudaGetLastError();
}
