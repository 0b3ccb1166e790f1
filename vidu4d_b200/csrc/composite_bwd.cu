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

// v[] order: 0..8 dT, 9 dopacity, 10..12 dcolor, 13..15 dnormal == slots 0..15 of the per-surfel accumulator
static_assert(SR_G_T == 0 && SR_G_OPAC == 9 && SR_G_COLOR == 10 && SR_G_NORMAL == 13, "butterfly order == accumulator order");
// NV consecutive totals -> one vector reduction (red.global.add.v2/v4.f32, sm_90+; address 4*NV-byte aligned)
template <int NV>
__device__ __forceinline__ void red_add(float* p, const float (&v)[16]) {
    if constexpr (NV == 16 || NV == 8) {
#pragma unroll
        for (int q = 0; q < NV; q += 4)
            asm volatile("red.global.v4.f32.add [%0], {%1, %2, %3, %4};" :: "l"(p + q), "f"(v[q]), "f"(v[q + 1]), "f"(v[q + 2]), "f"(v[q + 3]) : "memory");
    } else if constexpr (NV == 4) {
        asm volatile("red.global.v4.f32.add [%0], {%1, %2, %3, %4};" :: "l"(p), "f"(v[0]), "f"(v[1]), "f"(v[2]), "f"(v[3]) : "memory");
    } else if constexpr (NV == 2) {
        asm volatile("red.global.v2.f32.add [%0], {%1, %2};" :: "l"(p), "f"(v[0]), "f"(v[1]) : "memory");
    } else {
        atomicAdd(p, v[0]);
    }
}

// Per-pixel backward state (backward.cu:192-249) and the per-pair gradient arithmetic.
struct BwdPixel {
    float pixx, pixy;
    float T_final, T, final_D, final_D2, final_A, bg_dot_dpixel;
    int median_contributor;
    float dpix0, dpix1, dpix2, dL_ddepth, dL_daccum, dL_dreg, dn0, dn1, dn2, dL_dmedian_depth, dL_dmax_dweight;
    // "what lies behind the current contributor" accumulators (backward.cu:253-262 keeps last_alpha / last_color /
    // last_depth / last_normal and folds them in at the START of the next contributor; folding them in at the END of
    // the current one is the same arithmetic in the same order and needs 8 fewer live registers)
    float accum_rec0, accum_rec1, accum_rec2, accum_depth_rec, accum_alpha_rec, last_dL_dT, accum_n0, accum_n1, accum_n2;

    __device__ __forceinline__ void load(int pix_x, int pix_y, int W, int H, const float* __restrict__ bg,
                                         const float* __restrict__ final_Ts, const uint32_t* __restrict__ n_contrib,
                                         const float* __restrict__ dL_dpixels, const float* __restrict__ dL_dothers) {
        const bool inside = pix_x < W && pix_y < H;
        const size_t N = (size_t)W * H, pid = (size_t)W * pix_y + pix_x;
        pixx = (float)pix_x + 0.5f; pixy = (float)pix_y + 0.5f;
        T_final = inside ? final_Ts[pid] : 0.f;
        T = T_final;
        median_contributor = inside ? (int)n_contrib[pid + N] : 0;
        dpix0 = dpix1 = dpix2 = dL_ddepth = dL_daccum = dL_dreg = dn0 = dn1 = dn2 = dL_dmedian_depth = dL_dmax_dweight = 0.f;
        if (inside) {
            dpix0 = dL_dpixels[pid]; dpix1 = dL_dpixels[pid + N]; dpix2 = dL_dpixels[pid + 2 * N];
            dL_ddepth = dL_dothers[pid];
            dL_daccum = dL_dothers[pid + N];
            dn0 = dL_dothers[pid + 2 * N]; dn1 = dL_dothers[pid + 3 * N]; dn2 = dL_dothers[pid + 4 * N];
            dL_dmedian_depth = dL_dothers[pid + 5 * N];
            dL_dreg = dL_dothers[pid + 6 * N];
            dL_dmax_dweight = dL_dothers[pid + 7 * N];
        }
        final_D = inside ? final_Ts[pid + N] : 0.f;
        final_D2 = inside ? final_Ts[pid + 2 * N] : 0.f;
        final_A = 1.f - T_final;
        bg_dot_dpixel = __ldg(bg) * dpix0 + __ldg(bg + 1) * dpix1 + __ldg(bg + 2) * dpix2;
        accum_rec0 = accum_rec1 = accum_rec2 = accum_depth_rec = accum_alpha_rec = last_dL_dT = 0.f;
        accum_n0 = accum_n1 = accum_n2 = 0.f;
    }

    // One contributing (pixel, instance) pair; R = the instance's 5 x float4 record in shared memory, pos = its list
    // position.  The forward's contribution mask says this pair passed every test, so none is repeated.
    // v[0..8] dL/dT, v[9] dL/dopacity, v[10..12] dL/dcolour, v[13..15] dL/dnormal; (m2x, m2y) low-pass dL/dmean2D.
    __device__ __forceinline__ void pair(const float4* __restrict__ R, int pos, float (&v)[16], float& m2x, float& m2y,
                                         bool& lowpass) {
        const float4 r0 = R[0], r1 = R[1], r2 = R[2];
        // identical geometry / alpha arithmetic to the forward
        const float kx = ff(pixx, r1.z, -r0.x), ky = ff(pixx, r1.w, -r0.y), kz = ff(pixx, r2.x, -r0.z);
        const float lx_ = ff(pixy, r1.z, -r0.w), ly_ = ff(pixy, r1.w, -r1.x), lz_ = ff(pixy, r2.x, -r1.y);
        const float pz = ff(kx, ly_, -fm(ky, lx_));
        const float ppx = ff(ky, lz_, -fm(kz, ly_));
        const float ppy = ff(kz, lx_, -fm(kx, lz_));
        float sx, sy;
        div2_rn(ppx, ppy, pz, sx, sy);
        const float rho3d = ff(sx, sx, fm(sy, sy));
        const float dx = fa(r2.y, -pixx), dy = fa(r2.z, -pixy);
        const float q2 = ff(dx, dx, fm(dy, dy));
        const float rho2d = fa(q2, q2);
        const float rho = fminf(rho3d, rho2d);
        const float c_d = (rho3d <= rho2d) ? fa(r2.x, ff(r1.z, sx, fm(r1.w, sy))) : r2.x;
        const float power = fm(rho, -0.5f);
        const float G_ = expf(power);
        const float alpha = fminf(0.99f, fm(r2.w, G_));
        const float4 r3 = R[3], r4 = R[4];
        // one approximate reciprocal of (1 - alpha) serves the transmittance recurrence and the
        // background term (the reference divides twice, IEEE; the difference is ~1 ulp per step)
        const float r1ma = rcp_approx(1.f - alpha);
        T = T * r1ma;
        const float aT = alpha * T;
        float dL_dalpha = 0.f;
        // colour
        dL_dalpha += (r3.w - accum_rec0) * dpix0 + (r4.x - accum_rec1) * dpix1 + (r4.y - accum_rec2) * dpix2;
        v[10] = aT * dpix0; v[11] = aT * dpix1; v[12] = aT * dpix2;
        // distortion / median
        float dL_dz = 0.f, dL_dweight = 0.f;
        float m_d, dmd_dd;
        map_depth_vg(c_d, m_d, dmd_dd);
        if (pos == median_contributor - 1) { dL_dz += dL_dmedian_depth; dL_dweight += dL_dmax_dweight; }
        dL_dweight += (final_D2 + m_d * m_d * final_A - 2.f * m_d * final_D) * dL_dreg;
        dL_dalpha += dL_dweight - last_dL_dT;
        last_dL_dT = dL_dweight * alpha + (1.f - alpha) * last_dL_dT;
        const float dL_dmd = 2.0f * aT * (m_d * final_A - final_D) * dL_dreg;
        dL_dz += dL_dmd * dmd_dd;
        // depth, alpha
        dL_dalpha += (c_d - accum_depth_rec) * dL_ddepth;
        dL_dalpha += (1.f - accum_alpha_rec) * dL_daccum;
        // normal
        dL_dalpha += (r3.x - accum_n0) * dn0 + (r3.y - accum_n1) * dn1 + (r3.z - accum_n2) * dn2;
        v[13] = aT * dn0; v[14] = aT * dn1; v[15] = aT * dn2;
        // fold this contributor into the accumulators the NEXT (nearer) contributor sees
        const float oma = 1.f - alpha;
        accum_rec0 = alpha * r3.w + oma * accum_rec0;
        accum_rec1 = alpha * r4.x + oma * accum_rec1;
        accum_rec2 = alpha * r4.y + oma * accum_rec2;
        accum_depth_rec = alpha * c_d + oma * accum_depth_rec;
        accum_alpha_rec = alpha + oma * accum_alpha_rec;
        accum_n0 = alpha * r3.x + oma * accum_n0;
        accum_n1 = alpha * r3.y + oma * accum_n1;
        accum_n2 = alpha * r3.z + oma * accum_n2;

        dL_dalpha *= T;
        dL_dalpha += (-T_final * r1ma) * bg_dot_dpixel;
        const float dL_dG = r2.w * dL_dalpha;
        dL_dz += aT * dL_ddepth;
        if (rho3d <= rho2d) {
            const float dL_dsx = dL_dG * -G_ * sx + dL_dz * r1.z;
            const float dL_dsy = dL_dG * -G_ * sy + dL_dz * r1.w;
            const float rpz = rcp_approx(pz);
            const float dpx = dL_dsx * rpz, dpy = dL_dsy * rpz, dpz = -(dpx * sx + dpy * sy);
            // dL_dk = l x dL_dp ; dL_dl = dL_dp x k
            const float dkx = ly_ * dpz - lz_ * dpy, dky = lz_ * dpx - lx_ * dpz, dkz = lx_ * dpy - ly_ * dpx;
            const float dlx = dpy * kz - dpz * ky, dly = dpz * kx - dpx * kz, dlz = dpx * ky - dpy * kx;
            v[0] = -dkx; v[1] = -dky; v[2] = -dkz;
            v[3] = -dlx; v[4] = -dly; v[5] = -dlz;
            v[6] = pixx * dkx + pixy * dlx + dL_dz * sx;
            v[7] = pixx * dky + pixy * dly + dL_dz * sy;
            v[8] = pixx * dkz + pixy * dlz + dL_dz;
        } else {
            lowpass = true;
            m2x = dL_dG * (-G_ * 2.0f * dx);
            m2y = dL_dG * (-G_ * 2.0f * dy);
            v[8] = dL_dz;
        }
        v[9] = G_ * dL_dalpha;
    }
};

// one pixel per lane: the lane's 16 components leave as four vector reductions (+ one for the low-pass branch)
__device__ __forceinline__ void red_pixel(float* __restrict__ sgrad, const float4* __restrict__ R, const float (&v)[16],
                                          float m2x, float m2y, bool lowpass) {
    float* gp = sgrad + (size_t)__float_as_uint(R[4].z) * SR_GRAD_FLOATS;
    red_add<16>(gp, v);
    if (lowpass) {
        const float w2[16] = {m2x, m2y};
        red_add<2>(gp + SR_G_M2D, w2);
    }
}

template <int G, int MINB = 20>
__global__ void __launch_bounds__(32 * WPC, MINB / WPC)
composite_bwd_kernel(const uint2* __restrict__ ranges, const uint32_t* __restrict__ tile_order, int n_items, int tiles_x,
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
    const int item = blockIdx.x * WPC + (threadIdx.x >> 5);
    if (item >= n_items) return;
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
    px.load(tile_x * SR_TILE + sx0 + lx, tile_y * SR_TILE + sy0 + ly, W, H, bg, final_Ts, n_contrib, dL_dpixels, dL_dothers);

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
            if (contrib) px.pair(S + jj * REC4, pos, v, m2x, m2y, lowpass);
            if constexpr (GS::GL == 1) {
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
    cudaError_t e = cudaMemsetAsync(a.geom + a.gl.sgrad, 0, (size_t)(a.cam.P > 0 ? a.cam.P : 1) * SR_GRAD_FLOATS * 4, a.stream);
    if (e != cudaSuccess) return e;
    ProfileScope ps("composite_bwd", a.stream);
    auto launch = [&](auto kern) {
        kern<<<(a.il.tiles * 8 + comp::WPC - 1) / comp::WPC, 32 * comp::WPC, 0, a.stream>>>(
            (const uint2*)(a.img + a.il.ranges), (const uint32_t*)(a.img + a.il.tile_order), a.il.tiles * 8, a.il.tiles_x,
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
            else launch(composite_bwd_kernel<32, 28>);
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
