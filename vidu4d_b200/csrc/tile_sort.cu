// tile_sort.cu -- tile-local binning + sort (SR_FLAG_LOCAL_SORT), the fast path for the usual case where no
// 16x16 tile holds more than SR_LOCAL_SORT_CAP instances.
//
// Produces exactly what the reference's scan + duplicateWithKeys + global stable radix sort + identifyTileRanges
// produce (rasterizer_impl.cu:278-319): the instance list ordered by (tile, depth bits), ties in emission order.
// Emission order within a tile is ascending surfel id and a surfel touches a tile at most once, so that order is
// the unique ascending order of the 64-bit word (depth_bits << 32 | surfel id) inside each tile -- any sort gives
// it, stable or not.  So:
//   preprocess_fwd   counts instances per tile (one atomic per touched tile)
//   tile_scan        1 CTA: exclusive scan of the counts = tile ranges, R, overflow / cap checks, LPT launch order
//   emit_local       scatters (depth_bits, id) words into each tile's segment through an atomic cursor (any order)
//   tile_sort_gather 1 CTA per tile: bitonic sort of the segment in shared memory, then writes the sorted surfel
//                    list, the (tile | depth) keys and the 80-byte instance record stream for the composite kernels
// 4 launches and ~2 passes over the instances instead of histogram + plan + 6 onesweep passes + gather.
#include "common.cuh"

namespace {

// auxiliary.h:64-74, same arithmetic as preprocess.cu
__device__ __forceinline__ void get_rect(float px, float py, int max_radius, int gx, int gy, uint2& rmin, uint2& rmax) {
    const float r = (float)max_radius;
    int x0 = (int)__fmul_rn(__fadd_rn(px, -r), 0.0625f), y0 = (int)__fmul_rn(__fadd_rn(py, -r), 0.0625f);
    int x1 = (int)__fmul_rn(__fadd_rn(__fadd_rn(__fadd_rn(px, r), 16.f), -1.f), 0.0625f);
    int y1 = (int)__fmul_rn(__fadd_rn(__fadd_rn(__fadd_rn(py, r), 16.f), -1.f), 0.0625f);
    rmin.x = min((unsigned)gx, (unsigned)max(0, x0));
    rmin.y = min((unsigned)gy, (unsigned)max(0, y0));
    rmax.x = min((unsigned)gx, (unsigned)max(0, x1));
    rmax.y = min((unsigned)gy, (unsigned)max(0, y1));
}

__global__ void __launch_bounds__(1024)
tile_scan_kernel(const uint32_t* __restrict__ tile_count, int tiles, uint2* __restrict__ ranges,
                 uint32_t* __restrict__ order, uint32_t* __restrict__ num_rendered, long long capacity) {
    __shared__ uint32_t wtot[32];
    __shared__ uint32_t carry_s, maxc_s;
    __shared__ uint32_t cnt[33], start[33];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) { carry_s = 0; maxc_s = 0; }
    if (threadIdx.x < 33) cnt[threadIdx.x] = 0;
    __syncthreads();
    for (int base = 0; base < tiles; base += 1024) {
        const int i = base + threadIdx.x;
        const uint32_t v = i < tiles ? tile_count[i] : 0u;
        uint32_t inc = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { uint32_t n = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o) inc += n; }
        if (lane == 31) wtot[warp] = inc;
        __syncthreads();
        if (threadIdx.x < 32) {
            uint32_t w = wtot[threadIdx.x], winc = w;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) { uint32_t n = __shfl_up_sync(0xffffffffu, winc, o); if (threadIdx.x >= o) winc += n; }
            wtot[threadIdx.x] = winc - w;
        }
        __syncthreads();
        const uint32_t excl = carry_s + wtot[warp] + inc - v;
        if (i < tiles) {
            ranges[i] = v ? make_uint2(excl, excl + v) : make_uint2(0u, 0u);   // empty tiles stay (0,0) like the reference's memset
            atomicAdd(&cnt[v ? 32 - __clz(v) : 0], 1u);
            if (v > SR_LOCAL_SORT_CAP) atomicMax(&maxc_s, v);
        }
        __syncthreads();
        if (threadIdx.x == 1023) carry_s = excl + v;
        __syncthreads();
    }
    __shared__ uint32_t bad_s;
    if (threadIdx.x == 0) {
        const uint32_t R = carry_s;
        num_rendered[0] = R;
        uint32_t st = 0;
        if ((long long)R > capacity) st |= SR_STATUS_OVERFLOW;
        if (maxc_s) st |= SR_STATUS_SORT_CAP;
        if (st) atomicOr(num_rendered + 1, st);
        bad_s = st;
        uint32_t acc = 0;
        for (int b = 32; b >= 0; b--) { start[b] = acc; acc += cnt[b]; }
    }
    __syncthreads();
    if (bad_s) {
        // frame abandoned: the composite kernels must see empty tiles (ranges may point past the capacity)
        for (int t = threadIdx.x; t < tiles; t += 1024) ranges[t] = make_uint2(0u, 0u);
    }
    // longest-list-first launch order for the per-tile kernels (same bucketing as sort.cu:tile_order_kernel)
    for (int t = threadIdx.x; t < tiles; t += 1024) {
        const uint32_t v = tile_count[t];
        order[atomicAdd(&start[v ? 32 - __clz(v) : 0], 1u)] = (uint32_t)t;
    }
}

__global__ void __launch_bounds__(256)
emit_local_kernel(const CamParams c, const float4* __restrict__ srec, const float* __restrict__ depths,
                  const int* __restrict__ radii, const uint32_t* __restrict__ tiles_touched,
                  const uint2* __restrict__ ranges, uint32_t* __restrict__ cursor, uint64_t* __restrict__ keys,
                  const uint32_t* __restrict__ num_rendered) {
    if (num_rendered[1] & (SR_STATUS_OVERFLOW | SR_STATUS_SORT_CAP)) return;   // frame abandoned
    const int idx = blockIdx.x * 256 + threadIdx.x;
    if (idx >= c.P) return;
    if (tiles_touched[idx] == 0) return;
    const float4 r2 = __ldg(srec + (size_t)idx * 5 + 2);   // (Tw.z, xy.x, xy.y, opacity)
    uint2 rmin, rmax;
    get_rect(r2.y, r2.z, radii[idx], c.tiles_x, c.tiles_y, rmin, rmax);
    const uint64_t word = ((uint64_t)__float_as_uint(depths[idx]) << 32) | (uint32_t)idx;
    for (uint32_t y = rmin.y; y < rmax.y; y++)
        for (uint32_t x = rmin.x; x < rmax.x; x++) {
            const uint32_t t = y * (uint32_t)c.tiles_x + x;
            keys[ranges[t].x + atomicAdd(&cursor[t], 1u)] = word;
        }
}

__global__ void __launch_bounds__(256)
tile_sort_gather_kernel(const uint2* __restrict__ ranges, const uint32_t* __restrict__ order, int tiles_x,
                        const uint64_t* __restrict__ seg_keys, uint64_t* __restrict__ keys_sorted,
                        uint32_t* __restrict__ point_list, const float4* __restrict__ srec, float4* __restrict__ irec,
                        const uint32_t* __restrict__ num_rendered) {
    extern __shared__ uint64_t sk[];
    if (num_rendered[1] & (SR_STATUS_OVERFLOW | SR_STATUS_SORT_CAP)) return;
    const uint32_t tile = order[blockIdx.x];
    const uint2 range = ranges[tile];
    const int n = (int)(range.y - range.x);
    if (n == 0) return;
    int P2 = 2;
    while (P2 < n) P2 <<= 1;
    const int tid = threadIdx.x;
    for (int i = tid; i < P2; i += 256) sk[i] = i < n ? seg_keys[range.x + i] : ~0ull;
    __syncthreads();
    for (int k = 2; k <= P2; k <<= 1) {
        for (int j = k >> 1; j > 0; j >>= 1) {
            for (int i = tid; i < (P2 >> 1); i += 256) {
                const int a = ((i & ~(j - 1)) << 1) | (i & (j - 1));     // j is a power of two
                const int b = a | j;
                const uint64_t x = sk[a], y = sk[b];
                const bool up = (a & k) == 0;
                if ((x > y) == up) { sk[a] = y; sk[b] = x; }
            }
            __syncthreads();
        }
    }
    const int tx = (int)(tile % (uint32_t)tiles_x) * SR_TILE, ty = (int)(tile / (uint32_t)tiles_x) * SR_TILE;
    for (int i = tid; i < n; i += 256) {
        const uint64_t w = sk[i];
        const uint32_t id = (uint32_t)w;
        const size_t o = (size_t)range.x + i;
        point_list[o] = id;
        keys_sorted[o] = ((uint64_t)tile << 32) | (w >> 32);
        const float4* s = srec + (size_t)id * 5;
        const float4 a0 = __ldg(s), a1 = __ldg(s + 1), a2 = __ldg(s + 2), a3 = __ldg(s + 3);
        float4 a4 = __ldg(s + 4);
        const uint32_t bx = __float_as_uint(a4.z), by = __float_as_uint(a4.w);
        const int x0 = max((int)(bx & 0xffffu) - tx, 0), x1 = min((int)(bx >> 16) - tx, SR_TILE - 1);
        const int y0 = max((int)(by & 0xffffu) - ty, 0), y1 = min((int)(by >> 16) - ty, SR_TILE - 1);
        uint32_t cull = 0;
        if (x0 <= x1 && y0 <= y1) {
            // rho_cut, see sort.cu:ranges_gather_kernel
            const float c2 = 2.0f * logf(255.0f * a2.w) + 1e-4f;
            const uint32_t q = (uint32_t)min(16383.0f, fmaxf(0.0f, ceilf(c2 * 1024.0f)));
            cull = (uint32_t)x0 | ((uint32_t)x1 << 4) | ((uint32_t)y0 << 8) | ((uint32_t)y1 << 12) | (1u << 16) | (q << 17);
        }
        a4.z = __uint_as_float(id);
        a4.w = __uint_as_float(cull);
        float4* d = irec + o * 5;
        d[0] = a0; d[1] = a1; d[2] = a2; d[3] = a3; d[4] = a4;
    }
}

}  // namespace

cudaError_t launch_tile_scan_emit(const FwdArgs& a) {
    uint2* ranges = (uint2*)(a.img + a.il.ranges);
    {
        ProfileScope ps("tile_scan", a.stream);
        tile_scan_kernel<<<1, 1024, 0, a.stream>>>((const uint32_t*)(a.img + a.il.tile_count), a.il.tiles, ranges,
                                                    (uint32_t*)(a.img + a.il.tile_order), a.num_rendered_dev,
                                                    (long long)a.bl.capacity);
    }
    ProfileScope ps("emit_local", a.stream);
    emit_local_kernel<<<a.gl.nblocks, 256, 0, a.stream>>>(
        a.cam, (const float4*)(a.geom + a.gl.surfel_rec), (const float*)(a.geom + a.gl.depths), a.radii,
        (const uint32_t*)(a.geom + a.gl.tiles_touched), ranges, (uint32_t*)(a.img + a.il.tile_cursor),
        (uint64_t*)(a.bin + a.bl.keys[0]), a.num_rendered_dev);
    sr_count_launch(2);
    return cudaGetLastError();
}

cudaError_t launch_tile_sort_gather(const FwdArgs& a) {
    const size_t smem = (size_t)SR_LOCAL_SORT_CAP * sizeof(uint64_t);
    static bool attr_set = false;
    if (!attr_set) {
        cudaError_t e = cudaFuncSetAttribute(tile_sort_gather_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        attr_set = true;
    }
    // the sorted arrays live in the "pong" halves: sort_ctl[SORTED_SEL] = 1 (set by the caller's memset + this store)
    ProfileScope ps("tile_sort_gather", a.stream);
    tile_sort_gather_kernel<<<a.il.tiles, 256, smem, a.stream>>>(
        (const uint2*)(a.img + a.il.ranges), (const uint32_t*)(a.img + a.il.tile_order), a.il.tiles_x,
        (const uint64_t*)(a.bin + a.bl.keys[0]), (uint64_t*)(a.bin + a.bl.keys[1]), (uint32_t*)(a.bin + a.bl.values[1]),
        (const float4*)(a.geom + a.gl.surfel_rec), (float4*)(a.bin + a.bl.inst_rec), a.num_rendered_dev);
    sr_count_launch();
    return cudaGetLastError();
}
