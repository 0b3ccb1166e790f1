// api.cu -- the C ABI declared in include/surfel_raster.h: argument checking, buffer layout,
// launch orchestration.  Replaces the host side of the reference:
//   Rasterizer::forward / backward / markVisible   RAST/cuda_rasterizer/rasterizer_impl.cu:141-153,198-342,346-448
//   GeometryState/ImageState/BinningState::fromChunk, required<T>   rasterizer_impl.cu:155-194, rasterizer_impl.h:66-72
// No torch types; no host<->device synchronisation on the forward/backward path (debug mode excepted).
#include <atomic>
#include <map>
#include <mutex>
#include <string>
#include <vector>
#include <cstdarg>
#include <cstdio>
#include <cstring>

#include "common.cuh"

namespace {
thread_local char g_err[512] = "";
std::atomic<uint64_t> g_launches{0};

int fail(int code, const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    return code;
}
int cuda_fail(cudaError_t e, const char* where) {
    return fail(SR_ECUDA, "CUDA error in %s: %s", where, cudaGetErrorString(e));
}
#define CK(call, where)                                      \
    do {                                                     \
        cudaError_t e_ = (call);                             \
        if (e_ != cudaSuccess) return cuda_fail(e_, where);  \
    } while (0)
#define DBG(where)                                           \
    do {                                                     \
        if (debug) CK(cudaStreamSynchronize(stream), where); \
    } while (0)

CamParams make_cam(const sr_frame* f, const float* vm, const float* campos, const float* bg) {
    CamParams c{};
    c.vm = vm; c.campos = campos; c.bg = bg;
    c.W = f->width; c.H = f->height;
    c.tiles_x = (f->width + SR_TILE - 1) / SR_TILE;
    c.tiles_y = (f->height + SR_TILE - 1) / SR_TILE;
    c.P = f->P; c.D = f->sh_degree; c.M = f->sh_coeffs;
    // rasterizer_impl.cu:223-224
    c.focal_y = f->height / (2.0f * f->tan_fovy);
    c.focal_x = f->width / (2.0f * f->tan_fovx);
    // forward.cu:208  {focal_x, focal_y, float(W)/2.0, float(H)/2.0}
    c.cx = (float)((float)f->width / 2.0);
    c.cy = (float)((float)f->height / 2.0);
    // backward.cu:570 / 683-684
    c.bcx = c.focal_x * f->tan_fovx;
    c.bcy = c.focal_y * f->tan_fovy;
    return c;
}

int check_frame(const sr_frame* f) {
    if (!f) return fail(SR_EINVAL, "frame descriptor is NULL");
    if (f->P < 0) return fail(SR_EINVAL, "P must be >= 0");
    if (f->width <= 0 || f->height <= 0) return fail(SR_EINVAL, "image size must be positive");
    if (f->width > 65535 || f->height > 65535) return fail(SR_EINVAL, "image side must be < 65536");
    if (f->sh_degree < 0 || f->sh_degree > 3) return fail(SR_EINVAL, "sh_degree must be 0..3");
    if (f->sh_coeffs < 0) return fail(SR_EINVAL, "sh_coeffs must be >= 0");
    if (f->sh_coeffs > 0 && (f->sh_degree + 1) * (f->sh_degree + 1) > f->sh_coeffs)
        return fail(SR_EINVAL, "sh_degree %d needs %d coefficients, shs has %d", f->sh_degree,
                    (f->sh_degree + 1) * (f->sh_degree + 1), f->sh_coeffs);
    return 0;
}
}  // namespace

void sr_count_launch(int n) { g_launches.fetch_add((uint64_t)n, std::memory_order_relaxed); }

cudaError_t sr_memset_frames(void* p, size_t pitch, size_t bytes, int frames, cudaStream_t s) {
    if (frames == 1) return cudaMemsetAsync(p, 0, bytes, s);
    return cudaMemset2DAsync(p, pitch, 0, bytes, (size_t)frames, s);
}

namespace {
std::atomic<bool> g_profiling{false};
std::mutex g_prof_mu;
struct ProfRec { const char* name; cudaEvent_t e0, e1; };
std::vector<ProfRec> g_prof;
std::string g_prof_json;
}  // namespace
bool sr_profiling_on() { return g_profiling.load(std::memory_order_relaxed); }
void sr_profile_push(const char* name, cudaEvent_t e0, cudaEvent_t e1) {
    std::lock_guard<std::mutex> lk(g_prof_mu);
    g_prof.push_back({name, e0, e1});
}

extern "C" {

int sr_abi_version(void) { return SR_ABI_VERSION; }
const char* sr_last_error(void) { return g_err; }
uint64_t sr_launch_count(void) { return g_launches.load(); }

void sr_set_profiling(int on) { g_profiling.store(on != 0); }

// Synchronises the device, then returns {"kernel": {"ms": total, "count": n}, ...} for everything recorded
// since the previous call.  The returned pointer stays valid until the next call.
const char* sr_get_profile(void) {
    cudaDeviceSynchronize();
    std::lock_guard<std::mutex> lk(g_prof_mu);
    std::map<std::string, std::pair<double, int>> acc;
    for (auto& r : g_prof) {
        float ms = 0.f;
        if (cudaEventElapsedTime(&ms, r.e0, r.e1) == cudaSuccess) { acc[r.name].first += ms; acc[r.name].second += 1; }
        cudaEventDestroy(r.e0); cudaEventDestroy(r.e1);
    }
    g_prof.clear();
    g_prof_json = "{";
    bool first = true;
    for (auto& kv : acc) {
        char buf[256];
        snprintf(buf, sizeof(buf), "%s\"%s\": {\"ms\": %.6f, \"count\": %d}", first ? "" : ", ", kv.first.c_str(), kv.second.first, kv.second.second);
        g_prof_json += buf; first = false;
    }
    g_prof_json += "}";
    return g_prof_json.c_str();
}

size_t sr_geom_bytes(int32_t P) { return geom_layout(P).total; }
size_t sr_image_bytes(int32_t width, int32_t height) { return image_layout(width, height).total; }
size_t sr_binning_bytes(int64_t capacity, int32_t width, int32_t height) {
    return bin_layout(capacity, image_layout(width > 0 ? width : 0, height > 0 ? height : 0).tiles).total;
}

int sr_debug_view(int32_t P, int32_t width, int32_t height, int64_t capacity, sr_debug_layout* out) {
    if (!out) return fail(SR_EINVAL, "out is NULL");
    const GeomLayout g = geom_layout(P);
    const ImageLayout i = image_layout(width, height);
    const BinLayout b = bin_layout(capacity, i.tiles);
    out->surfel_rec = g.surfel_rec; out->depths = g.depths; out->tiles_touched = g.tiles_touched;
    out->point_offsets = g.point_offsets; out->clamped = g.clamped;
    out->keys[0] = b.keys[0]; out->keys[1] = b.keys[1]; out->values[0] = b.values[0]; out->values[1] = b.values[1];
    out->sort_ctl = b.sort_ctl; out->inst_rec = b.inst_rec; out->contrib = b.contrib;
    out->final_T = i.final_T; out->n_contrib = i.n_contrib; out->ranges = i.ranges; out->sub_last = i.tile_last;
    return 0;
}

int sr_forward_batch(const sr_frame* f, const sr_batch* b, const float* background, const float* means3D, const float* shs,
                     const float* colors_precomp, const float* opacities, const float* scales, const float* rotations,
                     const float* viewmatrix, const float* projmatrix, const float* campos, float* out_color,
                     float* out_others, int32_t* radii, void* geom_buffer, void* binning_buffer, void* image_buffer,
                     int64_t capacity, uint32_t* num_rendered_dev, uint32_t* num_rendered_host, void* stream_) {
    (void)projmatrix;   // only feeds dead code in the reference (auxiliary.h:170-172)
    if (int rc = check_frame(f)) return rc;
    if (!b || b->frames < 1 || b->frames > 65535) return fail(SR_EINVAL, "batch descriptor: frames must be in [1, 65535]");
    cudaStream_t stream = (cudaStream_t)stream_;
    const bool debug = f->debug != 0;
    const int P = f->P, M = b->frames;
    const size_t N = (size_t)f->width * f->height;
    if (!out_color || !out_others || !num_rendered_dev) return fail(SR_EINVAL, "output pointer is NULL");
    if (!background || !viewmatrix || !campos) return fail(SR_EINVAL, "camera / background pointer is NULL");
    CK(cudaMemsetAsync(num_rendered_dev, 0, (size_t)M * 2 * sizeof(uint32_t), stream), "memset(num_rendered)");
    if (P == 0) {
        // rasterize_points.cu:106 -- P == 0 short-circuits to zero-filled outputs
        CK(cudaMemsetAsync(out_color, 0, (size_t)M * 3 * N * sizeof(float), stream), "memset(out_color)");
        CK(cudaMemsetAsync(out_others, 0, (size_t)M * 8 * N * sizeof(float), stream), "memset(out_others)");
        if (num_rendered_host)
            CK(cudaMemcpyAsync(num_rendered_host, num_rendered_dev, (size_t)M * 2 * sizeof(uint32_t), cudaMemcpyDeviceToHost, stream), "copy(num_rendered)");
        return 0;
    }
    if (!means3D || !opacities || !scales || !rotations || !radii) return fail(SR_EINVAL, "surfel attribute pointer is NULL");
    if ((shs == nullptr) == (colors_precomp == nullptr))
        return fail(SR_EINVAL, "provide exactly one of shs / colors_precomp");
    if (shs && f->sh_coeffs == 0) return fail(SR_EINVAL, "shs given but sh_coeffs == 0");
    if (!geom_buffer || !binning_buffer || !image_buffer) return fail(SR_EINVAL, "scratch buffer is NULL");
    if (capacity <= 0 || capacity >= (1ll << 30)) return fail(SR_EINVAL, "capacity must be in (0, 2^30)");
    // 128-bit / 64-bit vector loads: rotations as float4, scales as float2, SH rows as float4 when 3M % 4 == 0
    if (((uintptr_t)rotations & 15) || ((uintptr_t)scales & 7) || (shs && ((uintptr_t)shs & 15)) ||
        ((b->rotations * 4) & 15) || ((b->scales * 4) & 7) || ((b->shs * 4) & 15))
        return fail(SR_EINVAL, "rotations / shs must be 16-byte aligned and scales 8-byte aligned (per frame)");

    FwdArgs a{};
    a.cam = make_cam(f, viewmatrix, campos, background);
    if (!shs) a.cam.M = 0;
    a.means3D = means3D; a.shs = shs; a.colors_precomp = colors_precomp; a.opacities = opacities;
    a.scales = scales; a.rotations = rotations;
    a.out_color = out_color; a.out_others = out_others; a.radii = radii;
    a.geom = (char*)geom_buffer; a.bin = (char*)binning_buffer; a.img = (char*)image_buffer;
    a.gl = geom_layout(P); a.il = image_layout(f->width, f->height); a.bl = bin_layout(capacity, a.il.tiles);
    a.num_rendered_dev = num_rendered_dev;
    a.prefiltered = f->prefiltered;
    a.key_bits = 32 + (int)sr_higher_msb((uint32_t)a.il.tiles);
    a.stream = stream; a.debug = debug;
    FrameStrides& fs = a.fs;
    fs.frames = M;
    fs.geom = (long long)a.gl.total; fs.bin = (long long)a.bl.total; fs.img = (long long)a.il.total; fs.nr = 2 * sizeof(uint32_t);
    fs.means3D = b->means3D * 4; fs.shs = b->shs * 4; fs.colors = b->colors_precomp * 4; fs.opac = b->opacities * 4;
    fs.scales = b->scales * 4; fs.rots = b->rotations * 4; fs.vm = 16 * sizeof(float); fs.campos = 3 * sizeof(float);
    fs.out_color = (long long)(3 * N * sizeof(float)); fs.out_others = (long long)(8 * N * sizeof(float));
    fs.radii = (long long)P * sizeof(int32_t);

    // per frame: tile state (ranges .. end of the image buffer); sort control words + digit histograms + look-back status
    CK(sr_memset_frames(a.img + a.il.ranges, a.il.total, a.il.total - a.il.ranges, M, stream), "memset(tile state)");
    CK(sr_memset_frames(a.bin + a.bl.sort_ctl, a.bl.total, a.bl.total - a.bl.sort_ctl, M, stream), "memset(sort state)");
    CK(launch_preprocess_fwd(a), "preprocess_fwd"); DBG("preprocess_fwd");
    CK(launch_scan_emit(a), "scan/emit_keys"); DBG("scan/emit_keys");
    CK(launch_sort(a), "sort"); DBG("sort");
    CK(launch_ranges_gather(a), "ranges_gather"); DBG("ranges_gather");
    CK(launch_tile_order(a), "tile_order"); DBG("tile_order");
    if (sr_composite_tile_mode(false)) { CK(launch_composite_tile_fwd(a), "composite_tile_fwd"); }
    else { CK(launch_composite_fwd(a), "composite_fwd"); }
    DBG("composite_fwd");
    if (num_rendered_host)
        CK(cudaMemcpyAsync(num_rendered_host, num_rendered_dev, (size_t)M * 2 * sizeof(uint32_t), cudaMemcpyDeviceToHost, stream), "copy(num_rendered)");
    return 0;
}

int sr_forward(const sr_frame* f, const float* background, const float* means3D, const float* shs,
               const float* colors_precomp, const float* opacities, const float* scales, const float* rotations,
               const float* viewmatrix, const float* projmatrix, const float* campos, float* out_color,
               float* out_others, int32_t* radii, void* geom_buffer, void* binning_buffer, void* image_buffer,
               int64_t capacity, uint32_t* num_rendered_dev, uint32_t* num_rendered_host, void* stream_) {
    const sr_batch one = {1, 0u, 0, 0, 0, 0, 0, 0};
    return sr_forward_batch(f, &one, background, means3D, shs, colors_precomp, opacities, scales, rotations, viewmatrix,
                            projmatrix, campos, out_color, out_others, radii, geom_buffer, binning_buffer, image_buffer,
                            capacity, num_rendered_dev, num_rendered_host, stream_);
}

int sr_backward_batch(const sr_frame* f, const sr_batch* b, const float* background, const float* means3D, const float* shs,
                      const float* colors_precomp, const float* scales, const float* rotations, const float* viewmatrix,
                      const float* projmatrix, const float* campos, const int32_t* radii, const float* dL_dout_color,
                      const float* dL_dout_others, const float* grad_scale, void* geom_buffer, void* binning_buffer,
                      void* image_buffer, int64_t capacity, float* dL_dmeans2D, float* dL_dcolors, float* dL_dopacity,
                      float* dL_dmeans3D, float* dL_dtransMat, float* dL_dsh, float* dL_dscales, float* dL_drotations,
                      void* stream_) {
    (void)projmatrix;
    if (int rc = check_frame(f)) return rc;
    if (!b || b->frames < 1 || b->frames > 65535) return fail(SR_EINVAL, "batch descriptor: frames must be in [1, 65535]");
    cudaStream_t stream = (cudaStream_t)stream_;
    const bool debug = f->debug != 0;
    if (f->P == 0) return 0;   // rasterize_points.cu:204
    if (!background || !means3D || !scales || !rotations || !viewmatrix || !campos || !radii || !dL_dout_color ||
        !dL_dout_others || !geom_buffer || !binning_buffer || !image_buffer)
        return fail(SR_EINVAL, "input pointer is NULL");
    if (!dL_dmeans2D || !dL_dcolors || !dL_dopacity || !dL_dmeans3D || !dL_dscales || !dL_drotations)
        return fail(SR_EINVAL, "gradient output pointer is NULL");
    if (!dL_dtransMat && b->frames == 1 && !(b->flags & SR_BATCH_SUM_SHARED)) return fail(SR_EINVAL, "dL_dtransMat is NULL");
    if (shs && !dL_dsh) return fail(SR_EINVAL, "dL_dsh is NULL");
    if (capacity <= 0) return fail(SR_EINVAL, "capacity must be positive");
    if (((uintptr_t)rotations & 15) || ((uintptr_t)scales & 7) || (shs && ((uintptr_t)shs & 15)) ||
        ((uintptr_t)dL_drotations & 15) || ((uintptr_t)dL_dscales & 7) || (dL_dsh && ((uintptr_t)dL_dsh & 15)) ||
        ((b->rotations * 4) & 15) || ((b->scales * 4) & 7) || ((b->shs * 4) & 15))
        return fail(SR_EINVAL, "rotations / shs (and their gradients) must be 16-byte aligned, scales 8-byte aligned");

    const int P = f->P, M = b->frames;
    const size_t N = (size_t)f->width * f->height;
    BwdArgs a{};
    a.cam = make_cam(f, viewmatrix, campos, background);
    if (!shs) a.cam.M = 0;
    a.grad_scale = grad_scale;
    a.means3D = means3D; a.shs = shs; a.colors_precomp = colors_precomp; a.scales = scales; a.rotations = rotations;
    a.radii = radii; a.dL_dcolor = dL_dout_color; a.dL_dothers = dL_dout_others;
    a.geom = (char*)geom_buffer; a.bin = (char*)binning_buffer; a.img = (char*)image_buffer;
    a.gl = geom_layout(P); a.il = image_layout(f->width, f->height); a.bl = bin_layout(capacity, a.il.tiles);
    a.dL_dmeans2D = dL_dmeans2D; a.dL_dcolors = dL_dcolors; a.dL_dopacity = dL_dopacity; a.dL_dmeans3D = dL_dmeans3D;
    a.dL_dtransMat = dL_dtransMat; a.dL_dsh = dL_dsh; a.dL_dscales = dL_dscales; a.dL_drotations = dL_drotations;
    a.stream = stream; a.debug = debug;
    FrameStrides& fs = a.fs;
    fs.frames = M;
    fs.geom = (long long)a.gl.total; fs.bin = (long long)a.bl.total; fs.img = (long long)a.il.total; fs.nr = 2 * sizeof(uint32_t);
    fs.means3D = b->means3D * 4; fs.shs = b->shs * 4; fs.colors = b->colors_precomp * 4; fs.opac = b->opacities * 4;
    fs.scales = b->scales * 4; fs.rots = b->rotations * 4; fs.vm = 16 * sizeof(float); fs.campos = 3 * sizeof(float);
    fs.radii = (long long)P * sizeof(int32_t);
    fs.dcolor = (long long)(3 * N * sizeof(float)); fs.dothers = (long long)(8 * N * sizeof(float));
    // gradients are written per frame, (M, P, .) -- or, with SR_BATCH_SUM_SHARED, once as the sum over frames, (P, .), for
    // the inputs the frames share (surfel_bwd.cu accumulates them inside the kernel; stride 0 selects that)
    const long long Pl = P;
    const bool sum = (b->flags & SR_BATCH_SUM_SHARED) != 0 && M > 1;
    fs.g_m2d = Pl * 3 * 4; fs.g_tm = Pl * 9 * 4;
    fs.g_col = (sum && b->colors_precomp == 0) ? 0 : Pl * 3 * 4;
    fs.g_opac = (sum && b->opacities == 0) ? 0 : Pl * 4;
    fs.g_m3d = (sum && b->means3D == 0) ? 0 : Pl * 3 * 4;
    fs.g_sh = (sum && b->shs == 0) ? 0 : Pl * a.cam.M * 3 * 4;
    fs.g_scales = (sum && b->scales == 0) ? 0 : Pl * 2 * 4;
    fs.g_rots = (sum && b->rotations == 0) ? 0 : Pl * 4 * 4;
    if (sr_composite_tile_mode(true)) { CK(launch_composite_tile_bwd(a), "composite_tile_bwd"); }
    else { CK(launch_composite_bwd(a), "composite_bwd"); }
    DBG("composite_bwd");
    CK(launch_surfel_bwd(a), "surfel_bwd"); DBG("surfel_bwd");
    return 0;
}

int sr_backward(const sr_frame* f, const float* background, const float* means3D, const float* shs,
                const float* colors_precomp, const float* scales, const float* rotations, const float* viewmatrix,
                const float* projmatrix, const float* campos, const int32_t* radii, const float* dL_dout_color,
                const float* dL_dout_others, void* geom_buffer, void* binning_buffer, void* image_buffer,
                int64_t capacity, float* dL_dmeans2D, float* dL_dcolors, float* dL_dopacity, float* dL_dmeans3D,
                float* dL_dtransMat, float* dL_dsh, float* dL_dscales, float* dL_drotations, void* stream_) {
    const sr_batch one = {1, 0u, 0, 0, 0, 0, 0, 0};
    return sr_backward_batch(f, &one, background, means3D, shs, colors_precomp, scales, rotations, viewmatrix, projmatrix,
                             campos, radii, dL_dout_color, dL_dout_others, nullptr, geom_buffer, binning_buffer, image_buffer,
                             capacity, dL_dmeans2D, dL_dcolors, dL_dopacity, dL_dmeans3D, dL_dtransMat, dL_dsh, dL_dscales,
                             dL_drotations, stream_);
}

int sr_mark_visible(int32_t P, const float* means3D, const float* viewmatrix, const float* projmatrix,
                    uint8_t* present, void* stream_) {
    (void)projmatrix;
    if (P < 0) return fail(SR_EINVAL, "P must be >= 0");
    if (P == 0) return 0;
    if (!means3D || !viewmatrix || !present) return fail(SR_EINVAL, "pointer is NULL");
    CK(launch_mark_visible(P, means3D, viewmatrix, present, (cudaStream_t)stream_), "mark_visible");
    return 0;
}

}  // extern "C"
