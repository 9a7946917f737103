// C-ABI glue for include/svslam.h: context, frame sets and the image-stage entry points
// (a0 half-resolution resize, a1 GFTT, a2/a3 pyramidal LK).  Host pointers in, host pointers out;
// staging through pinned buffers owned by the context; one CUDA stream per context.
#include "svs_internal.h"
#include <algorithm>
#include <atomic>
#include <mutex>
#include <set>
#include <utility>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <new>

int svs_i_gftt_overflow(svs_ctx *c, int n_img, int *flag_host);

static std::string g_create_err;
static std::mutex g_mutex;                       // guards g_create_err and g_smem_opted
static std::set<std::pair<int, const void *>> g_smem_opted;

cudaError_t svs_i_opt_in_smem(svs_ctx *c, const void *func)
{
    std::lock_guard<std::mutex> lk(g_mutex);
    if (g_smem_opted.count({c->device, func})) return cudaSuccess;
    int max_optin = 0;
    cudaError_t e = cudaDeviceGetAttribute(&max_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, c->device);
    if (e != cudaSuccess) return e;
    cudaFuncAttributes fa;
    if ((e = cudaFuncGetAttributes(&fa, func)) != cudaSuccess) return e;
    // the opt-in limit covers static + dynamic shared memory of the kernel
    e = cudaFuncSetAttribute(func, cudaFuncAttributeMaxDynamicSharedMemorySize, max_optin - (int)fa.sharedSizeBytes);
    if (e == cudaSuccess) g_smem_opted.insert({c->device, func});
    return e;
}
static std::atomic<int> g_live_ctx{0};
long long svs_i_regrowths = 0;     // diagnostic counter (not atomic: an approximate count is enough)
long long svs_buffer_regrowths(void) { return svs_i_regrowths; }

int svs_i_zc_grid(const svs_ctx *c)
{
    if (c->zc_ctas > 0) return c->zc_ctas;
    int live = std::max(1, g_live_ctx.load());
    return std::max(2, 32 / live);
}

void svs_i_prof_begin(svs_ctx *c, int kid)
{
    if (!c->prof) return;
    SvsPendingEv p;
    p.kid = kid;
    cudaEvent_t *ev[2] = {&p.a, &p.b};
    for (int i = 0; i < 2; i++) {
        if (!c->prof_free.empty()) { *ev[i] = c->prof_free.back(); c->prof_free.pop_back(); }
        else cudaEventCreate(ev[i]);
    }
    cudaEventRecord(p.a, c->stream);
    c->prof_pending.push_back(p);
}
void svs_i_prof_end(svs_ctx *c)
{
    if (!c->prof || c->prof_pending.empty()) return;
    cudaEventRecord(c->prof_pending.back().b, c->stream);
}
static void prof_harvest(svs_ctx *c)
{
    if (c->prof_pending.empty()) return;
    cudaStreamSynchronize(c->stream);
    for (SvsPendingEv &p : c->prof_pending) {
        float ms = 0.f;
        if (cudaEventElapsedTime(&ms, p.a, p.b) == cudaSuccess) { c->prof_ms[p.kid] += ms; c->prof_n[p.kid]++; }
        c->prof_free.push_back(p.a); c->prof_free.push_back(p.b);
    }
    c->prof_pending.clear();
}

extern "C" {

static int frameset_fill(svs_ctx *c, svs_frameset *fs, const PyrDesc &Lc, const PyrDesc *Rc, const uint8_t *dl, const uint8_t *dr,
                         size_t rs, size_t is, const uint8_t *const *pl, const uint8_t *const *pr, int aligned4, int rows_decimated,
                         int zero_copy, int n_img, const int32_t *ids_dev);
static int frameset_begin_push(svs_ctx *c, svs_frameset *fs);

int svs_version(void) { return 100; }

int svs_kernel_timing_enable(svs_ctx *c, int on)
{
    if (!c) return SVS_ERR_ARG;
    prof_harvest(c);
    c->prof = on != 0;
    return SVS_OK;
}
int svs_kernel_timing_reset(svs_ctx *c)
{
    if (!c) return SVS_ERR_ARG;
    prof_harvest(c);
    for (int i = 0; i < KID_COUNT; i++) { c->prof_ms[i] = 0; c->prof_n[i] = 0; }
    return SVS_OK;
}
int svs_kernel_timing_get(svs_ctx *c, double *ms /* n */, long long *count /* n */, int n)
{
    if (!c) return SVS_ERR_ARG;
    prof_harvest(c);
    for (int i = 0; i < n && i < KID_COUNT; i++) { if (ms) ms[i] = c->prof_ms[i]; if (count) count[i] = c->prof_n[i]; }
    return KID_COUNT;
}
const char *svs_kernel_name(int kid)
{
    static const char *names[KID_COUNT] = {"k_half_nearest", "k_copy2d", "k_pyr_down", "k_mask_boxes", "k_corner_response",
                                           "k_corner_select", "k_corner_greedy", "k_lk_track", "k_triangulate", "k_pose_only_lm",
                                           "k_ba_window", "k_bm_prefilter", "k_bm_sad", "k_backproject", "k_bgr2gray", "misc", "k_trk_state", "k_ba_build"};
    return (kid >= 0 && kid < KID_COUNT) ? names[kid] : "";
}
const char *svs_create_error(void) { return g_create_err.c_str(); }

svs_ctx *svs_create(int device)
{
    std::lock_guard<std::mutex> create_lock(g_mutex);
    // (CUDA_DEVICE_MAX_CONNECTIONS is the application's to set — bench.py does — a drop-in library must not edit the
    // process environment.)
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n == 0) {
        g_create_err = std::string("no CUDA device: ") + cudaGetErrorString(e);
        return nullptr;
    }
    if (device < 0 || device >= n) { g_create_err = "device index out of range"; return nullptr; }
    cudaDeviceProp p;
    if ((e = cudaGetDeviceProperties(&p, device)) != cudaSuccess) { g_create_err = cudaGetErrorString(e); return nullptr; }
    // the library carries ONE cubin, sm_100a: arch-specific ("a") targets are not forward compatible, so any other 10.x part
    // would fail at the first launch with "no kernel image" instead of here
    if (p.major != 10 || p.minor != 0) {
        g_create_err = "this library is built for sm_100a (B200) only; found compute capability " +
                       std::to_string(p.major) + "." + std::to_string(p.minor);
        return nullptr;
    }
    if ((e = cudaSetDevice(device)) != cudaSuccess) { g_create_err = cudaGetErrorString(e); return nullptr; }
    svs_ctx *c = new (std::nothrow) svs_ctx();
    if (!c) { g_create_err = "out of memory"; return nullptr; }
    c->device = device;
    c->sm_count = p.multiProcessorCount;
    if (const char *z = getenv("SVS_ZC_CTAS")) { int v = atoi(z); if (v > 0) c->zc_ctas = v; }
    if ((e = cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking)) != cudaSuccess) {
        g_create_err = cudaGetErrorString(e);
        delete c;
        return nullptr;
    }
    if (const char *z = getenv("SVS_WAIT_MODE")) c->wait_mode = atoi(z) == 1 ? 1 : 0;
    if ((e = cudaEventCreateWithFlags(&c->ev_wait, cudaEventBlockingSync | cudaEventDisableTiming)) != cudaSuccess) {
        g_create_err = cudaGetErrorString(e);
        cudaStreamDestroy(c->stream);
        delete c;
        return nullptr;
    }
    g_live_ctx++;
    {   // experiment knobs (the API is svs_set_ba_schedule)
        const char *pz = getenv("SVS_BA_PRIO"), *tz = getenv("SVS_BA_THREADS");
        if (pz || tz) svs_set_ba_schedule(c, pz ? atoi(pz) : 0, tz ? atoi(tz) : 0);
    }
    return c;
}

int svs_set_wait_mode(svs_ctx *c, int mode)
{
    if (!c || (mode != 0 && mode != 1)) return SVS_ERR_ARG;
    c->wait_mode = mode;
    return SVS_OK;
}

// Grow every variable-size scratch buffer of the context to factor x the largest request it has seen, once, at a quiet
// point.  Buffers grow geometrically on demand, but a regrowth is a cudaFree + cudaMalloc, i.e. a device-wide synchronisation
// that stalls every other context of the process for tens of milliseconds; a batch whose per-step sizes fluctuate (the
// number of streams that insert a keyframe in a step) calls this after its warm-up so that steady state never regrows.
int svs_reserve_headroom(svs_ctx *c, double factor)
{
    if (!c || !(factor >= 1.0) || factor > 64.0) return SVS_ERR_ARG;
    SVS_CUDA(c, cudaSetDevice(c->device));
    SVS_CUDA(c, cudaStreamSynchronize(c->stream));
    if (c->stream_in) SVS_CUDA(c, cudaStreamSynchronize(c->stream_in));
    if (c->stream_ba) SVS_CUDA(c, cudaStreamSynchronize(c->stream_ba));
    DevBuf *d[] = {&c->d_in, &c->d_in2, &c->d_out, &c->d_out2, &c->d_tmp, &c->d_tmp2, &c->d_tmp3, &c->d_tmp4, &c->d_tmp5, &c->d_tmp6, &c->d_tmp7, &c->d_tmp8};
    std::vector<DevBuf *> dv(d, d + sizeof(d) / sizeof(d[0]));
    dv.insert(dv.end(), c->reg_dev.begin(), c->reg_dev.end());
    std::vector<PinBuf *> pv = {&c->h_in, &c->h_out};
    pv.insert(pv.end(), c->reg_pin.begin(), c->reg_pin.end());
    for (DevBuf *b : dv) if (b->hwm) SVS_CUDA(c, b->grow_keep((size_t)(factor * (double)b->hwm) + 4096));
    for (PinBuf *b : pv) if (b->hwm) SVS_CUDA(c, b->grow_keep((size_t)(factor * (double)b->hwm) + 4096));
    return SVS_OK;
}

int svs_set_ba_schedule(svs_ctx *c, int high_priority, int threads_per_window)
{
    if (!c || (threads_per_window != 0 && threads_per_window != 256 && threads_per_window != 512)) return SVS_ERR_ARG;
    SVS_CUDA(c, cudaSetDevice(c->device));
    c->ba_threads = threads_per_window;
    if (high_priority && !c->stream_ba) {
        int lo = 0, hi = 0;
        SVS_CUDA(c, cudaDeviceGetStreamPriorityRange(&lo, &hi));      // numerically lower = higher priority
        SVS_CUDA(c, cudaStreamCreateWithPriority(&c->stream_ba, cudaStreamNonBlocking, hi));
        SVS_CUDA(c, cudaEventCreateWithFlags(&c->ev_ba, cudaEventDisableTiming));
    } else if (!high_priority && c->stream_ba) {
        SVS_CUDA(c, cudaStreamSynchronize(c->stream_ba));
        cudaStreamDestroy(c->stream_ba); c->stream_ba = nullptr;
        cudaEventDestroy(c->ev_ba); c->ev_ba = nullptr;
    }
    return SVS_OK;
}

void svs_destroy(svs_ctx *c)
{
    if (!c) return;
    cudaSetDevice(c->device);
    cudaStreamSynchronize(c->stream);
    prof_harvest(c);
    for (cudaEvent_t e : c->prof_free) cudaEventDestroy(e);
    DevBuf *d[] = {&c->d_in, &c->d_in2, &c->d_out, &c->d_out2, &c->d_tmp, &c->d_tmp2, &c->d_tmp3, &c->d_tmp4, &c->d_tmp5, &c->d_tmp6, &c->d_tmp7, &c->d_tmp8};
    for (DevBuf *b : d) b->release();
    c->h_in.release(); c->h_out.release();
    svs_i_ba_ws_free(c->ba_ws); c->ba_ws = nullptr;
    cudaStreamDestroy(c->stream);
    if (c->stream_in) cudaStreamDestroy(c->stream_in);
    if (c->stream_ba) cudaStreamDestroy(c->stream_ba);
    if (c->ev_ba) cudaEventDestroy(c->ev_ba);
    if (c->ev_wait) cudaEventDestroy(c->ev_wait);
    { std::lock_guard<std::mutex> lk(g_mutex); g_live_ctx--; }
    delete c;
}

const char *svs_last_error(svs_ctx *c) { return c ? c->err.c_str() : "null context"; }
int svs_sync(svs_ctx *c) { SVS_CUDA(c, svs_i_wait(c)); if (c->prof_pending.size() > 4096) prof_harvest(c); return SVS_OK; }
void *svs_stream(svs_ctx *c) { return (void *)c->stream; }
long long svs_launch_count(svs_ctx *c) { return c->launches; }
void *svs_host_alloc(size_t bytes) { void *p = nullptr; return cudaMallocHost(&p, bytes) == cudaSuccess ? p : nullptr; }
void svs_host_free(void *p) { if (p) cudaFreeHost(p); }

// ------------------------------------------------------------------ frame sets
svs_frameset *svs_frameset_create(svs_ctx *c, int n_streams, int in_w, int in_h, int half, int lk_win, int lk_max_level)
{
    if (!c) return nullptr;
    if (n_streams <= 0 || in_w < 6 || in_h < 6 || lk_max_level < 0) { c->err = "frameset: bad arguments"; return nullptr; }
    cudaSetDevice(c->device);
    svs_frameset *fs = new (std::nothrow) svs_frameset();
    if (!fs) { c->err = "out of memory"; return nullptr; }
    fs->B = n_streams; fs->in_w = in_w; fs->in_h = in_h; fs->half = half; fs->win = lk_win;
    fs->W = half ? (int)std::nearbyint(in_w * 0.5) : in_w;
    fs->H = half ? (int)std::nearbyint(in_h * 0.5) : in_h;
    size_t per = 0;
    PyrDesc d;
    svs_i_make_pyr_desc(&d, fs->W, fs->H, lk_win, lk_max_level, &per);
    fs->nlev = d.nlev;
    for (int i = 0; i < 5; i++) {
        if (fs->pyr[i].reserve(per * n_streams) != cudaSuccess) {
            c->err = "frameset: cudaMalloc failed";
            svs_frameset_destroy(c, fs);
            return nullptr;
        }
        cudaMemsetAsync(fs->pyr[i].p, 0, per * n_streams, c->stream);
    }
    for (int i = 0; i < 3; i++) { fs->L[i] = d; fs->L[i].base = fs->pyr[i].as<uint8_t>(); }
    for (int i = 0; i < 2; i++) { fs->R[i] = d; fs->R[i].base = fs->pyr[3 + i].as<uint8_t>(); }
    // TMA tensor maps (u8 [stream][H_l][W_l], out-of-image bytes read as zero) for the bulk-async staged kernels
    fs->has_tmaps = true;
    for (int i = 0; i < 5 && fs->has_tmaps; i++) {
        const PyrDesc &pd = i < 3 ? fs->L[i] : fs->R[i - 3];
        for (int l = 0; l < pd.nlev; l++)
            if (svs_i_tmap_u8_3d(&fs->tm_pyr[i][l], pd.base + pd.off[l], pd.w[l], pd.h[l], n_streams, (size_t)pd.stride[l], pd.img_pitch,
                                 SVS_PD_BOX_W, SVS_PD_BOX_H) != 0) { fs->has_tmaps = false; break; }
        if (i < 3 && fs->has_tmaps &&
            svs_i_tmap_u8_3d(&fs->tm_gftt[i], pd.base + pd.off[0], pd.w[0], pd.h[0], n_streams, (size_t)pd.stride[0], pd.img_pitch,
                             SVS_CR_BOX_W, SVS_CR_BOX_H) != 0) fs->has_tmaps = false;
    }
    if (!fs->has_tmaps) { c->err = "frameset: cuTensorMapEncodeTiled failed"; svs_frameset_destroy(c, fs); return nullptr; }
    c->reg_dev.push_back(&fs->staging); c->reg_dev.push_back(&fs->pf_staging);
    if (cudaEventCreateWithFlags(&fs->pf_done, cudaEventDisableTiming) != cudaSuccess ||
        cudaEventCreateWithFlags(&fs->pf_order, cudaEventDisableTiming) != cudaSuccess) {
        c->err = "frameset: cudaEventCreate failed";
        svs_frameset_destroy(c, fs);
        return nullptr;
    }
    return fs;
}

void svs_frameset_destroy(svs_ctx *c, svs_frameset *fs)
{
    if (!fs) return;
    if (c) { cudaSetDevice(c->device); cudaStreamSynchronize(c->stream); if (c->stream_in) cudaStreamSynchronize(c->stream_in); }
    if (c) { c->unregister(&fs->staging); c->unregister(&fs->pf_staging); }
    for (int i = 0; i < 5; i++) fs->pyr[i].release();
    fs->staging.release();
    fs->ptr_table.release();
    fs->ptr_table_h.release();
    fs->pf_staging.release();
    fs->pf_ptr_table.release();
    fs->pf_ptr_table_h.release();
    if (fs->pf_done) cudaEventDestroy(fs->pf_done);
    if (fs->pf_order) cudaEventDestroy(fs->pf_order);
    delete fs;
}

int svs_frameset_size(const svs_frameset *fs, int *w, int *h, int *n_levels)
{
    if (!fs) return SVS_ERR_ARG;
    if (w) *w = fs->W;
    if (h) *h = fs->H;
    if (n_levels) *n_levels = fs->nlev;
    return SVS_OK;
}

int svs_frameset_push(svs_ctx *c, svs_frameset *fs, const uint8_t *left, const uint8_t *right, size_t row_stride,
                      size_t img_stride, int on_device)
{
    if (!c || !fs || !left || !right) return SVS_ERR_ARG;
    if (row_stride < (size_t)fs->in_w || img_stride < row_stride * fs->in_h) SVS_FAIL(c, SVS_ERR_ARG, "frameset_push: bad strides");
    SVS_CUDA(c, cudaSetDevice(c->device));
    const uint8_t *dl = left, *dr = right;
    size_t rs = row_stride, is = img_stride;
    int decim = 0;
    if (!on_device) {
        // the half-resolution resize reads only the even rows (dst[y][x] = src[2y][2x]): only those cross PCIe
        const int rows = fs->half ? fs->H : fs->in_h;
        const size_t src_pitch = fs->half ? 2 * row_stride : row_stride;
        size_t dense = (size_t)fs->in_w * rows;
        SVS_CUDA(c, fs->staging.reserve(2 * dense * fs->B));
        uint8_t *sl = fs->staging.as<uint8_t>(), *sr = sl + dense * fs->B;
        for (int b = 0; b < fs->B; b++) {
            SVS_CUDA(c, cudaMemcpy2DAsync(sl + dense * b, fs->in_w, left + img_stride * b, src_pitch, fs->in_w, rows, cudaMemcpyHostToDevice, c->stream));
            SVS_CUDA(c, cudaMemcpy2DAsync(sr + dense * b, fs->in_w, right + img_stride * b, src_pitch, fs->in_w, rows, cudaMemcpyHostToDevice, c->stream));
        }
        dl = sl; dr = sr; rs = fs->in_w; is = dense; decim = fs->half ? 1 : 0;
    }
    SVS_TRY(frameset_begin_push(c, fs));
    if (!on_device) fs->h2d_bytes += 2ll * fs->B * fs->in_w * (fs->half ? fs->H : fs->in_h);
    return frameset_fill(c, fs, fs->L[fs->il_cur], &fs->R[fs->ir_cur], dl, dr, rs, is, nullptr, nullptr, 0, decim, 0, fs->B, nullptr);
}

// resize (or copy) + pyramids of n_img images per eye into the given target buffers, on c->stream.  Rc == nullptr: one eye
// only (the images go to Lc).  ids_dev != nullptr: source image k goes to pyramid slot ids_dev[k] (a subset of the streams).
static int frameset_fill(svs_ctx *c, svs_frameset *fs, const PyrDesc &Lc, const PyrDesc *Rc, const uint8_t *dl, const uint8_t *dr,
                         size_t rs, size_t is, const uint8_t *const *pl, const uint8_t *const *pr, int aligned4, int rows_decimated,
                         int zero_copy, int n_img, const int32_t *ids_dev)
{
    if (fs->half && zero_copy && pl && (!Rc || pr == pl + n_img)) {
        // frames live in pinned host memory: small persistent grid, both eyes in one launch (images.cu)
        SVS_TRY(svs_i_half_nearest_zc(c, pl, n_img, fs->in_w, fs->in_h, rs, Lc.base + Lc.off[0], Rc ? Rc->base + Rc->off[0] : nullptr,
                                      fs->W, fs->H, Lc.stride[0], Lc.img_pitch, aligned4, ids_dev));
    } else if (fs->half) {
        SVS_TRY(svs_i_half_nearest(c, dl, fs->in_w, fs->in_h, rs, is, n_img, Lc.base + Lc.off[0], fs->W, fs->H, Lc.stride[0], Lc.img_pitch, pl, aligned4, rows_decimated, ids_dev));
        if (Rc) SVS_TRY(svs_i_half_nearest(c, dr, fs->in_w, fs->in_h, rs, is, n_img, Rc->base + Rc->off[0], fs->W, fs->H, Rc->stride[0], Rc->img_pitch, pr, aligned4, rows_decimated, ids_dev));
    } else {
        SVS_TRY(svs_i_copy_level0(c, dl, fs->in_w, fs->in_h, rs, is, n_img, Lc, pl, ids_dev));
        if (Rc) SVS_TRY(svs_i_copy_level0(c, dr, fs->in_w, fs->in_h, rs, is, n_img, *Rc, pr, ids_dev));
    }
    SVS_TRY(svs_i_build_pyramid(c, Lc, n_img, ids_dev, fs->maps_of(Lc)));
    if (Rc) SVS_TRY(svs_i_build_pyramid(c, *Rc, n_img, ids_dev, fs->maps_of(*Rc)));
    return SVS_OK;
}

// A push without a usable prefetch: order after any pending prefetch (it wrote the buffers that become "current"),
// drop it, rotate the buffer roles.
static int frameset_begin_push(svs_ctx *c, svs_frameset *fs)
{
    if (fs->pf_pending) {
        SVS_CUDA(c, cudaStreamWaitEvent(c->stream, fs->pf_done, 0));
        fs->pf_pending = false;
        fs->prefetch_misses++;
    }
    fs->rotate();
    fs->pushes++;
    return SVS_OK;
}

// Host pointers (pinned or pageable) -> staged copies of the rows the resize reads; device / zero-copy pointers ->
// pointer table.  Everything is enqueued on c->stream (the caller may have swapped in the ingest stream).
// right == nullptr / Rc == nullptr: one eye.  ids_host != nullptr: n_img images for the pyramid slots ids_host[k].
static int frameset_ingest_ptrs(svs_ctx *c, svs_frameset *fs, const PyrDesc &Lc, const PyrDesc *Rc, const uint8_t *const *left,
                                const uint8_t *const *right, size_t row_stride, int on_device, DevBuf &ptr_table, PinBuf &ptr_table_h,
                                DevBuf &staging, bool table_may_be_in_flight, int n_img, const int32_t *ids_host)
{
    const int eyes = Rc ? 2 : 1;
    const int rows = fs->half ? fs->H : fs->in_h;
    if (on_device != 1) fs->h2d_bytes += (long long)eyes * n_img * fs->in_w * rows;     // image bytes that cross PCIe
    // table: [left pointers | right pointers | slot ids]
    const size_t ptr_b = (size_t)eyes * n_img * sizeof(void *), ids_b = ids_host ? align_up((size_t)n_img * 4, 8) : 0;
    SVS_CUDA(c, ptr_table.reserve(ptr_b + ids_b + 16));
    SVS_CUDA(c, ptr_table_h.reserve(ptr_b + ids_b + 16));
    // the previous table copy must have been consumed before the pinned table is overwritten
    if (table_may_be_in_flight) SVS_CUDA(c, svs_i_wait(c));
    const uint8_t **hp = ptr_table_h.as<const uint8_t *>();
    int aligned4 = 1;
    for (int b = 0; b < n_img; b++) {
        hp[b] = left[b];
        if (Rc) hp[n_img + b] = right[b];
        if ((reinterpret_cast<uintptr_t>(left[b]) | (Rc ? reinterpret_cast<uintptr_t>(right[b]) : 0)) & 3) aligned4 = 0;
    }
    if (ids_host) memcpy(ptr_table_h.as<uint8_t>() + ptr_b, ids_host, (size_t)n_img * 4);
    const int32_t *ids_dev = ids_host ? reinterpret_cast<const int32_t *>(ptr_table.as<uint8_t>() + ptr_b) : nullptr;
    // on_device == 2: the pointers are PINNED HOST memory that the device can address (cudaHostAlloc / UVA): the resize
    // kernel reads the frames straight over PCIe (each needed row exactly once), no staging copy.
    if (on_device) {
        SVS_CUDA(c, cudaMemcpyAsync(ptr_table.p, hp, ptr_b + ids_b, cudaMemcpyHostToDevice, c->stream));
        const uint8_t *const *dp = ptr_table.as<const uint8_t *>();
        return frameset_fill(c, fs, Lc, Rc, nullptr, nullptr, row_stride, 0, dp, Rc ? dp + n_img : nullptr, aligned4, 0, on_device == 2,
                             n_img, ids_dev);
    }
    // Staged path: only the even rows the half-resolution resize reads cross PCIe (strided 2-D DMA copies); the kernel
    // then reads the staged rows with a unit row step.
    if (ids_host) SVS_CUDA(c, cudaMemcpyAsync(ptr_table.as<uint8_t>() + ptr_b, ptr_table_h.as<uint8_t>() + ptr_b, ids_b, cudaMemcpyHostToDevice, c->stream));
    const size_t src_pitch = fs->half ? 2 * row_stride : row_stride;
    size_t dense = (size_t)fs->in_w * rows;
    SVS_CUDA(c, staging.reserve((size_t)eyes * dense * n_img));
    uint8_t *sl = staging.as<uint8_t>(), *sr = sl + dense * n_img;
    for (int b = 0; b < n_img; b++) {
        SVS_CUDA(c, cudaMemcpy2DAsync(sl + dense * b, fs->in_w, left[b], src_pitch, fs->in_w, rows, cudaMemcpyHostToDevice, c->stream));
        if (Rc) SVS_CUDA(c, cudaMemcpy2DAsync(sr + dense * b, fs->in_w, right[b], src_pitch, fs->in_w, rows, cudaMemcpyHostToDevice, c->stream));
    }
    return frameset_fill(c, fs, Lc, Rc, sl, sr, fs->in_w, dense, nullptr, nullptr, 0, fs->half ? 1 : 0, 0, n_img, ids_dev);
}

// right == NULL: LEFT eye only (the right images are fetched later, for the streams that need them, with
// svs_frameset_fetch_right_ptrs).
int svs_frameset_push_ptrs(svs_ctx *c, svs_frameset *fs, const uint8_t *const *left, const uint8_t *const *right, size_t row_stride,
                           int on_device)
{
    if (!c || !fs || !left) return SVS_ERR_ARG;
    if (row_stride < (size_t)fs->in_w) SVS_FAIL(c, SVS_ERR_ARG, "frameset_push_ptrs: bad row stride");
    SVS_CUDA(c, cudaSetDevice(c->device));
    const int B = fs->B;
    if (fs->pf_pending && fs->pf_mode == on_device && fs->pf_row_stride == row_stride && fs->pf_has_right == (right != nullptr) &&
        memcmp(fs->pf_ptrs.data(), left, (size_t)B * sizeof(void *)) == 0 &&
        (!right || memcmp(fs->pf_ptrs.data() + B, right, (size_t)B * sizeof(void *)) == 0)) {
        // this pair was prefetched into the "next" buffers on the ingest stream: rotate and order the main stream after it
        fs->pf_pending = false;
        fs->prefetch_hits++;
        fs->rotate();
        fs->pushes++;
        SVS_CUDA(c, cudaStreamWaitEvent(c->stream, fs->pf_done, 0));
        return SVS_OK;
    }
    SVS_TRY(frameset_begin_push(c, fs));
    return frameset_ingest_ptrs(c, fs, fs->L[fs->il_cur], right ? &fs->R[fs->ir_cur] : nullptr, left, right, row_stride, on_device,
                                fs->ptr_table, fs->ptr_table_h, fs->staging, true, B, nullptr);
}

int svs_frameset_prefetch_ptrs(svs_ctx *c, svs_frameset *fs, const uint8_t *const *left, const uint8_t *const *right,
                               size_t row_stride, int on_device)
{
    if (!c || !fs || !left) return SVS_ERR_ARG;
    if (row_stride < (size_t)fs->in_w) SVS_FAIL(c, SVS_ERR_ARG, "frameset_prefetch_ptrs: bad row stride");
    SVS_CUDA(c, cudaSetDevice(c->device));
    if (!c->stream_in) SVS_CUDA(c, cudaStreamCreateWithFlags(&c->stream_in, cudaStreamNonBlocking));
    const int B = fs->B;
    // the previous prefetch's kernels and its pointer-table copy must be done before its buffers / tables are reused
    // (a never-recorded event is complete)
    SVS_CUDA(c, cudaEventSynchronize(fs->pf_done));
    if (fs->pf_pending) { fs->pf_pending = false; fs->prefetch_misses++; }
    // the "next" buffers may still be read by work already queued on the main stream
    SVS_CUDA(c, cudaEventRecord(fs->pf_order, c->stream));
    SVS_CUDA(c, cudaStreamWaitEvent(c->stream_in, fs->pf_order, 0));
    fs->pf_ptrs.resize((size_t)2 * B);
    memcpy(fs->pf_ptrs.data(), left, (size_t)B * sizeof(void *));
    if (right) memcpy(fs->pf_ptrs.data() + B, right, (size_t)B * sizeof(void *));
    fs->pf_mode = on_device; fs->pf_row_stride = row_stride; fs->pf_has_right = right != nullptr;
    cudaStream_t main_stream = c->stream;
    c->stream = c->stream_in;           // every svs_i_* launch below goes to the ingest stream
    int rc = frameset_ingest_ptrs(c, fs, fs->L[fs->il_next], right ? &fs->R[fs->ir_next] : nullptr, left, right, row_stride, on_device,
                                  fs->pf_ptr_table, fs->pf_ptr_table_h, fs->pf_staging, false, B, nullptr);
    cudaError_t e = cudaEventRecord(fs->pf_done, c->stream_in);
    c->stream = main_stream;
    if (rc != SVS_OK) return rc;
    SVS_CUDA(c, e);
    fs->pf_pending = true;
    return SVS_OK;
}

// Lazy right-eye ingest: resize + pyramids of the CURRENT right image of the selected streams only (the frontend reads
// the right image only when a stream inserts a keyframe: Frontend::FindFeaturesInRight).
int svs_frameset_fetch_right_ptrs(svs_ctx *c, svs_frameset *fs, const int32_t *stream_ids, int n_sel, const uint8_t *const *right,
                                  size_t row_stride, int on_device)
{
    if (!c || !fs || n_sel < 0 || (n_sel > 0 && (!stream_ids || !right))) return SVS_ERR_ARG;
    if (n_sel == 0) return SVS_OK;
    if (row_stride < (size_t)fs->in_w) SVS_FAIL(c, SVS_ERR_ARG, "frameset_fetch_right_ptrs: bad row stride");
    for (int i = 0; i < n_sel; i++) if (stream_ids[i] < 0 || stream_ids[i] >= fs->B) SVS_FAIL(c, SVS_ERR_ARG, "frameset_fetch_right_ptrs: stream id out of range");
    SVS_CUDA(c, cudaSetDevice(c->device));
    return frameset_ingest_ptrs(c, fs, fs->R[fs->ir_cur], nullptr, right, nullptr, row_stride, on_device, fs->ptr_table, fs->ptr_table_h,
                                fs->staging, true, n_sel, stream_ids);
}

long long svs_frameset_h2d_bytes(const svs_frameset *fs) { return fs ? fs->h2d_bytes : 0; }

int svs_frameset_download(svs_ctx *c, svs_frameset *fs, int stream, int which, int level, uint8_t *out, int out_stride)
{
    if (!c || !fs || stream < 0 || stream >= fs->B || level < 0 || level >= fs->nlev || which < 0 || which > 2) return SVS_ERR_ARG;
    SVS_CUDA(c, cudaSetDevice(c->device));
    const PyrDesc &d = which == 0 ? fs->Lcur() : which == 1 ? fs->Lprev() : fs->Rcur();
    SVS_CUDA(c, cudaMemcpy2DAsync(out, out_stride, d.base + (size_t)stream * d.img_pitch + d.off[level], d.stride[level],
                                  d.w[level], d.h[level], cudaMemcpyDeviceToHost, c->stream));
    SVS_CUDA(c, svs_i_wait(c));
    return SVS_OK;
}

// ------------------------------------------------------------------ a0
int svs_half_nearest(svs_ctx *c, const uint8_t *src, int w, int h, int stride, int n, size_t img_stride, uint8_t *dst)
{
    if (!c || !src || !dst || w < 1 || h < 1 || n < 0 || stride < w) return SVS_ERR_ARG;
    SVS_CUDA(c, cudaSetDevice(c->device));
    int dw = (int)std::nearbyint(w * 0.5), dh = (int)std::nearbyint(h * 0.5);
    size_t in_bytes = img_stride * (size_t)n, out_bytes = (size_t)dw * dh * n;
    SVS_CUDA(c, c->d_in.reserve(in_bytes));
    SVS_CUDA(c, c->d_out.reserve(out_bytes));
    SVS_CUDA(c, cudaMemcpyAsync(c->d_in.p, src, in_bytes, cudaMemcpyHostToDevice, c->stream));
    SVS_TRY(svs_i_half_nearest(c, c->d_in.as<uint8_t>(), w, h, stride, img_stride, n, c->d_out.as<uint8_t>(), dw, dh, dw, (size_t)dw * dh));
    SVS_CUDA(c, cudaMemcpyAsync(dst, c->d_out.p, out_bytes, cudaMemcpyDeviceToHost, c->stream));
    SVS_CUDA(c, svs_i_wait(c));
    return SVS_OK;
}

// ------------------------------------------------------------------ a1
static int gftt_common(svs_ctx *c, const uint8_t *img_dev, int w, int h, int stride, size_t img_pitch, int n_img,
                       const int32_t *ids_host, const uint8_t *mask_host, int mask_stride,
                       const int32_t *occ_off_host, const float *occ_xy_host, int max_corners, double quality,
                       double min_distance, int granule, float *out_xy, float *out_resp, int32_t *out_n, const CUtensorMap *tm = nullptr)
{
    if (max_corners <= 0) SVS_FAIL(c, SVS_ERR_ARG, "gftt: max_corners must be > 0");
    int n_occ = occ_off_host ? occ_off_host[n_img] : 0;
    // device inputs: [ids][occ_off][occ_xy] in d_in2 ; mask in d_tmp5
    size_t ids_b = align_up((size_t)n_img * 4, 16), off_b = align_up((size_t)(n_img + 1) * 4, 16), xy_b = (size_t)n_occ * 8;
    SVS_CUDA(c, c->d_in2.reserve(ids_b + off_b + xy_b + 16));
    SVS_CUDA(c, c->h_in.reserve(ids_b + off_b + xy_b + 16));
    uint8_t *hb = c->h_in.as<uint8_t>(), *db = c->d_in2.as<uint8_t>();
    const int32_t *ids_dev = nullptr, *off_dev = nullptr;
    const float *xy_dev = nullptr;
    if (ids_host) { memcpy(hb, ids_host, (size_t)n_img * 4); ids_dev = reinterpret_cast<int32_t *>(db); }
    if (occ_off_host && n_occ > 0) {
        memcpy(hb + ids_b, occ_off_host, (size_t)(n_img + 1) * 4);
        memcpy(hb + ids_b + off_b, occ_xy_host, xy_b);
        off_dev = reinterpret_cast<int32_t *>(db + ids_b);
        xy_dev = reinterpret_cast<float *>(db + ids_b + off_b);
    }
    SVS_CUDA(c, cudaMemcpyAsync(db, hb, ids_b + off_b + xy_b, cudaMemcpyHostToDevice, c->stream));
    const uint8_t *mask_dev = nullptr;
    if (mask_host) {
        SVS_CUDA(c, c->d_tmp5.reserve((size_t)w * h * n_img));
        SVS_CUDA(c, cudaMemcpy2DAsync(c->d_tmp5.p, w, mask_host, mask_stride, w, (size_t)h * n_img, cudaMemcpyHostToDevice, c->stream));
        mask_dev = c->d_tmp5.as<uint8_t>();
    }
    size_t oxy_b = (size_t)n_img * max_corners * 8, orp_b = (size_t)n_img * max_corners * 4, on_b = (size_t)n_img * 4;
    SVS_CUDA(c, c->d_out.reserve(oxy_b + orp_b + on_b));
    SVS_CUDA(c, c->h_out.reserve(oxy_b + orp_b + on_b + 16));
    uint8_t *dob = c->d_out.as<uint8_t>();
    SVS_TRY(svs_i_gftt(c, img_dev, w, h, stride, img_pitch, n_img, ids_dev, mask_dev, off_dev, xy_dev, n_occ, max_corners,
                       quality, min_distance, granule, reinterpret_cast<float *>(dob), reinterpret_cast<float *>(dob + oxy_b),
                       reinterpret_cast<int32_t *>(dob + oxy_b + orp_b), nullptr, tm));
    uint8_t *hob = c->h_out.as<uint8_t>();
    SVS_CUDA(c, cudaMemcpyAsync(hob, dob, oxy_b + orp_b + on_b, cudaMemcpyDeviceToHost, c->stream));
    int *ovf = reinterpret_cast<int *>(hob + align_up(oxy_b + orp_b + on_b, 4));
    SVS_TRY(svs_i_gftt_overflow(c, n_img, ovf));
    SVS_CUDA(c, svs_i_wait(c));
    if (*ovf) SVS_FAIL(c, SVS_ERR_CAPACITY, "gftt: more local-maximum candidates than the candidate buffer holds (w*h/4)");
    memcpy(out_xy, hob, oxy_b);
    if (out_resp) memcpy(out_resp, hob + oxy_b, orp_b);
    memcpy(out_n, hob + oxy_b + orp_b, on_b);
    return SVS_OK;
}

int svs_gftt_detect(svs_ctx *c, const uint8_t *img, int w, int h, int stride, const uint8_t *mask, int mask_stride,
                    const float *occupied_xy, int n_occupied, int max_corners, double quality, double min_distance,
                    int granule, float *out_xy, float *out_response, int *out_n)
{
    if (!c || !img || !out_xy || !out_n || stride < w) return SVS_ERR_ARG;
    SVS_CUDA(c, cudaSetDevice(c->device));
    int dstride = (int)align_up((size_t)w, 16);
    SVS_CUDA(c, c->d_in.reserve((size_t)dstride * h));
    SVS_CUDA(c, cudaMemcpy2DAsync(c->d_in.p, dstride, img, stride, w, h, cudaMemcpyHostToDevice, c->stream));
    int32_t off[2] = {0, n_occupied};
    int32_t n = 0;
    int r = gftt_common(c, c->d_in.as<uint8_t>(), w, h, dstride, 0, 1, nullptr, mask, mask_stride,
                        (occupied_xy && n_occupied > 0) ? off : nullptr, occupied_xy, max_corners, quality, min_distance,
                        granule, out_xy, out_response, &n);
    *out_n = n;
    return r;
}

int svs_corner_min_eig(svs_ctx *c, const uint8_t *img, int w, int h, int stride, int granule, float *out)
{
    if (!c || !img || !out || stride < w) return SVS_ERR_ARG;
    SVS_CUDA(c, cudaSetDevice(c->device));
    int dstride = (int)align_up((size_t)w, 16);
    SVS_CUDA(c, c->d_in.reserve((size_t)dstride * h));
    SVS_CUDA(c, c->d_out2.reserve((size_t)w * h * 4));
    SVS_CUDA(c, cudaMemcpy2DAsync(c->d_in.p, dstride, img, stride, w, h, cudaMemcpyHostToDevice, c->stream));
    SVS_TRY(svs_i_gftt(c, c->d_in.as<uint8_t>(), w, h, dstride, 0, 1, nullptr, nullptr, nullptr, nullptr, 0, 0, 0.01, 1.0,
                       granule, nullptr, nullptr, nullptr, c->d_out2.as<float>()));
    SVS_CUDA(c, cudaMemcpyAsync(out, c->d_out2.p, (size_t)w * h * 4, cudaMemcpyDeviceToHost, c->stream));
    SVS_CUDA(c, svs_i_wait(c));
    return SVS_OK;
}

int svs_gftt_detect_batch(svs_ctx *c, svs_frameset *fs, const int32_t *stream_ids, int n_sel, const int32_t *occ_off,
                          const float *occupied_xy, int max_corners, double quality, double min_distance, int granule,
                          float *out_xy, float *out_response, int32_t *out_n)
{
    if (!c || !fs || !stream_ids || !out_xy || !out_n || n_sel < 0) return SVS_ERR_ARG;
    if (n_sel == 0) return SVS_OK;
    for (int i = 0; i < n_sel; i++) if (stream_ids[i] < 0 || stream_ids[i] >= fs->B) SVS_FAIL(c, SVS_ERR_ARG, "gftt_batch: stream id out of range");
    SVS_CUDA(c, cudaSetDevice(c->device));
    const PyrDesc &L = fs->Lcur();
    return gftt_common(c, L.base + L.off[0], fs->W, fs->H, L.stride[0], L.img_pitch, n_sel, stream_ids, nullptr, 0, occ_off,
                       occupied_xy, max_corners, quality, min_distance, granule, out_xy, out_response, out_n,
                       fs->has_tmaps ? &fs->tm_gftt[fs->il_cur] : nullptr);
}

// ------------------------------------------------------------------ a2 / a3
static int lk_common(svs_ctx *c, const PyrDesc &prev, const PyrDesc &next, int n_img, const int32_t *off_host,
                     const float *prev_xy, float *next_xy, int win, int max_iter, double eps, uint8_t *status)
{
    int n = off_host[n_img];
    if (n <= 0) return SVS_OK;
    size_t id_b = align_up((size_t)n * 4, 16), xy_b = (size_t)n * 8;
    SVS_CUDA(c, c->h_in.reserve(id_b + 2 * xy_b));
    SVS_CUDA(c, c->d_in2.reserve(id_b + 2 * xy_b));
    SVS_CUDA(c, c->d_out.reserve(n));
    SVS_CUDA(c, c->h_out.reserve(xy_b + n));
    uint8_t *hb = c->h_in.as<uint8_t>(), *db = c->d_in2.as<uint8_t>();
    int32_t *ids = reinterpret_cast<int32_t *>(hb);
    for (int b = 0; b < n_img; b++) for (int i = off_host[b]; i < off_host[b + 1]; i++) ids[i] = b;
    memcpy(hb + id_b, prev_xy, xy_b);
    memcpy(hb + id_b + xy_b, next_xy, xy_b);
    SVS_CUDA(c, cudaMemcpyAsync(db, hb, id_b + 2 * xy_b, cudaMemcpyHostToDevice, c->stream));
    float *nxt_dev = reinterpret_cast<float *>(db + id_b + xy_b);
    SVS_TRY(svs_i_lk(c, prev, next, reinterpret_cast<int32_t *>(db), reinterpret_cast<float *>(db + id_b), nxt_dev, n, win,
                     max_iter, eps, c->d_out.as<uint8_t>()));
    uint8_t *ho = c->h_out.as<uint8_t>();
    SVS_CUDA(c, cudaMemcpyAsync(ho, nxt_dev, xy_b, cudaMemcpyDeviceToHost, c->stream));
    SVS_CUDA(c, cudaMemcpyAsync(ho + xy_b, c->d_out.p, n, cudaMemcpyDeviceToHost, c->stream));
    SVS_CUDA(c, svs_i_wait(c));
    memcpy(next_xy, ho, xy_b);
    memcpy(status, ho + xy_b, n);
    return SVS_OK;
}

int svs_lk_track(svs_ctx *c, const uint8_t *prev, const uint8_t *next, int w, int h, int stride, const float *prev_xy,
                 float *next_xy, int n, int win, int max_level, int max_iter, double eps, uint8_t *status)
{
    if (!c || !prev || !next || n < 0 || stride < w || (n > 0 && (!prev_xy || !next_xy || !status))) return SVS_ERR_ARG;
    if (n == 0) return SVS_OK;
    SVS_CUDA(c, cudaSetDevice(c->device));
    PyrDesc dp, dn;
    size_t per = 0;
    svs_i_make_pyr_desc(&dp, w, h, win, max_level, &per);
    dn = dp;
    SVS_CUDA(c, c->d_tmp6.reserve(2 * per));
    dp.base = c->d_tmp6.as<uint8_t>();
    dn.base = dp.base + per;
    SVS_CUDA(c, cudaMemcpy2DAsync(dp.base, dp.stride[0], prev, stride, w, h, cudaMemcpyHostToDevice, c->stream));
    SVS_CUDA(c, cudaMemcpy2DAsync(dn.base, dn.stride[0], next, stride, w, h, cudaMemcpyHostToDevice, c->stream));
    SVS_TRY(svs_i_build_pyramid(c, dp, 1));
    SVS_TRY(svs_i_build_pyramid(c, dn, 1));
    int32_t off[2] = {0, n};
    return lk_common(c, dp, dn, 1, off, prev_xy, next_xy, win, max_iter, eps, status);
}

int svs_lk_track_batch(svs_ctx *c, svs_frameset *fs, int pair, const int32_t *off, const float *prev_xy, float *next_xy,
                       int max_iter, double eps, uint8_t *status)
{
    if (!c || !fs || !off || (pair != 0 && pair != 1)) return SVS_ERR_ARG;
    SVS_CUDA(c, cudaSetDevice(c->device));
    const PyrDesc &prev = pair == 0 ? fs->Lprev() : fs->Lcur();
    const PyrDesc &next = pair == 0 ? fs->Lcur() : fs->Rcur();
    return lk_common(c, prev, next, fs->B, off, prev_xy, next_xy, fs->win, max_iter, eps, status);
}

}  // extern "C"
