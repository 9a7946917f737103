// Internal declarations shared by the CUDA translation units behind include/svslam.h.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>
#include "../../include/svslam.h"

#define SVS_MAX_LEVELS 8
// TMA boxes.  The innermost (byte) coordinate of a bulk tensor copy must be a multiple of 16 (measured: any other start traps
// with "illegal instruction", tests/tools/tma_probe.cu), so a tile's box starts 16 bytes left of its first output column:
// k_pyr_down_tma reads image bytes [2*ox0 - 16, 2*ox0 + 144) x 35 rows, the corner response [ox0 - 16, ox0 + 80) x 20 rows.
#define SVS_PD_BOX_W 160
#define SVS_PD_BOX_H 35
#define SVS_CR_BOX_W 96
#define SVS_CR_BOX_H 20

// Device image pyramid batch: image b, level l, row y at base + b*img_pitch + off[l] + y*stride[l].
struct PyrDesc {
    uint8_t *base;
    size_t img_pitch;
    int nlev;
    int w[SVS_MAX_LEVELS], h[SVS_MAX_LEVELS], stride[SVS_MAX_LEVELS];
    size_t off[SVS_MAX_LEVELS];
};

// Grow-only device / pinned-host buffers.  svs_i_regrowths counts every (re)allocation: a regrowth frees and allocates, which
// synchronises the whole device, so a steady-state step must not see one (bench.py reports the count per timed region).
extern long long svs_i_regrowths;
struct DevBuf {
    void *p = nullptr;
    size_t cap = 0;
    size_t hwm = 0;     // largest request so far (svs_reserve_headroom sizes the buffer from it)
    cudaError_t reserve(size_t bytes) {
        if (bytes > hwm) hwm = bytes;
        if (bytes <= cap) return cudaSuccess;
        if (p) cudaFree(p);
        p = nullptr; cap = 0;
        size_t want = 2 * bytes + 4096;   // geometric growth: regrowing is a device-wide sync
        svs_i_regrowths++;
        cudaError_t e = cudaMalloc(&p, want);
        if (e == cudaSuccess) cap = want;
        return e;
    }
    // deliberate growth at a quiet point (the caller has synchronised the device): contents are kept
    cudaError_t grow_keep(size_t want) {
        if (want <= cap) return cudaSuccess;
        void *q = nullptr;
        cudaError_t e = cudaMalloc(&q, want);
        if (e != cudaSuccess) return e;
        if (p) { e = cudaMemcpy(q, p, cap, cudaMemcpyDeviceToDevice); cudaFree(p); }
        p = q; cap = want;
        return e;
    }
    void release() { if (p) cudaFree(p); p = nullptr; cap = 0; }
    template <class T> T *as() { return reinterpret_cast<T *>(p); }
};
struct PinBuf {
    void *p = nullptr;
    size_t cap = 0;
    size_t hwm = 0;
    cudaError_t grow_keep(size_t want) {
        if (want <= cap) return cudaSuccess;
        void *q = nullptr;
        cudaError_t e = cudaMallocHost(&q, want);
        if (e != cudaSuccess) return e;
        if (p) { memcpy(q, p, cap); cudaFreeHost(p); }
        p = q; cap = want;
        return cudaSuccess;
    }
    cudaError_t reserve(size_t bytes) {
        if (bytes > hwm) hwm = bytes;
        if (bytes <= cap) return cudaSuccess;
        if (p) cudaFreeHost(p);
        p = nullptr; cap = 0;
        size_t want = 2 * bytes + 4096;
        svs_i_regrowths++;
        cudaError_t e = cudaMallocHost(&p, want);
        if (e == cudaSuccess) cap = want;
        return e;
    }
    void release() { if (p) cudaFreeHost(p); p = nullptr; cap = 0; }
    template <class T> T *as() { return reinterpret_cast<T *>(p); }
};

// Kernel classes for the optional per-kernel CUDA-event timing (svs_kernel_timing_*; bench.py's roofline).
enum SvsKernelId { KID_HALF = 0, KID_COPY0, KID_PYRDOWN, KID_MASK, KID_CORNER_RESPONSE, KID_CORNER_SELECT, KID_CORNER_GREEDY,
                   KID_LK, KID_TRIANGULATE, KID_POSE_LM, KID_BA_WINDOW, KID_BM_PREFILTER, KID_BM_SAD, KID_BACKPROJECT,
                   KID_BGR2GRAY, KID_MISC, KID_TRACK_STATE, KID_BA_BUILD, KID_COUNT };
struct SvsPendingEv { int kid; cudaEvent_t a, b; };

struct svs_ctx {
    bool prof = false;
    std::vector<SvsPendingEv> prof_pending;
    std::vector<cudaEvent_t> prof_free;
    double prof_ms[KID_COUNT] = {0};
    long long prof_n[KID_COUNT] = {0};
    int device = 0;
    int sm_count = 0;
    int host_threads = 0;   // OpenMP team size for host-side problem construction in this context's calls (0 = OpenMP's default)
    // Host wait policy: 0 = cudaStreamSynchronize (the driver spins: lowest latency, one busy core per waiting thread),
    // 1 = record + wait on a cudaEventBlockingSync event (the thread sleeps: for boxes with fewer cores than waiting threads)
    int wait_mode = 0;
    cudaEvent_t ev_wait = nullptr;
    int zc_ctas = 0;    // persistent grid of the zero-copy (PCIe) ingest kernel: 0 = automatic (svs_i_zc_grid), env SVS_ZC_CTAS overrides
    cudaStream_t stream = nullptr;
    cudaStream_t stream_ba = nullptr;   // optional high-priority stream of the window solver (svs_set_ba_schedule)
    cudaEvent_t ev_ba = nullptr;
    int ba_threads = 0;                 // 0 = automatic (512 per window, 256 x 2 per SM when a launch has more windows than SMs)
    cudaStream_t stream_in = nullptr;   // ingest stream: prefetch of the NEXT frame pair overlaps this step's compute
    std::string err;
    long long launches = 0;
    void *ba_ws = nullptr;              // svs_ba_optimize's persistent host workspace (ba.cu; freed by svs_i_ba_ws_free)
    double ba_host_s[3] = {0, 0, 0};   // svs_ba_optimize wall time: structure build | pack + enqueue | wait for the device + unpack
    // scratch (named by user)
    DevBuf d_in, d_in2, d_out, d_out2, d_tmp, d_tmp2, d_tmp3, d_tmp4, d_tmp5, d_tmp6, d_tmp7, d_tmp8;
    PinBuf h_in, h_out;
    // variable-size scratch buffers of objects created on this context (tracker uploads, frame-set staging): registered so
    // that svs_reserve_headroom reaches them; the owner unregisters before it dies
    std::vector<DevBuf *> reg_dev;
    std::vector<PinBuf *> reg_pin;
    void unregister(DevBuf *b) { for (size_t i = 0; i < reg_dev.size(); i++) if (reg_dev[i] == b) { reg_dev.erase(reg_dev.begin() + i); break; } }
    void unregister(PinBuf *b) { for (size_t i = 0; i < reg_pin.size(); i++) if (reg_pin[i] == b) { reg_pin.erase(reg_pin.begin() + i); break; } }
};

struct svs_frameset {
    int B = 0, in_w = 0, in_h = 0, half = 0, W = 0, H = 0, win = 11, nlev = 1;
    // Triple-buffered left pyramids (previous / current / next) and double-buffered right pyramids (current / next):
    // "next" is the target of svs_frameset_prefetch_ptrs, filled on the context's ingest stream while the current
    // step's kernels run; a push rotates the roles.
    PyrDesc L[3], R[2];
    // TMA tensor maps over [stream][H_l][W_l] of every level of the five pyramid buffers (box = k_pyr_down_tma's source tile)
    // and over level 0 with the corner-response tile box
    CUtensorMap tm_pyr[5][SVS_MAX_LEVELS];
    CUtensorMap tm_gftt[3];
    bool has_tmaps = false;
    const CUtensorMap *maps_of(const PyrDesc &d) const {
        if (!has_tmaps) return nullptr;
        for (int i = 0; i < 3; i++) if (d.base == L[i].base) return tm_pyr[i];
        for (int i = 0; i < 2; i++) if (d.base == R[i].base) return tm_pyr[3 + i];
        return nullptr;
    }
    int il_prev = 1, il_cur = 0, il_next = 2, ir_cur = 0, ir_next = 1;
    const PyrDesc &Lcur() const { return L[il_cur]; }
    const PyrDesc &Lprev() const { return L[il_prev]; }
    const PyrDesc &Rcur() const { return R[ir_cur]; }
    void rotate() { int p = il_prev; il_prev = il_cur; il_cur = il_next; il_next = p; int r = ir_cur; ir_cur = ir_next; ir_next = r; }
    long long pushes = 0;
    DevBuf pyr[5];     // storage of L[0..2], R[0..1]
    DevBuf staging;    // raw input images when pushed from the host
    DevBuf ptr_table;  // per-stream device pointers (push_ptrs with on_device)
    PinBuf ptr_table_h;
    // pending prefetch (svs_frameset_prefetch_ptrs)
    bool pf_pending = false, pf_has_right = true;
    long long h2d_bytes = 0;                // image bytes this frame set has read from host memory
    int pf_mode = 0;
    size_t pf_row_stride = 0;
    std::vector<const uint8_t *> pf_ptrs;   // [left.. | right..] as given by the caller
    cudaEvent_t pf_done = nullptr, pf_order = nullptr;
    DevBuf pf_staging, pf_ptr_table;
    PinBuf pf_ptr_table_h;
    long long prefetch_hits = 0, prefetch_misses = 0;
};

void svs_i_ba_ws_free(void *ws);

// wait for everything enqueued on the context stream, by the context's wait policy
inline cudaError_t svs_i_wait(svs_ctx *c)
{
    if (c->wait_mode == 0 || !c->ev_wait) return cudaStreamSynchronize(c->stream);
    cudaError_t e = cudaEventRecord(c->ev_wait, c->stream);
    return e != cudaSuccess ? e : cudaEventSynchronize(c->ev_wait);
}

#define SVS_CUDA(ctx, call)                                                              \
    do {                                                                                 \
        cudaError_t e__ = (call);                                                        \
        if (e__ != cudaSuccess) {                                                        \
            (ctx)->err = std::string(#call) + ": " + cudaGetErrorString(e__);            \
            return SVS_ERR_CUDA;                                                         \
        }                                                                                \
    } while (0)
#define SVS_FAIL(ctx, code, msg)                                                         \
    do { (ctx)->err = (msg); return (code); } while (0)
#define SVS_TRY(call)                                                                    \
    do { int r__ = (call); if (r__ != SVS_OK) return r__; } while (0)
#define SVS_LAUNCH_CHECK(ctx)                                                            \
    do { (ctx)->launches++; SVS_CUDA(ctx, cudaGetLastError()); } while (0)
// Bracket one kernel launch with CUDA events on the context stream when timing is enabled.
void svs_i_prof_begin(svs_ctx *c, int kid);
void svs_i_prof_end(svs_ctx *c);
#define SVS_KERNEL(ctx, kid, ...)                                                        \
    do { svs_i_prof_begin(ctx, kid); __VA_ARGS__; svs_i_prof_end(ctx); SVS_LAUNCH_CHECK(ctx); } while (0)

static inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

// Opt a kernel in to the device's maximum dynamic shared memory, once per (device, kernel), thread-safe: contexts are
// stepped from several host threads, and the attribute is per device.
cudaError_t svs_i_opt_in_smem(svs_ctx *c, const void *func);

// images.cu: u8 tensor map [n][h][w], byte strides (1, row_stride, img_pitch), box (box_w, box_h, 1); 0 on success
int svs_i_tmap_u8_3d(CUtensorMap *m, const void *base, int w, int h, int n, size_t row_stride, size_t img_pitch, int box_w, int box_h);

// ---- internal device-level entry points (all asynchronous on ctx->stream) ----
// images.cu
int svs_i_make_pyr_desc(PyrDesc *d, int w, int h, int win, int max_level, size_t *bytes_per_image);
int svs_i_half_nearest(svs_ctx *c, const uint8_t *src_dev, int w, int h, size_t row_stride, size_t img_stride,
                       int n, uint8_t *dst_dev, int dw, int dh, int dst_stride, size_t dst_img_pitch,
                       const uint8_t *const *src_ptrs_dev = nullptr, int ptrs_aligned4 = 0, int rows_decimated = 0,
                       const int32_t *dst_ids_dev = nullptr);
int svs_i_zc_grid(const svs_ctx *c);   // CTAs of the zero-copy ingest kernel: ~32 in total over all live contexts
int svs_i_half_nearest_zc(svs_ctx *c, const uint8_t *const *src_ptrs_dev, int n_per_eye, int w, int h, size_t row_stride,
                          uint8_t *dstL, uint8_t *dstR /* null: one eye */, int dw, int dh, int dst_stride, size_t dst_img_pitch,
                          int ptrs_aligned4, const int32_t *dst_ids_dev = nullptr);
int svs_i_copy_level0(svs_ctx *c, const uint8_t *src_dev, int w, int h, size_t row_stride, size_t img_stride,
                      int n, const PyrDesc &d, const uint8_t *const *src_ptrs_dev = nullptr, const int32_t *dst_ids_dev = nullptr);
int svs_i_build_pyramid(svs_ctx *c, const PyrDesc &d, int n_images, const int32_t *img_ids_dev = nullptr,
                        const CUtensorMap *level_maps = nullptr /* one per level: TMA-staged kernel */);
// gftt.cu
int svs_i_gftt(svs_ctx *c, const uint8_t *img_dev, int w, int h, int stride, size_t img_pitch, int n_img,
               const int32_t *img_ids_dev /* may be null: identity */,
               const uint8_t *mask_dev /* n_img*h*w or null */,
               const int32_t *occ_off_dev, const float *occ_xy_dev /* or null */, int n_occ_total,
               int max_corners, double quality, double min_distance, int granule,
               float *out_xy_dev, float *out_resp_dev, int32_t *out_n_dev, float *eig_out_dev /* optional */,
               const CUtensorMap *tm = nullptr /* u8 [slot][h][w] map with the SVS_CR box: TMA-staged tiled corner response */);
// lk.cu
int svs_i_lk(svs_ctx *c, const PyrDesc &prev, const PyrDesc &next, const int32_t *pt_img_dev /* per point */,
             const float *prev_xy_dev, float *next_xy_dev, int n_pts, int win, int max_iter, double eps,
             uint8_t *status_dev);
// geom.cu
int svs_i_pose_only_lm_dev(svs_ctx *c, int n_prob, const int32_t *off_dev, const int32_t *end_dev, const double *pts_w_dev,
                           const double *uv_dev, const double *K_dev, const double *T0_dev, double chi2_th, int rounds, int iters,
                           double *T_out_dev, uint8_t *outl_dev, int32_t *n_inlier_dev);
