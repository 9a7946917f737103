// Device-resident tracking state: Frontend::Track() for a batch of independent streams with NO host round trip between
// its seams (reference src/frontend.cpp:645-688).  Per stream the device keeps what the host Frontend keeps between two
// frames — the left features of the last frame (position, landmark id, landmark world position), the last pose and the
// relative motion — and one step is
//
//   k_trk_begin   motion model pose0 = relative_motion * last.Pose()  (:653-656); LK initial guesses = world2pixel of the
//                 landmark with pose0, else the old position (TrackLastFrame :331-347)
//   k_lk_track    calcOpticalFlowPyrLK last.left -> current.left (:353-357), csrc/lk.cu, reading the guesses in place
//   k_trk_mid     keep status && inside the image, inherit the landmark (:361-381); gather the pose-only problem (:439-471)
//   k_pose_only_lm  the g2o block (:408-527), csrc/geom.cu, reading the gathered edges in place
//   k_trk_end     pose, outliers lose their landmark (:546-553), status (:665-679), keyframe test (:587), relative motion
//                 (:685); one 128-byte record per stream — and, for streams that must insert a keyframe, their feature list —
//                 written straight into pinned host memory
//
// The SE3 / camera arithmetic is the host's own source (host/slam.h, __host__ __device__) and this file is compiled with
// --fmad=false like the host code's target has no fused multiply-add: the device-resident path is bit-identical to the
// host-driven one (tests/test_gpu_pipeline.py::test_device_tracking_is_bit_identical_to_host_tracking).
// One warp per stream; every list keeps the host's order (stable ballot compaction).
#include "track.h"
#include <cstring>
#include <vector>

using slam::CameraModel;
using slam::SE3;
using slam::Vec2;
using slam::Vec3;

#define TRK_WARPS 4

struct TrkDev {
    int B, cap;
    // per stream
    int32_t *status, *nfeat, *buf;      // FrontendStatus | features of the last frame | which of the two tables is "last"
    double *last_pose, *rel, *pose0;    // 7 each; pose0 = motion-model prediction = T0 of the pose-only problem
    double *K;                          // 4 per stream (k_pose_only_lm reads K per problem)
    // feature tables [2][B][cap]
    float2 *xy[2];
    int64_t *lm[2];
    double *pw[2];                      // 3 per feature
    // compact per-step arrays (points of stream b at off[b]..)
    int32_t *off, *lm_end, *pt_img, *lm_fidx, *n_inl;
    float2 *prev_xy, *next_xy;
    uint8_t *st, *outl;
    double *lm_pts, *lm_uv, *T_out;
    // pinned host (device-addressable)
    TrkOut *out;
    TrkFeat *kf_feats;                  // [B][cap]
};

struct svs_tracker {
    TrkParams p;
    TrkDev d;
    DevBuf mem, up_dev;
    PinBuf out_h, kf_h, off_h, up_h;
    std::vector<int32_t> h_nfeat, h_status;     // host mirrors (what the next step's offsets are built from)
    cudaEvent_t up_done = nullptr;
};

__device__ __forceinline__ bool trk_active(int st) { return st == 1 || st == 2; }

__global__ void __launch_bounds__(TRK_WARPS * 32)
k_trk_begin(TrkDev d, const __grid_constant__ CameraModel cam)
{
    const int b = blockIdx.x * TRK_WARPS + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (b >= d.B || !trk_active(d.status[b])) return;
    const SE3 T0 = SE3::fromArray(d.rel + 7 * (size_t)b) * SE3::fromArray(d.last_pose + 7 * (size_t)b);
    if (lane < 7) d.pose0[7 * (size_t)b + lane] = T0.d[lane];
    const int n = d.nfeat[b], base = d.off[b], q = d.buf[b];
    const float2 *xy = d.xy[q] + (size_t)b * d.cap;
    const int64_t *lm = d.lm[q] + (size_t)b * d.cap;
    const double *pw = d.pw[q] + 3 * (size_t)b * d.cap;
    for (int i = lane; i < n; i += 32) {
        const float2 p = xy[i];
        float2 g = p;
        if (lm[i] >= 0) {
            const Vec2 px = cam.world2pixel(Vec3(pw[3 * i], pw[3 * i + 1], pw[3 * i + 2]), T0);
            g = make_float2((float)px.x, (float)px.y);
        }
        d.prev_xy[base + i] = p;
        d.next_xy[base + i] = g;
        d.pt_img[base + i] = b;
    }
}

__global__ void __launch_bounds__(TRK_WARPS * 32)
k_trk_mid(TrkDev d, float W, float H)
{
    const int b = blockIdx.x * TRK_WARPS + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (b >= d.B) return;
    const int base = d.off[b];
    if (!trk_active(d.status[b])) { if (lane == 0) d.lm_end[b] = base; return; }
    const int n = d.nfeat[b], q = d.buf[b];
    const int64_t *lm = d.lm[q] + (size_t)b * d.cap;
    const double *pw = d.pw[q] + 3 * (size_t)b * d.cap;
    float2 *oxy = d.xy[q ^ 1] + (size_t)b * d.cap;
    int64_t *olm = d.lm[q ^ 1] + (size_t)b * d.cap;
    double *opw = d.pw[q ^ 1] + 3 * (size_t)b * d.cap;
    int kept = 0, edges = 0;
    const unsigned below = (1u << lane) - 1u;
    for (int i0 = 0; i0 < n; i0 += 32) {
        const int i = i0 + lane;
        bool keep = false, has = false;
        float2 p = make_float2(0.f, 0.f);
        int64_t id = -1;
        if (i < n) {
            p = d.next_xy[base + i];
            keep = d.st[base + i] != 0 && p.y >= 0.f && p.y < H && p.x >= 0.f && p.x < W;
            id = lm[i];
            has = keep && id >= 0;
        }
        const unsigned mk = __ballot_sync(0xffffffffu, keep), me = __ballot_sync(0xffffffffu, has);
        if (keep) {
            const int k = kept + __popc(mk & below);
            oxy[k] = p; olm[k] = id;
            double x = 0, y = 0, z = 0;
            if (id >= 0) { x = pw[3 * i]; y = pw[3 * i + 1]; z = pw[3 * i + 2]; }
            opw[3 * k] = x; opw[3 * k + 1] = y; opw[3 * k + 2] = z;
            if (has) {
                const int e = base + edges + __popc(me & below);
                d.lm_pts[3 * (size_t)e] = x; d.lm_pts[3 * (size_t)e + 1] = y; d.lm_pts[3 * (size_t)e + 2] = z;
                d.lm_uv[2 * (size_t)e] = (double)p.x; d.lm_uv[2 * (size_t)e + 1] = (double)p.y;
                d.lm_fidx[e] = k;
            }
        }
        kept += __popc(mk); edges += __popc(me);
    }
    if (lane == 0) { d.nfeat[b] = kept; d.lm_end[b] = base + edges; }
}

__global__ void __launch_bounds__(TRK_WARPS * 32)
k_trk_end(TrkDev d, int n_track, int n_bad, int n_kf)
{
    const int b = blockIdx.x * TRK_WARPS + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (b >= d.B) return;
    TrkOut *o = d.out + b;
    const int st0 = d.status[b];
    if (!trk_active(st0)) {
        if (lane == 0) { o->status = st0; o->inliers = 0; o->need_kf = 0; o->nfeat = d.nfeat[b]; o->n_edges = 0; }
        return;
    }
    const int q = d.buf[b] ^ 1;            // the table k_trk_mid filled becomes "last"
    const int base = d.off[b], end = d.lm_end[b];
    int64_t *lm = d.lm[q] + (size_t)b * d.cap;
    for (int e = base + lane; e < end; e += 32)
        if (d.outl[e]) lm[d.lm_fidx[e]] = -1;            // outliers lose their landmark (frontend.cpp:546-553)
    __syncwarp();
    const int inl = d.n_inl[b], n = d.nfeat[b];
    const int st = inl > n_track ? 1 : (inl > n_bad ? 2 : 3);
    const int kf = inl < n_kf ? 1 : 0;
    const SE3 T = SE3::fromArray(d.T_out + 7 * (size_t)b), Tl = SE3::fromArray(d.last_pose + 7 * (size_t)b);
    const SE3 rel = T * Tl.inverse();
    __syncwarp();                          // every lane has read last_pose before lanes 0-6 overwrite it
    if (lane < 7) {
        o->pose[lane] = T.d[lane]; o->last_pose[lane] = Tl.d[lane];
        d.rel[7 * (size_t)b + lane] = rel.d[lane];
        d.last_pose[7 * (size_t)b + lane] = T.d[lane];
    }
    if (lane == 0) { o->status = st; o->inliers = inl; o->need_kf = kf; o->nfeat = n; o->n_edges = end - base; d.status[b] = st; d.buf[b] = q; }
    if (kf) {     // this stream goes back to the host's InsertKeyframe path: hand it its tracked features
        const float2 *xy = d.xy[q] + (size_t)b * d.cap;
        TrkFeat *dst = d.kf_feats + (size_t)b * d.cap;
        for (int i = lane; i < n; i += 32) { TrkFeat f; f.x = xy[i].x; f.y = xy[i].y; f.lm = lm[i]; dst[i] = f; }
    }
}

__global__ void k_trk_scatter(TrkDev d, const TrkUpHdr *hdrs, const TrkUpFeat *feats)
{
    const TrkUpHdr h = hdrs[blockIdx.x];
    const int b = h.stream, q = d.buf[b];
    float2 *xy = d.xy[q] + (size_t)b * d.cap;
    int64_t *lm = d.lm[q] + (size_t)b * d.cap;
    double *pw = d.pw[q] + 3 * (size_t)b * d.cap;
    const TrkUpFeat *f = feats + h.feat_off;
    for (int i = threadIdx.x; i < h.n; i += blockDim.x) {
        xy[i] = make_float2(f[i].x, f[i].y); lm[i] = f[i].lm;
        pw[3 * i] = f[i].pw[0]; pw[3 * i + 1] = f[i].pw[1]; pw[3 * i + 2] = f[i].pw[2];
    }
    if (threadIdx.x < 7) { d.last_pose[7 * (size_t)b + threadIdx.x] = h.pose[threadIdx.x]; d.rel[7 * (size_t)b + threadIdx.x] = h.rel[threadIdx.x]; }
    if (threadIdx.x == 0) { d.nfeat[b] = h.n; d.status[b] = h.status; }
}

__global__ void k_trk_export(TrkDev d, int b)
{
    const int q = d.buf[b], n = d.nfeat[b];
    const float2 *xy = d.xy[q] + (size_t)b * d.cap;
    const int64_t *lm = d.lm[q] + (size_t)b * d.cap;
    TrkFeat *dst = d.kf_feats + (size_t)b * d.cap;
    for (int i = threadIdx.x; i < n; i += blockDim.x) { TrkFeat f; f.x = xy[i].x; f.y = xy[i].y; f.lm = lm[i]; dst[i] = f; }
    if (threadIdx.x == 0) d.out[b].nfeat = n;
}

__global__ void k_trk_init(TrkDev d, double fx, double fy, double cx, double cy)
{
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= d.B) return;
    d.status[b] = 0; d.nfeat[b] = 0; d.buf[b] = 0;
    d.K[4 * b] = fx; d.K[4 * b + 1] = fy; d.K[4 * b + 2] = cx; d.K[4 * b + 3] = cy;
    for (int i = 0; i < 7; i++) { d.last_pose[7 * b + i] = i == 3; d.rel[7 * b + i] = i == 3; d.pose0[7 * b + i] = i == 3; }
}

svs_tracker *svs_i_trk_create(svs_ctx *c, const TrkParams &p)
{
    if (cudaSetDevice(c->device) != cudaSuccess) return nullptr;
    svs_tracker *t = new (std::nothrow) svs_tracker();
    if (!t) return nullptr;
    t->p = p;
    const size_t B = p.B, cap = p.cap, tot = B * cap;
    // carve one device allocation
    size_t off = 0;
    auto take = [&](size_t bytes) { size_t o = off; off = align_up(off + bytes, 256); return o; };
    const size_t o_status = take(B * 4), o_nfeat = take(B * 4), o_buf = take(B * 4), o_last = take(B * 56), o_rel = take(B * 56),
                 o_pose0 = take(B * 56), o_K = take(B * 32), o_xy0 = take(tot * 8), o_xy1 = take(tot * 8), o_lm0 = take(tot * 8),
                 o_lm1 = take(tot * 8), o_pw0 = take(tot * 24), o_pw1 = take(tot * 24), o_off = take((B + 1) * 4), o_end = take(B * 4),
                 o_img = take(tot * 4), o_fidx = take(tot * 4), o_ninl = take(B * 4), o_prev = take(tot * 8), o_next = take(tot * 8),
                 o_st = take(tot), o_outl = take(tot), o_pts = take(tot * 24), o_uv = take(tot * 16), o_T = take(B * 56);
    if (cudaEventCreateWithFlags(&t->up_done, cudaEventDisableTiming) != cudaSuccess || t->mem.reserve(off) != cudaSuccess || t->out_h.reserve(B * sizeof(TrkOut)) != cudaSuccess ||
        t->kf_h.reserve(tot * sizeof(TrkFeat)) != cudaSuccess || t->off_h.reserve((B + 1) * 4) != cudaSuccess) {
        c->err = "tracker: allocation failed";
        svs_i_trk_destroy(c, t);
        return nullptr;
    }
    uint8_t *m = t->mem.as<uint8_t>();
    TrkDev &d = t->d;
    d.B = p.B; d.cap = p.cap;
    d.status = (int32_t *)(m + o_status); d.nfeat = (int32_t *)(m + o_nfeat); d.buf = (int32_t *)(m + o_buf);
    d.last_pose = (double *)(m + o_last); d.rel = (double *)(m + o_rel); d.pose0 = (double *)(m + o_pose0); d.K = (double *)(m + o_K);
    d.xy[0] = (float2 *)(m + o_xy0); d.xy[1] = (float2 *)(m + o_xy1); d.lm[0] = (int64_t *)(m + o_lm0); d.lm[1] = (int64_t *)(m + o_lm1);
    d.pw[0] = (double *)(m + o_pw0); d.pw[1] = (double *)(m + o_pw1);
    d.off = (int32_t *)(m + o_off); d.lm_end = (int32_t *)(m + o_end); d.pt_img = (int32_t *)(m + o_img); d.lm_fidx = (int32_t *)(m + o_fidx);
    d.n_inl = (int32_t *)(m + o_ninl); d.prev_xy = (float2 *)(m + o_prev); d.next_xy = (float2 *)(m + o_next);
    d.st = m + o_st; d.outl = m + o_outl; d.lm_pts = (double *)(m + o_pts); d.lm_uv = (double *)(m + o_uv); d.T_out = (double *)(m + o_T);
    d.out = t->out_h.as<TrkOut>(); d.kf_feats = t->kf_h.as<TrkFeat>();
    memset(d.out, 0, B * sizeof(TrkOut));
    t->h_nfeat.assign(B, 0); t->h_status.assign(B, 0);
    cudaMemsetAsync(m, 0, off, c->stream);
    k_trk_init<<<(p.B + 127) / 128, 128, 0, c->stream>>>(d, p.cam_left.fx_, p.cam_left.fy_, p.cam_left.cx_, p.cam_left.cy_);
    c->launches++;
    if (cudaStreamSynchronize(c->stream) != cudaSuccess) { c->err = "tracker: init failed"; svs_i_trk_destroy(c, t); return nullptr; }
    c->reg_dev.push_back(&t->up_dev); c->reg_pin.push_back(&t->up_h);     // per-step upload scratch: svs_reserve_headroom sizes it
    return t;
}

void svs_i_trk_destroy(svs_ctx *c, svs_tracker *t)
{
    if (!t) return;
    if (c) { cudaSetDevice(c->device); cudaStreamSynchronize(c->stream); c->unregister(&t->up_dev); c->unregister(&t->up_h); }
    t->mem.release(); t->up_dev.release();
    t->out_h.release(); t->kf_h.release(); t->off_h.release(); t->up_h.release();
    if (t->up_done) cudaEventDestroy(t->up_done);
    delete t;
}

const TrkOut *svs_i_trk_out(const svs_tracker *t) { return t->d.out; }
const TrkFeat *svs_i_trk_kf_feats(const svs_tracker *t, int stream) { return t->d.kf_feats + (size_t)stream * t->p.cap; }
int svs_i_trk_is_active(const svs_tracker *t, int stream) { int s = t->h_status[stream]; return s == 1 || s == 2; }

int svs_i_trk_step(svs_ctx *c, svs_tracker *t, svs_frameset *fs, long long *lk_points, long long *pose_edges)
{
    SVS_CUDA(c, cudaSetDevice(c->device));
    const TrkParams &p = t->p;
    const int B = p.B;
    int32_t *off = t->off_h.as<int32_t>();
    int tot = 0, n_act = 0;
    for (int b = 0; b < B; b++) {
        off[b] = tot;
        const int s = t->h_status[b];
        if (s == 1 || s == 2) { tot += t->h_nfeat[b]; n_act++; }
    }
    off[B] = tot;
    if (n_act == 0) return SVS_OK;
    SVS_CUDA(c, cudaMemcpyAsync(t->d.off, off, (size_t)(B + 1) * 4, cudaMemcpyHostToDevice, c->stream));
    const int grid = (B + TRK_WARPS - 1) / TRK_WARPS;
    SVS_KERNEL(c, KID_TRACK_STATE, k_trk_begin<<<grid, TRK_WARPS * 32, 0, c->stream>>>(t->d, p.cam_left));
    SVS_TRY(svs_i_lk(c, fs->Lprev(), fs->Lcur(), t->d.pt_img, reinterpret_cast<const float *>(t->d.prev_xy),
                     reinterpret_cast<float *>(t->d.next_xy), tot, p.lk_win, p.lk_max_iter, p.lk_eps, t->d.st));
    SVS_KERNEL(c, KID_TRACK_STATE, k_trk_mid<<<grid, TRK_WARPS * 32, 0, c->stream>>>(t->d, (float)p.W, (float)p.H));
    SVS_TRY(svs_i_pose_only_lm_dev(c, B, t->d.off, t->d.lm_end, t->d.lm_pts, t->d.lm_uv, t->d.K, t->d.pose0, p.chi2_th, 4, 10, t->d.T_out,
                                   t->d.outl, t->d.n_inl));
    SVS_KERNEL(c, KID_TRACK_STATE, k_trk_end<<<grid, TRK_WARPS * 32, 0, c->stream>>>(t->d, p.num_features_tracking, p.num_features_tracking_bad,
                                                                                   p.num_features_needed_for_keyframe));
    SVS_CUDA(c, svs_i_wait(c));
    long long edges = 0;
    const TrkOut *o = t->d.out;
    for (int b = 0; b < B; b++) {
        const int s = t->h_status[b];
        if (s == 1 || s == 2) { t->h_nfeat[b] = o[b].nfeat; t->h_status[b] = o[b].status; edges += o[b].n_edges; }
    }
    if (lk_points) *lk_points += tot;
    if (pose_edges) *pose_edges += edges;
    return SVS_OK;
}

int svs_i_trk_upload(svs_ctx *c, svs_tracker *t, int n_sel, const TrkUpHdr *hdrs, const TrkUpFeat *feats, long long n_feats)
{
    if (n_sel <= 0) return SVS_OK;
    SVS_CUDA(c, cudaSetDevice(c->device));
    for (int k = 0; k < n_sel; k++) {
        if (hdrs[k].stream < 0 || hdrs[k].stream >= t->p.B) SVS_FAIL(c, SVS_ERR_ARG, "tracker upload: stream out of range");
        if (hdrs[k].n > t->p.cap) SVS_FAIL(c, SVS_ERR_CAPACITY, "tracker: more features in a frame than the device feature table holds (4 * num_features + 512)");
    }
    const size_t hb = align_up((size_t)n_sel * sizeof(TrkUpHdr), 256), fb = (size_t)n_feats * sizeof(TrkUpFeat);
    SVS_CUDA(c, cudaEventSynchronize(t->up_done));      // the previous upload has left the pinned staging buffer
    if (hb + fb + 16 > t->up_dev.cap) SVS_CUDA(c, svs_i_wait(c));   // regrow frees a buffer the scatter kernel may still read
    SVS_CUDA(c, t->up_h.reserve(hb + fb + 16));
    SVS_CUDA(c, t->up_dev.reserve(hb + fb + 16));
    uint8_t *h = t->up_h.as<uint8_t>(), *dv = t->up_dev.as<uint8_t>();
    memcpy(h, hdrs, (size_t)n_sel * sizeof(TrkUpHdr));
    if (fb) memcpy(h + hb, feats, fb);
    SVS_CUDA(c, cudaMemcpyAsync(dv, h, hb + fb, cudaMemcpyHostToDevice, c->stream));
    SVS_KERNEL(c, KID_TRACK_STATE, k_trk_scatter<<<n_sel, 128, 0, c->stream>>>(t->d, reinterpret_cast<const TrkUpHdr *>(dv),
                                                                                reinterpret_cast<const TrkUpFeat *>(dv + hb)));
    for (int k = 0; k < n_sel; k++) { t->h_nfeat[hdrs[k].stream] = hdrs[k].n; t->h_status[hdrs[k].stream] = hdrs[k].status; }
    SVS_CUDA(c, cudaEventRecord(t->up_done, c->stream));
    return SVS_OK;
}

int svs_i_trk_fetch(svs_ctx *c, svs_tracker *t, int stream, const TrkFeat **feats, int *n)
{
    if (stream < 0 || stream >= t->p.B) return SVS_ERR_ARG;
    SVS_CUDA(c, cudaSetDevice(c->device));
    SVS_KERNEL(c, KID_TRACK_STATE, k_trk_export<<<1, 128, 0, c->stream>>>(t->d, stream));
    SVS_CUDA(c, svs_i_wait(c));
    *feats = t->d.kf_feats + (size_t)stream * t->p.cap;
    *n = t->d.out[stream].nfeat;
    return SVS_OK;
}
