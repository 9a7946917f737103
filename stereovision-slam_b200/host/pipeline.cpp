// slam::StreamBatch — steps n independent stereo streams in lock-step: the host mirrors of the reference
// classes (slam.h) do the per-stream bookkeeping (OpenMP over streams), and every third-party seam of
// Frontend::AddFrame / Backend::Optimize becomes ONE batched call into the C ABI (include/svslam.h) per step.
// Exposed through the svs_slam_* entry points (C ABI) for tests, bench.py and a C++ caller.
#include <omp.h>
#include <chrono>
#include <cstring>
#include <new>
#include <string>
#include <vector>
#include "../../include/svslam.h"
#include "slam.h"
#include "../csrc/track.h"

namespace slam {

struct Stream {
    Map::Ptr map;
    std::shared_ptr<Frontend> frontend;
    std::shared_ptr<Backend> backend;
    LkRequest lk;
    PoseRequest pose;
    DetectRequest det;
    TriRequest tri;
    BaRequest ba;
    bool ran_track = false, ran_detect = false, ran_backend = false, is_kf = false;
    bool on_host = true;          // the current frame exists in the host Frontend (always, without device tracking)
    double out_pose[7] = {0, 0, 0, 1, 0, 0, 0};
};

class StreamBatch {
public:
    StreamBatch(svs_ctx *ctx, int n_streams, int in_w, int in_h, int half, const Config &cfg, const double K[4],
                double baseline)
        : ctx_(ctx), cfg_(cfg)
    {
        fs_ = svs_frameset_create(ctx, n_streams, in_w, in_h, half, cfg.lk_win, cfg.lk_max_level);
        if (!fs_) return;
        svs_frameset_size(fs_, &W_, &H_, nullptr);
        cam_left_ = std::make_shared<Camera>(K[0], K[1], K[2], K[3], 0.0, SE3());
        cam_right_ = std::make_shared<Camera>(K[0], K[1], K[2], K[3], baseline, SE3::fromTranslation(Vec3(-baseline, 0, 0)));
        baseline_ = baseline;
        streams_.resize(n_streams);
        for (Stream &s : streams_) {
            s.map = std::make_shared<Map>(cfg.num_active_keyframes);
            s.frontend = std::make_shared<Frontend>(cfg);
            s.frontend->SetMap(s.map);
            s.frontend->SetCameras(cam_left_, cam_right_);
            if (cfg.backend_on) {
                s.backend = std::make_shared<Backend>(cfg);
                s.backend->SetMap(s.map);
                s.backend->SetCameras(cam_left_, cam_right_);
                s.frontend->SetBackend(s.backend);
            }
        }
        if (cfg.device_tracking) {
            TrkParams p;
            p.B = n_streams; p.W = W_; p.H = H_;
            p.cap = (4 * cfg.num_features + 512 + 31) / 32 * 32;      // tracked + newly detected left features of one frame
            p.num_features_tracking = cfg.num_features_tracking; p.num_features_tracking_bad = cfg.num_features_tracking_bad;
            p.num_features_needed_for_keyframe = cfg.num_features_needed_for_keyframe;
            p.lk_win = cfg.lk_win; p.lk_max_iter = cfg.lk_max_iter; p.lk_eps = cfg.lk_eps; p.chi2_th = 5.991;
            p.cam_left = *cam_left_;
            trk_ = svs_i_trk_create(ctx, p);
            if (!trk_) { svs_frameset_destroy(ctx, fs_); fs_ = nullptr; }
        }
    }
    ~StreamBatch() { if (trk_) svs_i_trk_destroy(ctx_, trk_); if (fs_) svs_frameset_destroy(ctx_, fs_); }
    bool ok() const { return fs_ != nullptr; }
    svs_tracker *tracker() { return trk_; }
    svs_ctx *ctx() { return ctx_; }
    int n() const { return (int)streams_.size(); }
    Stream &stream(int i) { return streams_[i]; }
    svs_frameset *frameset() { return fs_; }

    int step(const uint8_t *const *left, const uint8_t *const *right, size_t row_stride, int on_device);
    int track_host(double &t0);      // Track() up to the keyframe decision, every seam a host round trip
    int track_device(double &t0);    // the same on the device-resident state (csrc/track.cu)
    int upload_states();             // hand the streams the host touched in this step back to the device
    int run_lk(int pair, bool (*sel)(const Stream &));

    void set_threads(int n) { threads_ = n > 0 ? n : 1; ctx_->host_threads = threads_; }
    void hint_next(const uint8_t *const *l, const uint8_t *const *r)
    {
        next_left_.assign(l, l + n()); next_right_.assign(r, r + n());
        has_next_ = true;
    }
    double t_phase[8] = {0, 0, 0, 0, 0, 0, 0, 0};   // push, track-lk, pose, detect, right-lk, triangulate, ba, host
    // host bookkeeping by section: begin+prepare track | finish track+prepare pose | finish pose+prepare detect |
    // finish detect+prepare right | finish right+prepare triangulate | finish triangulate+prepare BA | finish BA+end
    double t_host[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    long long frames = 0, keyframes = 0, ba_problems = 0, ba_iterations = 0, ba_trials = 0, ba_edges = 0, ba_lms = 0, ba_kfs = 0, lk_points = 0, pose_edges = 0, right_images = 0;

private:
    svs_ctx *ctx_;
    svs_frameset *fs_ = nullptr;
    svs_tracker *trk_ = nullptr;
    std::vector<TrkUpHdr> up_hdr_;
    std::vector<TrkUpFeat> up_feat_;
    Config cfg_;
    int W_ = 0, H_ = 0;
    int threads_ = omp_get_max_threads();
    double baseline_ = 0;
    Camera::Ptr cam_left_, cam_right_;
    std::vector<Stream> streams_;
    std::vector<const uint8_t *> next_left_, next_right_, rptr_;
    bool has_next_ = false;
    // gather buffers
    std::vector<int32_t> off_, off2_, off3_, ids_, det_ids_, host_ids_;
    std::vector<float> f0_, f1_, f2_;
    std::vector<uint8_t> u0_;
    std::vector<double> d0_, d1_, d2_, d3_, d4_;
    std::vector<int32_t> i0_, i1_, i2_;
    std::vector<svs_ba_stats> bast_;
};

static inline double now_s()
{
    return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

int StreamBatch::run_lk(int pair, bool (*sel)(const Stream &))
{
    const int B = n();
    off_.assign(B + 1, 0);
    for (int b = 0; b < B; b++) off_[b + 1] = off_[b] + (sel(streams_[b]) ? (int)streams_[b].lk.status.size() : 0);
    int tot = off_[B];
    if (tot == 0) return 0;
    lk_points += tot;
    f0_.resize((size_t)2 * tot); f1_.resize((size_t)2 * tot); u0_.resize(tot);
    for (int b = 0; b < B; b++) {
        if (!sel(streams_[b])) continue;
        const LkRequest &q = streams_[b].lk;
        if (q.status.empty()) continue;
        memcpy(&f0_[2 * (size_t)off_[b]], q.prev_xy.data(), q.prev_xy.size() * 4);
        memcpy(&f1_[2 * (size_t)off_[b]], q.next_xy.data(), q.next_xy.size() * 4);
    }
    int r = svs_lk_track_batch(ctx_, fs_, pair, off_.data(), f0_.data(), f1_.data(), cfg_.lk_max_iter, cfg_.lk_eps, u0_.data());
    if (r) return r;
    for (int b = 0; b < B; b++) {
        if (!sel(streams_[b])) continue;
        LkRequest &q = streams_[b].lk;
        if (q.status.empty()) continue;
        memcpy(q.next_xy.data(), &f1_[2 * (size_t)off_[b]], q.next_xy.size() * 4);
        memcpy(q.status.data(), &u0_[off_[b]], q.status.size());
    }
    return 0;
}

int StreamBatch::track_host(double &t0)
{
    const int B = n();
    double t1;
    int rc;
#pragma omp parallel for schedule(dynamic, 16) num_threads(threads_)
    for (int b = 0; b < B; b++) {
        Stream &s = streams_[b];
        s.ran_track = s.ran_detect = s.ran_backend = s.is_kf = false;
        Frame::Ptr f = s.frontend->CreateFrame();
        s.frontend->begin_AddFrame(f, W_, H_);
        s.lk.prev_xy.clear(); s.lk.next_xy.clear(); s.lk.status.clear();
        if (s.frontend->wants_track()) { s.frontend->prepare_TrackLastFrame(s.lk); s.ran_track = true; }
    }
    // ---------------- TrackLastFrame: LK previous-left -> current-left (src/frontend.cpp:353-357)
    t1 = now_s(); t_phase[7] += t1 - t0; t_host[0] += t1 - t0; t0 = t1;
    if ((rc = run_lk(0, [](const Stream &s) { return s.ran_track; }))) return rc;
    t1 = now_s(); t_phase[1] += t1 - t0; t0 = t1;

    // ---------------- EstimateCurrentPose: pose-only LM (src/frontend.cpp:408-527)
#pragma omp parallel for schedule(dynamic, 16) num_threads(threads_)
    for (int b = 0; b < B; b++) {
        Stream &s = streams_[b];
        if (!s.ran_track) continue;
        s.frontend->finish_TrackLastFrame(s.lk);
        s.frontend->prepare_EstimateCurrentPose(s.pose);
    }
    {
        ids_.clear();
        for (int b = 0; b < B; b++) if (streams_[b].ran_track) ids_.push_back(b);
        int np = (int)ids_.size();
        t1 = now_s(); t_phase[7] += t1 - t0; t_host[1] += t1 - t0; t0 = t1;
        if (np > 0) {
            off_.assign(np + 1, 0);
            for (int k = 0; k < np; k++) off_[k + 1] = off_[k] + (int)streams_[ids_[k]].pose.feat_index.size();
            int M = off_[np];
            pose_edges += M;
            d0_.resize((size_t)3 * M + 1); d1_.resize((size_t)2 * M + 1); d2_.resize((size_t)4 * np); d3_.resize((size_t)7 * np);
            d4_.resize((size_t)7 * np); u0_.resize(M + 1); i0_.resize(np);
            for (int k = 0; k < np; k++) {
                const PoseRequest &q = streams_[ids_[k]].pose;
                if (!q.feat_index.empty()) {
                    memcpy(&d0_[3 * (size_t)off_[k]], q.pts_w.data(), q.pts_w.size() * 8);
                    memcpy(&d1_[2 * (size_t)off_[k]], q.uv.data(), q.uv.size() * 8);
                }
                memcpy(&d2_[4 * (size_t)k], q.K, 32);
                memcpy(&d3_[7 * (size_t)k], q.T0, 56);
            }
            rc = svs_pose_only_lm(ctx_, np, off_.data(), d0_.data(), d1_.data(), d2_.data(), d3_.data(), 5.991, 4, 10, d4_.data(),
                                  u0_.data(), i0_.data(), nullptr);
            if (rc) return rc;
            for (int k = 0; k < np; k++) {
                PoseRequest &q = streams_[ids_[k]].pose;
                memcpy(q.T, &d4_[7 * (size_t)k], 56);
                if (!q.feat_index.empty()) memcpy(q.outlier.data(), &u0_[off_[k]], q.outlier.size());
                q.n_inlier = i0_[k];
            }
        }
        t1 = now_s(); t_phase[2] += t1 - t0; t0 = t1;
    }
#pragma omp parallel for schedule(dynamic, 16) num_threads(threads_)
    for (int b = 0; b < B; b++) {
        Stream &s = streams_[b];
        if (s.ran_track) s.frontend->finish_EstimateCurrentPose(s.pose);
        if (s.frontend->wants_detect()) { s.frontend->prepare_DetectFeatures(s.det); s.ran_detect = true; }
    }
    return 0;
}

int StreamBatch::track_device(double &t0)
{
    const int B = n();
    double t1;
    int rc = svs_i_trk_step(ctx_, trk_, fs_, &lk_points, &pose_edges);
    if (rc) return rc;
    t1 = now_s(); t_phase[1] += t1 - t0; t0 = t1;
    const TrkOut *o = svs_i_trk_out(trk_);
#pragma omp parallel for schedule(dynamic, 16) num_threads(threads_)
    for (int b = 0; b < B; b++) {
        Stream &s = streams_[b];
        s.ran_track = s.ran_detect = s.ran_backend = s.is_kf = false;
        s.on_host = false;
        const FrontendStatus st = s.frontend->GetStatus();
        if (st == FrontendStatus::INITING) {                       // StereoInit runs on the host classes
            Frame::Ptr f = s.frontend->CreateFrame();
            s.frontend->begin_AddFrame(f, W_, H_);
            s.on_host = true;
        } else if (st == FrontendStatus::TRACKING_GOOD || st == FrontendStatus::TRACKING_BAD) {
            const TrkOut &r = o[b];
            s.ran_track = true;
            if (r.need_kf) {                                       // InsertKeyframe: the stream comes back to the host
                s.frontend->adopt_tracked_keyframe(r.pose, r.last_pose, reinterpret_cast<const float *>(svs_i_trk_kf_feats(trk_, b)), r.nfeat,
                                                   r.status, r.inliers, W_, H_);
                s.on_host = true;
            } else {
                s.frontend->note_tracked_frame(r.status, r.inliers);
                memcpy(s.out_pose, r.pose, 56);
            }
        } else {                                                    // LOST: Reset() is "not implemented" in the reference
            s.frontend->skip_frame();
            const double I[7] = {0, 0, 0, 1, 0, 0, 0};
            memcpy(s.out_pose, I, 56);
        }
        if (s.on_host && s.frontend->wants_detect()) { s.frontend->prepare_DetectFeatures(s.det); s.ran_detect = true; }
    }
    return 0;
}

// Every stream the host classes touched in this step (keyframe inserted, initialisation tried) goes back to the device:
// its current frame's left features with the landmark positions (refined by BA), its pose and the relative motion.
int StreamBatch::upload_states()
{
    const int B = n();
    ids_.clear();
    for (int b = 0; b < B; b++) if (streams_[b].on_host) ids_.push_back(b);
    const int ns = (int)ids_.size();
    if (ns == 0) return 0;
    up_hdr_.resize(ns);
    long long tot = 0;
    for (int k = 0; k < ns; k++) {
        Stream &s = streams_[ids_[k]];
        TrkUpHdr &h = up_hdr_[k];
        const Frame &f = *s.frontend->current_frame_;
        h.stream = ids_[k]; h.n = (int)f.feature_left_.size(); h.status = (int)s.frontend->GetStatus(); h.pad_ = 0; h.pad2_ = 0;
        memcpy(h.pose, f.pose_.d, 56);
        memcpy(h.rel, s.frontend->relative_motion().d, 56);
        h.feat_off = tot;
        tot += h.n;
    }
    up_feat_.resize((size_t)tot);
#pragma omp parallel for schedule(dynamic, 1) num_threads(threads_)
    for (int k = 0; k < ns; k++) {
        Stream &s = streams_[ids_[k]];
        const Frame &f = *s.frontend->current_frame_;
        Map *map = s.frontend->map();
        TrkUpFeat *o = up_feat_.data() + up_hdr_[k].feat_off;
        for (size_t i = 0; i < f.feature_left_.size(); i++) {
            const Feature &ft = f.feature_left_[i];
            o[i].x = ft.x; o[i].y = ft.y; o[i].lm = ft.map_point_;
            o[i].pw[0] = o[i].pw[1] = o[i].pw[2] = 0.0;
            if (const MapPoint *mp = map->GetMapPoint(ft.map_point_)) { o[i].pw[0] = mp->pos_.x; o[i].pw[1] = mp->pos_.y; o[i].pw[2] = mp->pos_.z; }
        }
    }
    return svs_i_trk_upload(ctx_, trk_, ns, up_hdr_.data(), up_feat_.data(), tot);
}

int StreamBatch::step(const uint8_t *const *left, const uint8_t *const *right, size_t row_stride, int on_device)
{
    const int B = n();
    double t0 = now_s(), t1;
    const bool lazy = cfg_.lazy_right_ingest != 0;
    int rc = svs_frameset_push_ptrs(ctx_, fs_, left, lazy ? nullptr : right, row_stride, on_device);
    if (rc) return rc;
    if (has_next_) {   // double buffering: the next pair's PCIe transfer / resize / pyramids overlap this step's kernels
        has_next_ = false;
        if ((rc = svs_frameset_prefetch_ptrs(ctx_, fs_, next_left_.data(), lazy ? nullptr : next_right_.data(), row_stride, on_device))) return rc;
    }
    t1 = now_s(); t_phase[0] += t1 - t0; t0 = t1;

    if ((rc = trk_ ? track_device(t0) : track_host(t0))) return rc;
    // ---------------- DetectFeatures: GFTT with the tracked-feature mask (src/frontend.cpp:42-51)
    {
        ids_.clear();
        for (int b = 0; b < B; b++) if (streams_[b].ran_detect) ids_.push_back(b);
        int ns = (int)ids_.size();
        t1 = now_s(); t_phase[7] += t1 - t0; t_host[2] += t1 - t0; t0 = t1;
        if (ns > 0) {
            const int mc = cfg_.num_features;
            off_.assign(ns + 1, 0);
            for (int k = 0; k < ns; k++) off_[k + 1] = off_[k] + (int)streams_[ids_[k]].det.occupied_xy.size() / 2;
            f0_.resize((size_t)2 * off_[ns] + 2); f1_.resize((size_t)2 * mc * ns); f2_.resize((size_t)mc * ns); i0_.resize(ns);
            for (int k = 0; k < ns; k++) {
                const DetectRequest &q = streams_[ids_[k]].det;
                if (!q.occupied_xy.empty()) memcpy(&f0_[2 * (size_t)off_[k]], q.occupied_xy.data(), q.occupied_xy.size() * 4);
            }
            rc = svs_gftt_detect_batch(ctx_, fs_, ids_.data(), ns, off_.data(), f0_.data(), mc, cfg_.gftt_quality, cfg_.gftt_min_distance,
                                       cfg_.oracle_simd_granule, f1_.data(), f2_.data(), i0_.data());
            if (rc) return rc;
            for (int k = 0; k < ns; k++) {
                DetectRequest &q = streams_[ids_[k]].det;
                q.out_n = i0_[k];
                memcpy(q.out_xy.data(), &f1_[(size_t)2 * mc * k], (size_t)2 * q.out_n * 4);
                memcpy(q.out_resp.data(), &f2_[(size_t)mc * k], (size_t)q.out_n * 4);
            }
        }
        t1 = now_s(); t_phase[3] += t1 - t0; t0 = t1;
    }
    // ---------------- FindFeaturesInRight: LK current-left -> current-right (src/frontend.cpp:105-109)
    // the keyframe streams are a few per cent of the batch: the host sections below walk their id list, one stream per task
    det_ids_.clear();
    for (int b = 0; b < B; b++) if (streams_[b].ran_detect) det_ids_.push_back(b);
    const int n_det = (int)det_ids_.size();
#pragma omp parallel for schedule(dynamic, 1) num_threads(threads_) if (n_det > 1)
    for (int k = 0; k < n_det; k++) {
        Stream &s = streams_[det_ids_[k]];
        s.frontend->finish_DetectFeatures(s.det);
        s.frontend->prepare_FindFeaturesInRight(s.lk);
    }
    t1 = now_s(); t_phase[7] += t1 - t0; t_host[3] += t1 - t0; t0 = t1;
    if (lazy) {   // the right image is needed now, and only by the streams that are inserting a keyframe (or initialising)
        ids_.clear(); rptr_.clear();
        for (int b = 0; b < B; b++) if (streams_[b].ran_detect) { ids_.push_back(b); rptr_.push_back(right[b]); }
        right_images += (long long)ids_.size();
        if ((rc = svs_frameset_fetch_right_ptrs(ctx_, fs_, ids_.data(), (int)ids_.size(), rptr_.data(), row_stride, on_device))) return rc;
    } else {
        right_images += B;
    }
    if ((rc = run_lk(1, [](const Stream &s) { return s.ran_detect; }))) return rc;
    t1 = now_s(); t_phase[4] += t1 - t0; t0 = t1;
    // ---------------- triangulation of new landmarks (src/frontend.cpp:174, :286)
#pragma omp parallel for schedule(dynamic, 1) num_threads(threads_) if (n_det > 1)
    for (int k = 0; k < n_det; k++) {
        Stream &s = streams_[det_ids_[k]];
        s.frontend->finish_FindFeaturesInRight(s.lk);
        s.frontend->prepare_Triangulate(s.tri);
    }
    {
        off_.assign(B + 1, 0);
        for (int b = 0; b < B; b++) off_[b + 1] = off_[b] + (streams_[b].ran_detect ? (int)streams_[b].tri.feat_index.size() : 0);
        int tot = off_[B];
        t1 = now_s(); t_phase[7] += t1 - t0; t_host[4] += t1 - t0; t0 = t1;
        if (tot > 0) {
            f0_.resize((size_t)2 * tot); f1_.resize((size_t)2 * tot); d0_.resize((size_t)3 * tot); u0_.resize(tot);
            for (int b = 0; b < B; b++) {
                const TriRequest &q = streams_[b].tri;
                if (!streams_[b].ran_detect || q.feat_index.empty()) continue;
                memcpy(&f0_[2 * (size_t)off_[b]], q.left_xy.data(), q.left_xy.size() * 4);
                memcpy(&f1_[2 * (size_t)off_[b]], q.right_xy.data(), q.right_xy.size() * 4);
            }
            double Kl[4], Kr[4];
            cam_left_->K(Kl); cam_right_->K(Kr);
            rc = svs_triangulate(ctx_, f0_.data(), f1_.data(), tot, Kl, Kr, baseline_, d0_.data(), u0_.data());
            if (rc) return rc;
            for (int b = 0; b < B; b++) {
                TriRequest &q = streams_[b].tri;
                if (!streams_[b].ran_detect || q.feat_index.empty()) continue;
                memcpy(q.xyz.data(), &d0_[3 * (size_t)off_[b]], q.xyz.size() * 8);
                memcpy(q.ok.data(), &u0_[off_[b]], q.ok.size());
            }
        }
        t1 = now_s(); t_phase[5] += t1 - t0; t0 = t1;
    }
    // ---------------- Backend::UpdateMap -> Optimize, synchronous schedule (src/backend.cpp:9-248)
#pragma omp parallel for schedule(dynamic, 1) num_threads(threads_) if (n_det > 1)
    for (int k = 0; k < n_det; k++) {
        Stream &s = streams_[det_ids_[k]];
        s.is_kf = s.frontend->finish_Triangulate(s.tri) != 0;
        if (s.is_kf && s.frontend->wants_backend()) s.ran_backend = s.backend->prepare_Optimize(s.ba);
    }
    {
        ids_.clear();
        for (int b = 0; b < B; b++) if (streams_[b].ran_backend) ids_.push_back(b);
        int np = (int)ids_.size();
        t1 = now_s(); t_phase[7] += t1 - t0; t_host[5] += t1 - t0; t0 = t1;
        if (np > 0) {
            off_.assign(np + 1, 0); off2_.assign(np + 1, 0); off3_.assign(np + 1, 0);
            for (int k = 0; k < np; k++) {
                const BaRequest &q = streams_[ids_[k]].ba;
                off_[k + 1] = off_[k] + (int)q.kf_ids.size();
                off2_[k + 1] = off2_[k] + (int)q.lm_ids.size();
                off3_[k + 1] = off3_[k] + (int)q.edge_kf.size();
            }
            int sN = off_[np], sL = off2_[np], sE = off3_[np];
            d0_.resize((size_t)7 * sN); d1_.resize((size_t)3 * sL + 1); d2_.resize((size_t)2 * sE); d3_.resize(sE);
            i0_.resize(sE); i1_.resize(sE); u0_.resize(sE); bast_.resize(np);
#pragma omp parallel for schedule(dynamic, 1) num_threads(threads_)
            for (int k = 0; k < np; k++) {
                const BaRequest &q = streams_[ids_[k]].ba;
                memcpy(&d0_[7 * (size_t)off_[k]], q.poses.data(), q.poses.size() * 8);
                if (!q.lms.empty()) memcpy(&d1_[3 * (size_t)off2_[k]], q.lms.data(), q.lms.size() * 8);
                memcpy(&d2_[2 * (size_t)off3_[k]], q.edge_uv.data(), q.edge_uv.size() * 8);
                memcpy(&i0_[off3_[k]], q.edge_kf.data(), q.edge_kf.size() * 4);
                memcpy(&i1_[off3_[k]], q.edge_lm.data(), q.edge_lm.size() * 4);
                memcpy(&u0_[off3_[k]], q.edge_cam.data(), q.edge_cam.size());
            }
            double Kl[4], Kr[4];
            cam_left_->K(Kl); cam_right_->K(Kr);
            SE3 el = cam_left_->pose(), er = cam_right_->pose();
            rc = svs_ba_optimize(ctx_, np, off_.data(), d0_.data(), off2_.data(), d1_.data(), off3_.data(), i0_.data(), i1_.data(),
                                 u0_.data(), d2_.data(), Kl, Kr, el.d, er.d, cfg_.chi2_th, cfg_.ba_max_iter, cfg_.ba_jacobian_mode,
                                 d3_.data(), bast_.data());
            if (rc) return rc;
#pragma omp parallel for schedule(dynamic, 1) num_threads(threads_)
            for (int k = 0; k < np; k++) {
                BaRequest &q = streams_[ids_[k]].ba;
                memcpy(q.poses.data(), &d0_[7 * (size_t)off_[k]], q.poses.size() * 8);
                if (!q.lms.empty()) memcpy(q.lms.data(), &d1_[3 * (size_t)off2_[k]], q.lms.size() * 8);
                memcpy(q.chi2.data(), &d3_[off3_[k]], q.chi2.size() * 8);
            }
            for (int k = 0; k < np; k++) {
                const BaRequest &q = streams_[ids_[k]].ba;
                ba_iterations += bast_[k].iterations; ba_trials += bast_[k].trials; ba_edges += (long long)q.edge_kf.size();
                ba_lms += (long long)q.lm_ids.size(); ba_kfs += (long long)q.kf_ids.size();
            }
            ba_problems += np;
        }
        t1 = now_s(); t_phase[6] += t1 - t0; t0 = t1;
    }
    long long nkf = 0;
    host_ids_.clear();                      // device-resident tracking: the other streams' frames never left the GPU
    for (int b = 0; b < B; b++) if (streams_[b].on_host) host_ids_.push_back(b);
    const int n_host = (int)host_ids_.size();
#pragma omp parallel for schedule(dynamic, 1) reduction(+ : nkf) num_threads(threads_) if (n_host > 1)
    for (int k = 0; k < n_host; k++) {
        Stream &s = streams_[host_ids_[k]];
        if (s.ran_backend) s.backend->finish_Optimize(s.ba);
        s.frontend->end_AddFrame();
        memcpy(s.out_pose, s.frontend->current_frame_->pose_.d, 56);
        nkf += s.is_kf ? 1 : 0;
    }
    keyframes += nkf;
    frames += B;
    if (trk_ && (rc = upload_states())) return rc;
    t1 = now_s(); t_phase[7] += t1 - t0; t_host[6] += t1 - t0;
    return 0;
}

}  // namespace slam

// ================================================================== C ABI
struct svs_slam {
    slam::StreamBatch *batch = nullptr;
    svs_ctx *ctx = nullptr;
};

extern "C" {

svs_slam *svs_slam_create(svs_ctx *ctx, int n_streams, int in_w, int in_h, int half, const svs_slam_config *c, const double K[4],
                          double baseline)
{
    if (!ctx || !c || !K || n_streams <= 0) return nullptr;
    slam::Config cfg;
    cfg.num_features = c->num_features; cfg.num_features_init = c->num_features_init;
    cfg.num_features_tracking = c->num_features_tracking; cfg.num_features_tracking_bad = c->num_features_tracking_bad;
    cfg.num_features_needed_for_keyframe = c->num_features_needed_for_keyframe;
    cfg.max_triangulation_depth = c->max_triangulation_depth; cfg.num_active_keyframes = c->num_active_keyframes;
    cfg.backend_on = c->backend_on; cfg.chi2_th = c->chi2_th; cfg.gftt_quality = c->gftt_quality;
    cfg.gftt_min_distance = c->gftt_min_distance; cfg.lk_win = c->lk_win; cfg.lk_max_level = c->lk_max_level;
    cfg.lk_max_iter = c->lk_max_iter; cfg.lk_eps = c->lk_eps; cfg.ba_max_iter = c->ba_max_iter;
    cfg.ba_jacobian_mode = c->ba_jacobian_mode; cfg.oracle_simd_granule = c->oracle_simd_granule;
    cfg.lazy_right_ingest = c->lazy_right_ingest;
    cfg.device_tracking = c->device_tracking;
    svs_slam *s = new (std::nothrow) svs_slam();
    if (!s) return nullptr;
    s->ctx = ctx;
    s->batch = new (std::nothrow) slam::StreamBatch(ctx, n_streams, in_w, in_h, half, cfg, K, baseline);
    if (!s->batch || !s->batch->ok()) { delete s->batch; delete s; return nullptr; }
    return s;
}

void svs_slam_destroy(svs_slam *s)
{
    if (!s) return;
    delete s->batch;
    delete s;
}

void svs_slam_default_config(svs_slam_config *c)
{
    slam::Config d;
    c->num_features = d.num_features; c->num_features_init = d.num_features_init;
    c->num_features_tracking = d.num_features_tracking; c->num_features_tracking_bad = d.num_features_tracking_bad;
    c->num_features_needed_for_keyframe = d.num_features_needed_for_keyframe;
    c->max_triangulation_depth = d.max_triangulation_depth; c->num_active_keyframes = d.num_active_keyframes;
    c->backend_on = d.backend_on; c->chi2_th = d.chi2_th; c->gftt_quality = d.gftt_quality;
    c->gftt_min_distance = d.gftt_min_distance; c->lk_win = d.lk_win; c->lk_max_level = d.lk_max_level;
    c->lk_max_iter = d.lk_max_iter; c->lk_eps = d.lk_eps; c->ba_max_iter = d.ba_max_iter;
    c->ba_jacobian_mode = d.ba_jacobian_mode; c->oracle_simd_granule = d.oracle_simd_granule;
    c->lazy_right_ingest = d.lazy_right_ingest; c->device_tracking = d.device_tracking;
}

int svs_slam_add_frames(svs_slam *s, const uint8_t *const *left, const uint8_t *const *right, size_t row_stride, int on_device,
                        double *poses_out, int32_t *status_out, int32_t *keyframe_out, int32_t *inliers_out)
{
    if (!s || !left || !right) return SVS_ERR_ARG;
    int rc = s->batch->step(left, right, row_stride, on_device);
    if (rc) return rc;
    for (int b = 0; b < s->batch->n(); b++) {
        slam::Stream &st = s->batch->stream(b);
        if (poses_out) memcpy(poses_out + 7 * (size_t)b, st.out_pose, 56);
        if (status_out) status_out[b] = (int)st.frontend->GetStatus();
        if (keyframe_out) keyframe_out[b] = st.is_kf ? 1 : 0;
        if (inliers_out) inliers_out[b] = st.frontend->tracking_inliers_;
    }
    return SVS_OK;
}

int svs_slam_get_features(svs_slam *s, int stream, int right, float *xy, int64_t *map_point_ids, uint8_t *valid, int cap, int *n)
{
    if (!s || stream < 0 || stream >= s->batch->n() || !n) return SVS_ERR_ARG;
    slam::Stream &st = s->batch->stream(stream);
    if (!st.on_host) {      // device-resident tracking: the frame lives on the GPU (left features only; no right features
                            // exist outside keyframes)
        *n = 0;
        if (right || st.frontend->GetStatus() == slam::FrontendStatus::LOST) return SVS_OK;
        const TrkFeat *ft = nullptr;
        int rc = svs_i_trk_fetch(s->batch->ctx(), s->batch->tracker(), stream, &ft, n);
        if (rc) return rc;
        for (int i = 0; i < *n && i < cap; i++) {
            if (xy) { xy[2 * i] = ft[i].x; xy[2 * i + 1] = ft[i].y; }
            if (map_point_ids) map_point_ids[i] = ft[i].lm;
            if (valid) valid[i] = 1;
        }
        return SVS_OK;
    }
    slam::Frame::Ptr f = st.frontend->current_frame_;
    if (!f) { *n = 0; return SVS_OK; }
    const std::vector<slam::Feature> &v = right ? f->feature_right_ : f->feature_left_;
    *n = (int)v.size();
    for (int i = 0; i < (int)v.size() && i < cap; i++) {
        if (xy) { xy[2 * i] = v[i].x; xy[2 * i + 1] = v[i].y; }
        if (map_point_ids) map_point_ids[i] = v[i].map_point_;
        if (valid) valid[i] = v[i].valid ? 1 : 0;
    }
    return SVS_OK;
}

int svs_slam_get_keyframes(svs_slam *s, int stream, int active_only, int64_t *kf_ids, int64_t *frame_ids, double *poses, int cap, int *n)
{
    if (!s || stream < 0 || stream >= s->batch->n() || !n) return SVS_ERR_ARG;
    slam::Map &m = *s->batch->stream(stream).map;
    const slam::Map::KeyframesType &k = active_only ? m.GetActiveKeyFrames() : m.GetAllKeyFrames();
    *n = (int)k.size();
    int i = 0;
    for (const auto &kv : k) {
        if (i >= cap) break;
        if (kf_ids) kf_ids[i] = (int64_t)kv.first;
        if (frame_ids) frame_ids[i] = (int64_t)kv.second->id_;
        if (poses) memcpy(poses + 7 * (size_t)i, kv.second->pose_.d, 56);
        i++;
    }
    return SVS_OK;
}

int svs_slam_get_landmarks(svs_slam *s, int stream, int active_only, int64_t *ids, double *xyz, int32_t *observed_times, int cap, int *n)
{
    if (!s || stream < 0 || stream >= s->batch->n() || !n) return SVS_ERR_ARG;
    slam::Map &m = *s->batch->stream(stream).map;
    const slam::Map::LandmarksType &l = active_only ? m.GetActiveMapPoints() : m.GetAllMapPoints();
    *n = (int)l.size();
    int i = 0;
    for (const auto &kv : l) {
        if (i >= cap) break;
        if (ids) ids[i] = (int64_t)kv.first;
        if (xyz) { xyz[3 * i] = kv.second->pos_.x; xyz[3 * i + 1] = kv.second->pos_.y; xyz[3 * i + 2] = kv.second->pos_.z; }
        if (observed_times) observed_times[i] = kv.second->observed_times_;
        i++;
    }
    return SVS_OK;
}

int svs_slam_get_counters(svs_slam *s, double *phase_seconds /* 8 */, long long *counters /* 12 */)
{
    if (!s) return SVS_ERR_ARG;
    slam::StreamBatch &b = *s->batch;
    if (phase_seconds) memcpy(phase_seconds, b.t_phase, sizeof(b.t_phase));
    if (counters) {
        counters[0] = b.frames; counters[1] = b.keyframes; counters[2] = b.ba_problems;
        counters[3] = b.ba_iterations; counters[4] = b.ba_trials; counters[5] = b.ba_edges;
        counters[6] = b.ba_lms; counters[7] = b.ba_kfs; counters[8] = b.lk_points; counters[9] = b.pose_edges;
        counters[10] = svs_frameset_h2d_bytes(b.frameset()); counters[11] = b.right_images;
    }
    return SVS_OK;
}

int svs_slam_hint_next(svs_slam *s, const uint8_t *const *next_left, const uint8_t *const *next_right)
{
    if (!s || !next_left || !next_right) return SVS_ERR_ARG;
    s->batch->hint_next(next_left, next_right);
    return SVS_OK;
}

int svs_slam_get_host_seconds(svs_slam *s, double *host_seconds /* 8 */)
{
    if (!s || !host_seconds) return SVS_ERR_ARG;
    memcpy(host_seconds, s->batch->t_host, sizeof(s->batch->t_host));
    return SVS_OK;
}

svs_frameset *svs_slam_frameset(svs_slam *s) { return s ? s->batch->frameset() : nullptr; }

int svs_slam_set_threads(svs_slam *s, int n)
{
    if (!s) return SVS_ERR_ARG;
    s->batch->set_threads(n);
    return SVS_OK;
}

}  // extern "C"
