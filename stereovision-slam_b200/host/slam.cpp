// Host-side mirror of the reference's Frontend / Backend / Map logic (see slam.h for the file:line map).
// Pure host bookkeeping: every piece of arithmetic the reference delegates to OpenCV / g2o is NOT here —
// it is requested through the prepare_* / finish_* pairs and executed on the GPU by slam::StreamBatch.
#include "slam.h"
#include <algorithm>
#include <cfloat>

namespace slam {

// ------------------------------------------------------------------ SE3 (Sophus::SE3d formulas)
SE3 SE3::exp(const double *a)
{
    const double *w = a + 3;
    double th2 = w[0] * w[0] + w[1] * w[1] + w[2] * w[2], th = std::sqrt(th2), imag, real;
    if (th2 < 1e-10 * 1e-10) { double th4 = th2 * th2; imag = 0.5 - th2 / 48.0 + th4 / 3840.0; real = 1.0 - th2 / 8.0 + th4 / 384.0; }
    else { double h = 0.5 * th; imag = std::sin(h) / th; real = std::cos(h); }
    SE3 T;
    T.d[0] = imag * w[0]; T.d[1] = imag * w[1]; T.d[2] = imag * w[2]; T.d[3] = real;
    double A, B;
    if (th < 1e-10) { A = 0.5; B = 1.0 / 6.0; }
    else { A = (1.0 - std::cos(th)) / th2; B = (th - std::sin(th)) / (th2 * th); }
    double c0 = w[1] * a[2] - w[2] * a[1], c1 = w[2] * a[0] - w[0] * a[2], c2 = w[0] * a[1] - w[1] * a[0];
    double e0 = w[1] * c2 - w[2] * c1, e1 = w[2] * c0 - w[0] * c2, e2 = w[0] * c1 - w[1] * c0;
    T.d[4] = a[0] + A * c0 + B * e0; T.d[5] = a[1] + A * c1 + B * e1; T.d[6] = a[2] + A * c2 + B * e2;
    return T;
}
void SE3::log(double *a) const
{
    double n2 = d[0] * d[0] + d[1] * d[1] + d[2] * d[2], w = d[3], f, th;
    if (n2 < 1e-10 * 1e-10) { f = 2.0 / w - (2.0 / 3.0) * n2 / (w * w * w); th = f * std::sqrt(n2); }
    else { double n = std::sqrt(n2); double at = (w < 0) ? std::atan2(-n, -w) : std::atan2(n, w); f = 2.0 * at / n; th = f * n; }
    double om[3] = {f * d[0], f * d[1], f * d[2]}, c;
    if (std::fabs(th) < 1e-10) c = 1.0 / 12.0;
    else { double h = 0.5 * th; c = (1.0 - th * std::cos(h) / (2.0 * std::sin(h))) / (th * th); }
    const double *t = d + 4;
    double x0 = om[1] * t[2] - om[2] * t[1], x1 = om[2] * t[0] - om[0] * t[2], x2 = om[0] * t[1] - om[1] * t[0];
    double y0 = om[1] * x2 - om[2] * x1, y1 = om[2] * x0 - om[0] * x2, y2 = om[0] * x1 - om[1] * x0;
    a[0] = t[0] - 0.5 * x0 + c * y0; a[1] = t[1] - 0.5 * x1 + c * y1; a[2] = t[2] - 0.5 * x2 + c * y2;
    a[3] = om[0]; a[4] = om[1]; a[5] = om[2];
}

// ------------------------------------------------------------------ MapPoint / Map
void MapPoint::RemoveObservation(const Observation &o)
{   // src/mappoint.cpp:38-78
    for (auto it = observations_->begin(); it != observations_->end(); ++it) {
        if (*it == o) {
            observations_->erase(it);
            Feature &f = o.feature();
            if (f.outlier_) f.map_point_ = -1;
            observed_times_--;
            break;
        }
    }
}

MapPoint *Map::CreateNewMappoint()
{
    if (n_points_ % kChunk == 0) { chunks_.emplace_back(new MapPoint[kChunk]); obs_chunks_.emplace_back(new ObsList[kChunk]); }
    MapPoint *mp = GetMapPoint((long)n_points_);
    mp->observations_ = &obs_chunks_[n_points_ / kChunk][n_points_ % kChunk];
    mp->id_ = n_points_++;
    return mp;
}
void Map::InsertMapPoint(MapPoint *mp) { landmarks_.insert_or_assign(mp->id_, mp); active_landmarks_.insert_or_assign(mp->id_, mp); }

void Map::CleanMap()
{
    active_landmarks_.erase_if([](const LandmarkMap::value_type &kv) { return kv.second->observed_times_ == 0; });
}

void Map::InsertKeyFrame(Frame::Ptr frame)
{
    current_frame_ = frame;
    keyframes_[frame->keyframe_id_] = frame;
    active_keyframes_[frame->keyframe_id_] = frame;
    if ((int)active_keyframes_.size() > num_active_keyframes_) RemoveOldKeyframe();
}

void Map::RemoveOldKeyframe()
{
    if (!current_frame_) return;
    double max_dis = 0, min_dis = 999999;
    unsigned long max_kf_id = 0, min_kf_id = 0;
    SE3 Twc = current_frame_->Pose().inverse();
    for (const auto &kf : active_keyframes_) {
        if (kf.second == current_frame_) continue;
        double lg[6];
        (kf.second->Pose() * Twc).log(lg);
        double dis = std::sqrt(lg[0] * lg[0] + lg[1] * lg[1] + lg[2] * lg[2] + lg[3] * lg[3] + lg[4] * lg[4] + lg[5] * lg[5]);
        if (dis > max_dis) { max_dis = dis; max_kf_id = kf.first; }
        if (dis < min_dis) { min_dis = dis; min_kf_id = kf.first; }
    }
    const double min_dis_th = 0.2;
    Frame::Ptr frame_to_remove = (min_dis < min_dis_th) ? active_keyframes_.at(min_kf_id) : active_keyframes_.at(max_kf_id);
    active_keyframes_.erase(frame_to_remove->keyframe_id_);
    for (size_t i = 0; i < frame_to_remove->feature_left_.size(); i++) {
        Feature &f = frame_to_remove->feature_left_[i];
        if (MapPoint *mp = GetMapPoint(f.map_point_)) mp->RemoveObservation(Observation{frame_to_remove.get(), true, (int)i});
    }
    for (size_t i = 0; i < frame_to_remove->feature_right_.size(); i++) {
        Feature &f = frame_to_remove->feature_right_[i];
        if (!f.valid) continue;
        if (MapPoint *mp = GetMapPoint(f.map_point_)) mp->RemoveObservation(Observation{frame_to_remove.get(), false, (int)i});
    }
    CleanMap();
}

// ------------------------------------------------------------------ Frontend
Frontend::Frontend(const Config &cfg) : cfg_(cfg) {}

Frame::Ptr Frontend::CreateFrame()
{
    Frame::Ptr f = std::make_shared<Frame>();
    f->id_ = frame_factory_id_++;
    f->feature_left_.reserve(256);      // tracked + newly detected features of one frame: one allocation
    return f;
}

void Frontend::begin_AddFrame(Frame::Ptr frame, int img_w, int img_h)
{
    current_frame_ = frame;
    img_w_ = img_w; img_h_ = img_h;
    phase_track_ = phase_detect_ = phase_backend_ = initing_ = init_ok_ = false;
    switch (status_) {
    case FrontendStatus::INITING:
        phase_detect_ = true; initing_ = true;      // StereoInit :216-249
        break;
    case FrontendStatus::TRACKING_GOOD:
    case FrontendStatus::TRACKING_BAD:
        phase_track_ = true;                        // Track :645-688
        if (last_frame_) current_frame_->SetPose(relative_motion_ * last_frame_->Pose());
        break;
    case FrontendStatus::LOST:
        break;                                      // Reset() is "not implemented" in the reference (:723-731)
    }
}

void Frontend::prepare_TrackLastFrame(LkRequest &rq)
{
    const std::vector<Feature> &lf = last_frame_->feature_left_;
    const size_t n = lf.size();
    rq.prev_xy.resize(2 * n); rq.next_xy.resize(2 * n);
    float *pv = rq.prev_xy.data(), *nx = rq.next_xy.data();
    const SE3 Tcw = current_frame_->Pose();
    for (size_t i = 0; i < n; i++) {
        const Feature &f = lf[i];
        if (i + 8 < n) map_->PrefetchMapPoint(lf[i + 8].map_point_);
        pv[2 * i] = f.x; pv[2 * i + 1] = f.y;
        if (MapPoint *mp = map_->GetMapPoint(f.map_point_)) {
            Vec2 px = camera_left_->world2pixel(mp->pos_, Tcw);
            nx[2 * i] = (float)px.x; nx[2 * i + 1] = (float)px.y;
        } else {
            nx[2 * i] = f.x; nx[2 * i + 1] = f.y;
        }
    }
    rq.status.assign(n, 0);
}

int Frontend::finish_TrackLastFrame(const LkRequest &rq)
{
    const size_t n = rq.status.size();
    const float w = (float)img_w_, h = (float)img_h_;
    const Feature *lf = last_frame_->feature_left_.data();
    std::vector<Feature> &cf = current_frame_->feature_left_;
    cf.resize(n);
    Feature *out = cf.data();
    int num_good_pts = 0;
    for (size_t i = 0; i < n; i++) {
        if (!rq.status[i]) continue;
        float x = rq.next_xy[2 * i], y = rq.next_xy[2 * i + 1];
        if (!(y >= 0 && y < h && x >= 0 && x < w)) continue;      // written positively: NaN is out of bounds
        Feature &f = out[num_good_pts++];
        f.x = x; f.y = y; f.size = 7;
        f.map_point_ = lf[i].map_point_;
    }
    cf.resize((size_t)num_good_pts);
    last_tracked = num_good_pts;
    return num_good_pts;
}

void Frontend::prepare_EstimateCurrentPose(PoseRequest &rq)
{
    camera_left_->K(rq.K);
    SE3 T = current_frame_->Pose();
    for (int i = 0; i < 7; i++) rq.T0[i] = T.d[i];
    const std::vector<Feature> &cf = current_frame_->feature_left_;
    const size_t n = cf.size();
    rq.pts_w.resize(3 * n); rq.uv.resize(2 * n); rq.feat_index.resize(n);
    double *pw = rq.pts_w.data(), *uv = rq.uv.data();
    int *fi = rq.feat_index.data();
    size_t m = 0;
    for (size_t i = 0; i < n; i++) {
        const Feature &f = cf[i];
        if (i + 8 < n) map_->PrefetchMapPoint(cf[i + 8].map_point_);
        if (MapPoint *mp = map_->GetMapPoint(f.map_point_)) {
            fi[m] = (int)i;
            pw[3 * m] = mp->pos_.x; pw[3 * m + 1] = mp->pos_.y; pw[3 * m + 2] = mp->pos_.z;
            uv[2 * m] = (double)f.x; uv[2 * m + 1] = (double)f.y;
            m++;
        }
    }
    rq.pts_w.resize(3 * m); rq.uv.resize(2 * m); rq.feat_index.resize(m);
    rq.outlier.assign(m, 0);
}

int Frontend::finish_EstimateCurrentPose(const PoseRequest &rq)
{
    current_frame_->SetPose(SE3::fromArray(rq.T));
    int cnt_outlier = 0;
    for (size_t k = 0; k < rq.feat_index.size(); k++) {
        Feature &f = current_frame_->feature_left_[rq.feat_index[k]];
        if (rq.outlier[k]) { f.map_point_ = -1; f.outlier_ = false; cnt_outlier++; }
    }
    tracking_inliers_ = (int)rq.feat_index.size() - cnt_outlier;
    if (tracking_inliers_ > cfg_.num_features_tracking) status_ = FrontendStatus::TRACKING_GOOD;
    else if (tracking_inliers_ > cfg_.num_features_tracking_bad) status_ = FrontendStatus::TRACKING_BAD;
    else status_ = FrontendStatus::LOST;
    // InsertKeyframe :576-643
    if (tracking_inliers_ < cfg_.num_features_needed_for_keyframe) InsertKeyframe_begin();
    return tracking_inliers_;
}

void Frontend::adopt_tracked_keyframe(const double pose[7], const double last_pose[7], const float *recs, int n, int status,
                                      int inliers, int img_w, int img_h)
{   // the state Track() would have reached on the host at the point where it calls InsertKeyframe() (frontend.cpp:681)
    struct Rec { float x, y; int64_t lm; };
    const Rec *r = reinterpret_cast<const Rec *>(recs);
    img_w_ = img_w; img_h_ = img_h;
    phase_track_ = true;
    phase_detect_ = phase_backend_ = initing_ = init_ok_ = false;
    if (!(last_frame_ && last_frame_->id_ + 1 == frame_factory_id_)) {
        Frame::Ptr last = std::make_shared<Frame>();      // only its pose is read again (end_AddFrame, relative_motion_)
        last->pose_ = SE3::fromArray(last_pose);
        last_frame_ = last;
    }   // else: the previous frame was handled here too (a keyframe): keep the shared object — this step's BA may refine its
        // pose before relative_motion_ is taken, exactly as with the reference's shared Frame
    current_frame_ = CreateFrame();
    current_frame_->pose_ = SE3::fromArray(pose);
    std::vector<Feature> &cf = current_frame_->feature_left_;
    cf.resize((size_t)n);
    for (int i = 0; i < n; i++) { Feature &f = cf[i]; f.x = r[i].x; f.y = r[i].y; f.size = 7; f.map_point_ = (long)r[i].lm; }
    last_tracked = n;
    status_ = (FrontendStatus)status;
    tracking_inliers_ = inliers;
    InsertKeyframe_begin();
}

void Frontend::InsertKeyframe_begin()
{
    current_frame_->is_keyframe_ = true;                     // Frame::SetKeyFrame
    current_frame_->keyframe_id_ = keyframe_factory_id_++;
    map_->InsertKeyFrame(current_frame_);
    frontend_prev_kf_ = frontend_current_kf_;
    frontend_current_kf_ = current_frame_;
    current_frame_->prev_keyframe_ = (long)frontend_prev_kf_->keyframe_id_;
    current_frame_->relative_pose_pkf_ = current_frame_->Pose() * frontend_prev_kf_->Pose().inverse();
    SetObservationsForKeyFrame();
    phase_detect_ = true;
}

void Frontend::SetObservationsForKeyFrame()
{
    for (size_t i = 0; i < current_frame_->feature_left_.size(); i++) {
        Feature &f = current_frame_->feature_left_[i];
        if (MapPoint *mp = map_->GetMapPoint(f.map_point_)) mp->AddObservation(Observation{current_frame_.get(), true, (int)i});
    }
}

void Frontend::prepare_DetectFeatures(DetectRequest &rq)
{
    rq.occupied_xy.clear();
    for (Feature &f : current_frame_->feature_left_) { rq.occupied_xy.push_back(f.x); rq.occupied_xy.push_back(f.y); }
    rq.out_xy.assign((size_t)2 * cfg_.num_features, 0.f);
    rq.out_resp.assign((size_t)cfg_.num_features, 0.f);
    rq.out_n = 0;
}

int Frontend::finish_DetectFeatures(const DetectRequest &rq)
{
    for (int i = 0; i < rq.out_n; i++) {
        Feature f;
        f.x = rq.out_xy[2 * i]; f.y = rq.out_xy[2 * i + 1]; f.size = 3; f.response = rq.out_resp[i];
        current_frame_->feature_left_.push_back(f);
    }
    last_detected = rq.out_n;
    return rq.out_n;
}

void Frontend::prepare_FindFeaturesInRight(LkRequest &rq)
{
    rq.prev_xy.clear(); rq.next_xy.clear();
    const std::vector<Feature> &cf = current_frame_->feature_left_;
    for (size_t i = 0; i < cf.size(); i++) {
        const Feature &f = cf[i];
        if (i + 8 < cf.size()) map_->PrefetchMapPoint(cf[i + 8].map_point_);
        rq.prev_xy.push_back(f.x); rq.prev_xy.push_back(f.y);
        if (MapPoint *mp = map_->GetMapPoint(f.map_point_)) {
            Vec2 px = camera_right_->world2pixel(mp->pos_, current_frame_->Pose());
            rq.next_xy.push_back((float)px.x); rq.next_xy.push_back((float)px.y);
        } else {
            rq.next_xy.push_back(f.x); rq.next_xy.push_back(f.y);
        }
    }
    rq.status.assign(current_frame_->feature_left_.size(), 0);
}

int Frontend::finish_FindFeaturesInRight(const LkRequest &rq)
{
    int num_good_pts = 0;
    for (size_t i = 0; i < rq.status.size(); i++) {
        float x = rq.next_xy[2 * i], y = rq.next_xy[2 * i + 1];
        Feature f;
        f.is_on_left_image_ = false;
        if (rq.status[i] && y >= 0 && y < (float)img_h_ && x >= 0 && x < (float)img_w_) {
            f.x = x; f.y = y; f.size = 7;
            num_good_pts++;
        } else {
            f.valid = false;
        }
        current_frame_->feature_right_.push_back(f);
    }
    last_right = num_good_pts;
    if (initing_) init_ok_ = num_good_pts >= cfg_.num_features_init;     // StereoInit :227-230
    return num_good_pts;
}

void Frontend::prepare_Triangulate(TriRequest &rq)
{
    rq.left_xy.clear(); rq.right_xy.clear(); rq.feat_index.clear();
    if (initing_ && !init_ok_) { rq.xyz.clear(); rq.ok.clear(); return; }
    for (size_t i = 0; i < current_frame_->feature_left_.size(); i++) {
        Feature &l = current_frame_->feature_left_[i], &r = current_frame_->feature_right_[i];
        if (!r.valid) continue;
        if (!initing_ && l.map_point_ >= 0) continue;        // TriangulateNewPoints: only features without a map point
        rq.feat_index.push_back((int)i);
        rq.left_xy.push_back(l.x); rq.left_xy.push_back(l.y);
        rq.right_xy.push_back(r.x); rq.right_xy.push_back(r.y);
    }
    rq.xyz.assign(rq.feat_index.size() * 3, 0.0);
    rq.ok.assign(rq.feat_index.size(), 0);
}

int Frontend::finish_Triangulate(const TriRequest &rq)
{
    if (initing_ && !init_ok_) return 0;                     // StereoInit failed: stay INITING
    if (initing_) frontend_current_kf_ = current_frame_;
    SE3 current_pose_Twc = current_frame_->Pose().inverse();
    int cnt = 0;
    for (size_t k = 0; k < rq.feat_index.size(); k++) {
        Vec3 p(rq.xyz[3 * k], rq.xyz[3 * k + 1], rq.xyz[3 * k + 2]);
        bool good = rq.ok[k] && p.z > 0;
        if (!initing_) good = good && p.z <= cfg_.max_triangulation_depth;
        if (!good) continue;
        int i = rq.feat_index[k];
        MapPoint *mp = map_->CreateNewMappoint();
        if (!initing_) p = current_pose_Twc * p;
        mp->SetPos(p);
        mp->AddObservation(Observation{current_frame_.get(), true, i});
        mp->AddObservation(Observation{current_frame_.get(), false, i});
        current_frame_->feature_left_[i].map_point_ = (long)mp->id_;
        current_frame_->feature_right_[i].map_point_ = (long)mp->id_;
        map_->InsertMapPoint(mp);
        cnt++;
    }
    last_triangulated = cnt;
    if (initing_) {                                           // BuildInitMap :195-203
        current_frame_->is_keyframe_ = true;
        current_frame_->keyframe_id_ = keyframe_factory_id_++;
        map_->InsertKeyFrame(current_frame_);
        status_ = FrontendStatus::TRACKING_GOOD;
    }
    phase_backend_ = (bool)backend_;
    return 1;
}

void Frontend::end_AddFrame()
{
    if (phase_track_) relative_motion_ = current_frame_->Pose() * last_frame_->Pose().inverse();
    last_frame_ = current_frame_;
}

// ------------------------------------------------------------------ Backend
bool Backend::prepare_Optimize(BaRequest &rq)
{
    rq.poses.clear(); rq.lms.clear(); rq.edge_uv.clear(); rq.chi2.clear(); rq.edge_kf.clear(); rq.edge_lm.clear();
    rq.edge_cam.clear(); rq.kf_ids.clear(); rq.lm_ids.clear(); rq.edge_obs.clear();     // keep the capacity of the last window
    const Map::KeyframesType &keyframes = map_->GetActiveKeyFrames();
    const Map::LandmarksType &landmarks = map_->GetActiveMapPoints();
    // window rows: Frame::ba_index_ (reset below for the keyframes of this window; every other frame keeps -1)
    unsigned long max_kf_id = 0, min_kf_id = 10000000000UL;
    for (const auto &kv : keyframes) {
        kv.second->ba_index_ = (int)rq.kf_ids.size();
        rq.kf_ids.push_back(kv.first);
        SE3 T = kv.second->Pose();
        rq.poses.insert(rq.poses.end(), T.d, T.d + 7);
        max_kf_id = std::max(max_kf_id, kv.first);
        min_kf_id = std::min(min_kf_id, kv.first);
    }
    max_keyframe_id_in_pipeline_ = max_kf_id; min_keyframe_id_in_pipeline_ = min_kf_id;
    for (const auto &lv : landmarks) {
        MapPoint *mp = lv.second;
        if (mp->is_outlier_) continue;
        int lm_index = -1;
        for (const Observation &obs : mp->GetObs()) {
            Feature &feat = obs.feature();
            if (feat.outlier_) continue;
            if (lm_index < 0) {
                lm_index = (int)rq.lm_ids.size();
                rq.lm_ids.push_back(mp->id_);
                Vec3 p = mp->Pos();
                rq.lms.push_back(p.x); rq.lms.push_back(p.y); rq.lms.push_back(p.z);
            }
            const int row = obs.frame->ba_index_;
            if (row < 0) continue;
            rq.edge_kf.push_back(row);
            rq.edge_lm.push_back(lm_index);
            rq.edge_cam.push_back(feat.is_on_left_image_ ? 0 : 1);
            rq.edge_uv.push_back((double)feat.x); rq.edge_uv.push_back((double)feat.y);
            rq.edge_obs.push_back(obs);
        }
    }
    rq.chi2.assign(rq.edge_kf.size(), 0.0);
    for (const auto &kv : keyframes) kv.second->ba_index_ = -1;
    return !rq.edge_kf.empty();
}

void Backend::finish_Optimize(BaRequest &rq)
{
    int cnt_outlier = 0, cnt_inlier = 0, iteration = 0;
    double chi2_th = chi2_th_;
    while (iteration < 5) {
        cnt_outlier = 0; cnt_inlier = 0;
        for (double c : rq.chi2) { if (c > chi2_th) cnt_outlier++; else cnt_inlier++; }
        double inlier_ratio = cnt_inlier / double(cnt_inlier + cnt_outlier);
        if (inlier_ratio > 0.5) break;
        chi2_th *= 2; iteration++;
    }
    last_outliers = cnt_outlier; last_inliers = cnt_inlier;
    for (size_t e = 0; e < rq.chi2.size(); e++) {
        Feature &feat = rq.edge_obs[e].feature();
        if (rq.chi2[e] > chi2_th) {
            feat.outlier_ = true;
            if (MapPoint *mp = map_->GetMapPoint(feat.map_point_)) mp->RemoveObservation(rq.edge_obs[e]);
        } else {
            feat.outlier_ = false;
        }
    }
    const Map::KeyframesType &keyframes = map_->GetActiveKeyFrames();
    for (size_t k = 0; k < rq.kf_ids.size(); k++) keyframes.at(rq.kf_ids[k])->SetPose(SE3::fromArray(&rq.poses[7 * k]));
    for (size_t l = 0; l < rq.lm_ids.size(); l++)
        map_->GetMapPoint((long)rq.lm_ids[l])->SetPos(Vec3(rq.lms[3 * l], rq.lms[3 * l + 1], rq.lms[3 * l + 2]));
    const Map::KeyframesType &all = map_->GetAllKeyFrames();
    for (const auto &kv : keyframes) {
        if (kv.first == 0) continue;
        auto it = all.find((unsigned long)kv.second->prev_keyframe_);
        if (it == all.end()) continue;
        kv.second->relative_pose_pkf_ = kv.second->Pose() * it->second->Pose().inverse();
    }
}

}  // namespace slam
