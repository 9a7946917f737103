// Host-side mirror of the reference's classes on the hot path, written above the C ABI of
// include/svslam.h (the reference is compiled C++, so the host side is C++ too).  Same class,
// method and field names as the reference so its call sites read the same:
//
//   slam::Camera    include/StereoVisionSLAM/camera.h:9-57,   src/camera.cpp:6-86
//   slam::Feature   include/StereoVisionSLAM/feature.h:14-35
//   slam::Frame     include/StereoVisionSLAM/frame.h:9-78,    src/frame.cpp:10-35
//   slam::MapPoint  include/StereoVisionSLAM/mappoint.h:14-53, src/mappoint.cpp:9-98
//   slam::Map       include/StereoVisionSLAM/map.h:10-59,     src/map.cpp:6-209
//   slam::Frontend  include/StereoVisionSLAM/frontend.h,      src/frontend.cpp:10-768
//   slam::Backend   include/StereoVisionSLAM/backend.h,       src/backend.cpp:9-346
//
// Differences that are deliberate:
//  * Eigen / Sophus / OpenCV types are replaced by the tiny fixed-size types below (those libraries are
//    not available here); SE3 keeps Sophus' representation (unit quaternion + translation) and formulas.
//  * Every third-party call site (GFTT detect, calcOpticalFlowPyrLK, triangulation, the two g2o blocks)
//    is split into a prepare_* / finish_* pair so that many independent streams can be stepped in
//    lock-step and each seam becomes ONE batched svs_* call over all streams (slam::StreamBatch).
//  * Bundle adjustment runs on the synchronous schedule (inside UpdateMap), SURVEY.md §5.  One consequence: on a keyframe
//    step relative_motion_ (src/frontend.cpp:685) is formed AFTER Backend::Optimize has refined the current keyframe's
//    pose, whereas the reference's backend thread usually has not touched it yet when Track() reads it (a race the
//    reference leaves open).  The constant-velocity prior of the next frame therefore already contains the BA correction.
//  * Hash-map iteration orders that the reference leaves unspecified are fixed to ascending id.
#pragma once
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <map>
#include <memory>
#include <vector>

// The small math types are shared with the device-resident tracker (csrc/track.cu): the SAME source text computes the
// motion-model pose, the LK initial guesses and the relative motion on either side, and neither side contracts a*b+c into a
// fused multiply-add (track.cu is built with --fmad=false, the host compiler targets baseline x86-64), so the two paths agree
// bit for bit (tests/test_gpu_pipeline.py::test_device_tracking_is_bit_identical_to_host_tracking).
#ifdef __CUDACC__
#define SLAM_HD __host__ __device__
#else
#define SLAM_HD
#endif

namespace slam {

// ------------------------------------------------------------------ math
struct Vec2 { double x = 0, y = 0; };
struct Vec3 {
    double x = 0, y = 0, z = 0;
    SLAM_HD Vec3() {}
    SLAM_HD Vec3(double a, double b, double c) : x(a), y(b), z(c) {}
    SLAM_HD Vec3 operator+(const Vec3 &o) const { return {x + o.x, y + o.y, z + o.z}; }
    SLAM_HD Vec3 operator-(const Vec3 &o) const { return {x - o.x, y - o.y, z - o.z}; }
    double norm() const { return std::sqrt(x * x + y * y + z * z); }
};

// Sophus::SE3d: unit quaternion (x,y,z,w) + translation; d = [qx qy qz qw tx ty tz] is the C-ABI layout.
struct SE3 {
    double d[7] = {0, 0, 0, 1, 0, 0, 0};
    SLAM_HD SE3() {}
    SLAM_HD static SE3 fromArray(const double *p) { SE3 T; for (int i = 0; i < 7; i++) T.d[i] = p[i]; return T; }
    SLAM_HD static SE3 fromTranslation(const Vec3 &t) { SE3 T; T.d[4] = t.x; T.d[5] = t.y; T.d[6] = t.z; return T; }
    // Eigen QuaternionBase::_transformVector; inline: called once or twice per feature per frame
    SLAM_HD Vec3 rotate(const Vec3 &p) const
    {
        double ux = 2.0 * (d[1] * p.z - d[2] * p.y), uy = 2.0 * (d[2] * p.x - d[0] * p.z), uz = 2.0 * (d[0] * p.y - d[1] * p.x);
        return {p.x + d[3] * ux + (d[1] * uz - d[2] * uy), p.y + d[3] * uy + (d[2] * ux - d[0] * uz), p.z + d[3] * uz + (d[0] * uy - d[1] * ux)};
    }
    SLAM_HD Vec3 operator*(const Vec3 &p) const { Vec3 r = rotate(p); return {r.x + d[4], r.y + d[5], r.z + d[6]}; }   // point action
    // a pure translation (the rectified cameras' extrinsics): rotate() returns its argument bit for bit, so it can be skipped
    SLAM_HD bool rotation_is_identity() const { return d[0] == 0.0 && d[1] == 0.0 && d[2] == 0.0 && d[3] == 1.0; }
    SLAM_HD SE3 operator*(const SE3 &o) const      // composition (Sophus renormalisation included)
    {
        const double *A = d, *B = o.d;
        SE3 C;
        C.d[0] = A[3] * B[0] + A[0] * B[3] + A[1] * B[2] - A[2] * B[1];
        C.d[1] = A[3] * B[1] + A[1] * B[3] + A[2] * B[0] - A[0] * B[2];
        C.d[2] = A[3] * B[2] + A[2] * B[3] + A[0] * B[1] - A[1] * B[0];
        C.d[3] = A[3] * B[3] - A[0] * B[0] - A[1] * B[1] - A[2] * B[2];
        double n2 = C.d[0] * C.d[0] + C.d[1] * C.d[1] + C.d[2] * C.d[2] + C.d[3] * C.d[3];
        if (n2 != 1.0) { double s = 2.0 / (1.0 + n2); for (int i = 0; i < 4; i++) C.d[i] *= s; }
        Vec3 t = rotate(o.translation());
        C.d[4] = t.x + d[4]; C.d[5] = t.y + d[5]; C.d[6] = t.z + d[6];
        return C;
    }
    SLAM_HD SE3 inverse() const
    {
        SE3 I;
        I.d[0] = -d[0]; I.d[1] = -d[1]; I.d[2] = -d[2]; I.d[3] = d[3];
        Vec3 t = I.rotate(translation());
        I.d[4] = -t.x; I.d[5] = -t.y; I.d[6] = -t.z;
        return I;
    }
    SLAM_HD Vec3 translation() const { return {d[4], d[5], d[6]}; }
    static SE3 exp(const double *tangent6);         // (upsilon, omega)
    void log(double *tangent6) const;
};

// ------------------------------------------------------------------ camera
// The plain-data part of Camera (usable as a kernel argument by the device-resident tracker).
struct CameraModel {
    double fx_ = 0, fy_ = 0, cx_ = 0, cy_ = 0, baseline_ = 0;
    SE3 pose_, pose_inv_;   // extrinsic: stereo-system frame -> this camera
    bool pure_translation_ = false;
    SLAM_HD Vec3 world2camera(const Vec3 &p_w, const SE3 &T_c_w) const
    {
        Vec3 v = T_c_w * p_w;
        if (pure_translation_) return {v.x + pose_.d[4], v.y + pose_.d[5], v.z + pose_.d[6]};   // == pose_ * v, bit for bit
        return pose_ * v;
    }
    SLAM_HD Vec2 camera2pixel(const Vec3 &p_c) const { Vec2 r; r.x = fx_ * p_c.x / p_c.z + cx_; r.y = fy_ * p_c.y / p_c.z + cy_; return r; }
    SLAM_HD Vec2 world2pixel(const Vec3 &p_w, const SE3 &T_c_w) const { return camera2pixel(world2camera(p_w, T_c_w)); }
};

class Camera : public CameraModel {
public:
    typedef std::shared_ptr<Camera> Ptr;
    Camera() {}
    Camera(double fx, double fy, double cx, double cy, double baseline, const SE3 &pose)
    {
        fx_ = fx; fy_ = fy; cx_ = cx; cy_ = cy; baseline_ = baseline; pose_ = pose;
        pose_inv_ = pose_.inverse();
        pure_translation_ = pose_.rotation_is_identity();
    }
    SE3 pose() const { return pose_; }
    void K(double k4[4]) const { k4[0] = fx_; k4[1] = fy_; k4[2] = cx_; k4[3] = cy_; }
    Vec3 camera2world(const Vec3 &p_c, const SE3 &T_c_w) const { return T_c_w.inverse() * (pose_inv_ * p_c); }
    Vec3 pixel2camera(const Vec2 &p_p, double depth = 1) const { return {(p_p.x - cx_) * depth / fx_, (p_p.y - cy_) * depth / fy_, depth}; }
    Vec3 pixel2world(const Vec2 &p_p, const SE3 &T_c_w, double depth = 1) const { return camera2world(pixel2camera(p_p, depth), T_c_w); }
};

// ------------------------------------------------------------------ data model
struct Feature {                // feature.h:23-30 (cv::KeyPoint reduced to what the path reads)
    float x = 0, y = 0;         // position_.pt
    float size = 0, response = 0;
    long map_point_ = -1;       // weak_ptr<MapPoint> -> landmark id, -1 = expired / none
    bool outlier_ = false;
    bool is_on_left_image_ = true;
    bool valid = true;          // false = the nullptr entries of Frame::feature_right_
};

class Frame {
public:
    typedef std::shared_ptr<Frame> Ptr;
    unsigned long id_ = 0, keyframe_id_ = 0;
    bool is_keyframe_ = false;
    SE3 pose_;                  // T_cw
    double time_stamp_ = 0;
    std::vector<Feature> feature_left_, feature_right_;
    long prev_keyframe_ = -1;   // keyframe id of the previous keyframe
    int ba_index_ = -1;         // row of this keyframe in the window problem being built (Backend::prepare_Optimize)
    SE3 relative_pose_pkf_;
    SE3 Pose() const { return pose_; }
    void SetPose(const SE3 &p) { pose_ = p; }
};

struct Observation {            // weak_ptr<Feature> of a keyframe feature
    Frame *frame = nullptr;
    bool left = true;
    int index = 0;
    bool operator==(const Observation &o) const { return frame == o.frame && left == o.left && index == o.index; }
    Feature &feature() const { return left ? frame->feature_left_[index] : frame->feature_right_[index]; }
};

// Observation list of a landmark (std::list<weak_ptr<Feature>> in the reference): insertion order kept.  The first kInline
// entries are stored in place (most landmarks are seen 2-4 times) and the lists live in an arena parallel to the landmark
// arena, so walking the active landmarks of a window (Backend::Optimize) is two linear walks instead of one heap hop per
// landmark, while the per-frame lookups (position only) touch 48-byte MapPoints.
class ObsList {
public:
    static constexpr int kInline = 4;
    typedef const Observation *const_iterator;
    ObsList() {}
    ObsList(const ObsList &) = delete;
    ObsList &operator=(const ObsList &) = delete;
    ~ObsList() { delete[] heap_; }
    size_t size() const { return n_; }
    bool empty() const { return n_ == 0; }
    const Observation *begin() const { return data(); }
    const Observation *end() const { return data() + n_; }
    void push_back(const Observation &o)
    {
        if (n_ == cap_) grow();
        data()[n_++] = o;
    }
    void erase(const Observation *it)
    {
        Observation *d = data();
        for (size_t i = (size_t)(it - d); i + 1 < n_; i++) d[i] = d[i + 1];
        n_--;
    }
private:
    Observation *data() { return heap_ ? heap_ : inline_; }
    const Observation *data() const { return heap_ ? heap_ : inline_; }
    void grow()
    {
        uint32_t nc = cap_ * 2;
        Observation *h = new Observation[nc];
        for (uint32_t i = 0; i < n_; i++) h[i] = data()[i];
        delete[] heap_;
        heap_ = h; cap_ = nc;
    }
    Observation inline_[kInline];
    Observation *heap_ = nullptr;
    uint32_t n_ = 0, cap_ = kInline;
};

// Sorted flat map id -> MapPoint* with the few std::map operations the path uses (the reference's unordered_map, iterated
// in ascending id here): landmark ids only grow, so insertion is an append and iteration is a linear scan.
class LandmarkMap {
public:
    typedef std::pair<unsigned long, class MapPoint *> value_type;
    typedef std::vector<value_type>::const_iterator const_iterator;
    typedef std::vector<value_type>::iterator iterator;
    size_t size() const { return v_.size(); }
    const_iterator begin() const { return v_.begin(); }
    const_iterator end() const { return v_.end(); }
    iterator begin() { return v_.begin(); }
    iterator end() { return v_.end(); }
    void insert_or_assign(unsigned long id, class MapPoint *mp)
    {
        if (v_.empty() || v_.back().first < id) { v_.emplace_back(id, mp); return; }
        iterator it = std::lower_bound(v_.begin(), v_.end(), id, [](const value_type &a, unsigned long b) { return a.first < b; });
        if (it != v_.end() && it->first == id) it->second = mp;
        else v_.insert(it, value_type(id, mp));
    }
    template <class Pred> void erase_if(Pred pred) { v_.erase(std::remove_if(v_.begin(), v_.end(), pred), v_.end()); }
private:
    std::vector<value_type> v_;
};

class MapPoint {
public:
    unsigned long id_ = 0;                    // 48 bytes: the per-frame path reads pos_ of ~190 landmarks per stream
    Vec3 pos_;
    int observed_times_ = 0;
    bool is_outlier_ = false;
    ObsList *observations_ = nullptr;         // std::list in the reference; insertion order kept; storage: Map's arena
    Vec3 Pos() const { return pos_; }
    void SetPos(const Vec3 &p) { pos_ = p; }
    void AddObservation(const Observation &o) { observations_->push_back(o); observed_times_++; }  // mappoint.cpp:22-36
    void RemoveObservation(const Observation &o);                                                  // mappoint.cpp:38-78
    const ObsList &GetObs() const { return *observations_; }
};

class Map {
public:
    typedef std::shared_ptr<Map> Ptr;
    typedef LandmarkMap LandmarksType;                             // ascending id (reference: unordered_map)
    typedef std::map<unsigned long, Frame::Ptr> KeyframesType;
    explicit Map(int num_active_keyframes) : num_active_keyframes_(num_active_keyframes) {}
    void CleanMap();                                  // map.cpp:21-40
    void InsertKeyFrame(Frame::Ptr frame);            // map.cpp:53-67
    MapPoint *CreateNewMappoint();                    // mappoint.cpp:88-97 (id factory is per map = per stream)
    void InsertMapPoint(MapPoint *mp);                // map.cpp:69-74
    // Landmarks live in fixed chunks (stable addresses, creation order = memory order): features tracked from the same
    // keyframe have neighbouring ids, so the per-feature lookups of a frame walk a few cache lines instead of chasing one
    // heap pointer per landmark.
    MapPoint *GetMapPoint(long id) { return id >= 0 ? &chunks_[(size_t)id / kChunk][(size_t)id % kChunk] : nullptr; }
    void PrefetchMapPoint(long id) const { if (id >= 0) __builtin_prefetch(&chunks_[(size_t)id / kChunk][(size_t)id % kChunk]); }
    const LandmarksType &GetAllMapPoints() const { return landmarks_; }
    const KeyframesType &GetAllKeyFrames() const { return keyframes_; }
    const LandmarksType &GetActiveMapPoints() const { return active_landmarks_; }
    const KeyframesType &GetActiveKeyFrames() const { return active_keyframes_; }
private:
    void RemoveOldKeyframe();                         // map.cpp:76-181
    static constexpr size_t kChunk = 256;
    std::vector<std::unique_ptr<MapPoint[]>> chunks_;
    std::vector<std::unique_ptr<ObsList[]>> obs_chunks_;
    size_t n_points_ = 0;
    LandmarksType landmarks_, active_landmarks_;
    KeyframesType keyframes_, active_keyframes_;
    Frame::Ptr current_frame_;
    int num_active_keyframes_ = 9;
};

// ------------------------------------------------------------------ configuration (Appendix C of SURVEY.md)
struct Config {
    int num_features = 150, num_features_init = 50, num_features_tracking = 50, num_features_tracking_bad = 20;
    int num_features_needed_for_keyframe = 80;
    double max_triangulation_depth = 300.0;
    int num_active_keyframes = 10;
    int backend_on = 1;
    double chi2_th = 5.991;
    // hard-coded constants of the reference, exposed so the synthetic config 4 can override them
    double gftt_quality = 0.01, gftt_min_distance = 20.0;        // src/frontend.cpp:24
    int lk_win = 11, lk_max_level = 3, lk_max_iter = 30;         // src/frontend.cpp:107-108
    double lk_eps = 0.01;
    int ba_max_iter = 10, ba_jacobian_mode = 0;                  // src/backend.cpp:164
    int oracle_simd_granule = 32;
    // The frontend reads the right image only in FindFeaturesInRight (keyframes / init): ingest it only then (identical
    // results, about half the image traffic).  0 = ingest both eyes of every frame like Dataset::NextFrame.
    int lazy_right_ingest = 1;
    // Track()'s per-frame arithmetic on the device with device-resident feature / pose state (csrc/track.cu); 0 = every seam
    // is a host round trip (the round-1 path, kept as the bit-identity reference of the device path)
    int device_tracking = 1;
};

enum class FrontendStatus { INITING, TRACKING_GOOD, TRACKING_BAD, LOST };

// Ragged request/response buffers of one stream for one batched seam call.
struct LkRequest { std::vector<float> prev_xy, next_xy; std::vector<uint8_t> status; };
struct PoseRequest { std::vector<double> pts_w, uv; double K[4]; double T0[7]; double T[7]; std::vector<uint8_t> outlier; int n_inlier = 0; std::vector<int> feat_index; };
struct DetectRequest { std::vector<float> occupied_xy; std::vector<float> out_xy, out_resp; int out_n = 0; };
struct TriRequest { std::vector<float> left_xy, right_xy; std::vector<double> xyz; std::vector<uint8_t> ok; std::vector<int> feat_index; };
struct BaRequest {
    std::vector<double> poses, lms, edge_uv, chi2;
    std::vector<int32_t> edge_kf, edge_lm;
    std::vector<uint8_t> edge_cam;
    std::vector<unsigned long> kf_ids, lm_ids;
    std::vector<Observation> edge_obs;
};

class Backend {
public:
    Backend(const Config &cfg) : chi2_th_(cfg.chi2_th) {}
    void SetCameras(Camera::Ptr l, Camera::Ptr r) { cam_left_ = l; cam_right_ = r; }
    void SetMap(Map::Ptr m) { map_ = m; }
    // Backend::Optimize (src/backend.cpp:9-248) split around the g2o block (svs_ba_optimize):
    bool prepare_Optimize(BaRequest &rq);      // graph construction :39-158, ascending ids
    void finish_Optimize(BaRequest &rq);       // chi2 post-pass :167-213, write-back :224-231, relative poses :235-246
    unsigned long max_keyframe_id_in_pipeline_ = 0, min_keyframe_id_in_pipeline_ = 0;
    int last_outliers = 0, last_inliers = 0;
private:
    Camera::Ptr cam_left_, cam_right_;
    Map::Ptr map_;
    double chi2_th_;
};

class Frontend {
public:
    explicit Frontend(const Config &cfg);
    void SetMap(Map::Ptr map) { map_ = map; }
    void SetBackend(std::shared_ptr<Backend> b) { backend_ = b; }
    void SetCameras(Camera::Ptr l, Camera::Ptr r) { camera_left_ = l; camera_right_ = r; }
    FrontendStatus GetStatus() const { return status_; }
    Frame::Ptr GetLastFrame() const { return last_frame_; }
    Frame::Ptr CreateFrame();                     // Frame::CreateFrame (id factory per stream)

    // Frontend::AddFrame (src/frontend.cpp:690-721) is driven by slam::StreamBatch through these phases.
    void begin_AddFrame(Frame::Ptr frame, int img_w, int img_h);
    //   Track (:645-688)
    bool wants_track() const { return phase_track_; }
    void prepare_TrackLastFrame(LkRequest &rq);          // :331-347
    int finish_TrackLastFrame(const LkRequest &rq);      // :361-381
    void prepare_EstimateCurrentPose(PoseRequest &rq);   // :408-471
    int finish_EstimateCurrentPose(const PoseRequest &rq);   // :542-556, then status :665-679 and keyframe test :587
    //   keyframe / init branch
    bool wants_detect() const { return phase_detect_; }
    void prepare_DetectFeatures(DetectRequest &rq);      // :42-47
    int finish_DetectFeatures(const DetectRequest &rq);  // :55-59
    void prepare_FindFeaturesInRight(LkRequest &rq);     // :82-100
    int finish_FindFeaturesInRight(const LkRequest &rq); // :113-130
    void prepare_Triangulate(TriRequest &rq);            // BuildInitMap :155-169 / TriangulateNewPoints :267-281
    int finish_Triangulate(const TriRequest &rq);        // :174-192 / :286-307 ; returns 1 when the backend must run
    bool wants_backend() const { return phase_backend_; }
    void end_AddFrame();                                 // relative_motion_ :685, last_frame_ :718

    // ---- device-resident tracking (slam::StreamBatch with Config::device_tracking, csrc/track.cu): Track()'s per-frame
    // arithmetic runs on the GPU and the host sees one record per stream per step; a stream comes back to these host
    // classes only when it inserts a keyframe (or initialises), with the tracked frame the device hands over.
    void note_tracked_frame(int status, int inliers) { frame_factory_id_++; status_ = (FrontendStatus)status; tracking_inliers_ = inliers; }
    void skip_frame() { frame_factory_id_++; }
    void adopt_tracked_keyframe(const double pose[7], const double last_pose[7], const float *xy_lm /* TrkFeat records */, int n,
                                int status, int inliers, int img_w, int img_h);
    const SE3 &relative_motion() const { return relative_motion_; }
    Map *map() const { return map_.get(); }

    int tracking_inliers_ = 0;
    int last_detected = 0, last_right = 0, last_triangulated = 0, last_tracked = 0;
    Frame::Ptr current_frame_, last_frame_;
private:
    void SetObservationsForKeyFrame();                   // :560-574
    void InsertKeyframe_begin();                         // :587-616
    Config cfg_;
    FrontendStatus status_ = FrontendStatus::INITING;
    Frame::Ptr frontend_current_kf_, frontend_prev_kf_;
    SE3 relative_motion_;
    Camera::Ptr camera_left_, camera_right_;
    Map::Ptr map_;
    std::shared_ptr<Backend> backend_;
    int img_w_ = 0, img_h_ = 0;
    unsigned long frame_factory_id_ = 0, keyframe_factory_id_ = 0;
    bool phase_track_ = false, phase_detect_ = false, phase_backend_ = false, initing_ = false, init_ok_ = false;
};

}  // namespace slam
