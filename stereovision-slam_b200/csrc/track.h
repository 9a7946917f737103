// Device-resident per-stream tracking state (csrc/track.cu): the per-frame arithmetic of Frontend::Track()
// (reference src/frontend.cpp:645-688 — motion model :653-656, TrackLastFrame :322-392, EstimateCurrentPose :394-558,
// status :665-679, keyframe test :587, relative motion :685) for a batch of independent streams without a host round
// trip between the seams.  The host (slam::StreamBatch) sees one record per stream per step and takes a stream back
// only when it inserts a keyframe or initialises.  Internal C++ interface between pipeline.cpp and track.cu.
#pragma once
#include <cstdint>
#include "svs_internal.h"
#include "../host/slam.h"

struct TrkOut {            // one per stream per step, written by the device into pinned host memory
    double pose[7];        // T_cw after EstimateCurrentPose
    double last_pose[7];   // pose of the previous frame (what Frontend::Track's relative_motion_ is computed against)
    int32_t status;        // FrontendStatus after the step
    int32_t inliers;       // tracking_inliers_
    int32_t need_kf;       // inliers < num_features_needed_for_keyframe
    int32_t nfeat;         // left features of the current frame
    int32_t n_edges;       // edges of the pose-only problem (features with a landmark)
    int32_t pad_;
};
struct TrkFeat { float x, y; int64_t lm; };                       // a left feature: position + landmark id (-1: none)
struct TrkUpFeat { float x, y; int64_t lm; double pw[3]; };       // + the landmark's world position
struct TrkUpHdr {          // state of one stream handed (back) to the device after the host's keyframe / init path
    int32_t stream, n, status, pad_;
    double pose[7];        // pose of the current frame = the next step's "last frame"
    double rel[7];         // relative_motion_
    int64_t feat_off;      // first TrkUpFeat of this stream in the packed feature array
    int64_t pad2_;
};

struct svs_tracker;
struct TrkParams {
    int B, cap, W, H;
    int num_features_tracking, num_features_tracking_bad, num_features_needed_for_keyframe;
    int lk_win, lk_max_iter;
    double lk_eps, chi2_th;
    slam::CameraModel cam_left;
};
svs_tracker *svs_i_trk_create(svs_ctx *c, const TrkParams &p);
void svs_i_trk_destroy(svs_ctx *c, svs_tracker *t);
// One Track() for every stream whose device status is TRACKING_GOOD / TRACKING_BAD; synchronises the context stream.
// After it: svs_i_trk_out(t)[b] for every stream, and svs_i_trk_kf_feats(t, b) for the streams with need_kf.
int svs_i_trk_step(svs_ctx *c, svs_tracker *t, svs_frameset *fs, long long *lk_points, long long *pose_edges);
const TrkOut *svs_i_trk_out(const svs_tracker *t);
const TrkFeat *svs_i_trk_kf_feats(const svs_tracker *t, int stream);
// Hand n_sel streams (back) to the device (asynchronous on the context stream; the arrays are copied before returning).
int svs_i_trk_upload(svs_ctx *c, svs_tracker *t, int n_sel, const TrkUpHdr *hdrs, const TrkUpFeat *feats, long long n_feats);
// Current left features of one stream (tests / getters): synchronises.
int svs_i_trk_fetch(svs_ctx *c, svs_tracker *t, int stream, const TrkFeat **feats, int *n);
int svs_i_trk_is_active(const svs_tracker *t, int stream);
