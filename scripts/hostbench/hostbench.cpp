// CPU-only micro-benchmark of the host bookkeeping on the per-frame (non-keyframe) path, with faked GPU results.
#include <algorithm>
#include <chrono>
#include <cstdio>
#include <random>
#include "slam.h"
using namespace slam;
static double now() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }
struct St { Map::Ptr map; std::shared_ptr<Frontend> fe; std::shared_ptr<Backend> be; LkRequest lk; PoseRequest pose; DetectRequest det; TriRequest tri; BaRequest ba; };
int main(int argc, char **argv)
{
    int B = argc > 1 ? atoi(argv[1]) : 4096, steps = argc > 2 ? atoi(argv[2]) : 30;
    Config cfg; cfg.backend_on = 1;
    const int kf_period = argc > 3 ? atoi(argv[3]) : 25;   // every stream inserts a keyframe every kf_period frames (staggered)
    double K[4] = {353.5, 353.5, 300.9, 91.6};
    auto cl = std::make_shared<Camera>(K[0], K[1], K[2], K[3], 0.0, SE3());
    auto cr = std::make_shared<Camera>(K[0], K[1], K[2], K[3], 0.537, SE3::fromTranslation(Vec3(-0.537, 0, 0)));
    std::vector<St> S(B);
    std::mt19937 rng(1);
    std::uniform_real_distribution<float> ux(20, 590), uy(20, 160);
    for (auto &s : S) {
        s.map = std::make_shared<Map>(10); s.fe = std::make_shared<Frontend>(cfg); s.fe->SetMap(s.map); s.fe->SetCameras(cl, cr);
        s.be = std::make_shared<Backend>(cfg); s.be->SetMap(s.map); s.be->SetCameras(cl, cr); s.fe->SetBackend(s.be);
        // init frame: detect 190, right match, triangulate
        Frame::Ptr f = s.fe->CreateFrame(); s.fe->begin_AddFrame(f, 613, 185);
        s.fe->prepare_DetectFeatures(s.det);
        s.det.out_xy.resize(2 * 190); s.det.out_resp.resize(190); s.det.out_n = 150;
        for (int i = 0; i < 150; i++) { s.det.out_xy[2 * i] = ux(rng); s.det.out_xy[2 * i + 1] = uy(rng); s.det.out_resp[i] = 1; }
        s.fe->finish_DetectFeatures(s.det);
        s.fe->prepare_FindFeaturesInRight(s.lk);
        for (size_t i = 0; i < s.lk.status.size(); i++) { s.lk.next_xy[2 * i] = s.lk.prev_xy[2 * i] - 10; s.lk.next_xy[2 * i + 1] = s.lk.prev_xy[2 * i + 1]; s.lk.status[i] = 1; }
        s.fe->finish_FindFeaturesInRight(s.lk);
        s.fe->prepare_Triangulate(s.tri);
        for (size_t k = 0; k < s.tri.feat_index.size(); k++) {
            double z = 0.537 * 353.5 / 10.0, x = (s.tri.left_xy[2 * k] - K[2]) * z / K[0], y = (s.tri.left_xy[2 * k + 1] - K[3]) * z / K[1];
            s.tri.xyz[3 * k] = x; s.tri.xyz[3 * k + 1] = y; s.tri.xyz[3 * k + 2] = z; s.tri.ok[k] = 1;
        }
        s.fe->finish_Triangulate(s.tri);
        s.fe->end_AddFrame();
    }
    double t[4] = {0, 0, 0, 0}, tk[4] = {0, 0, 0, 0};
    long nkf = 0;
    for (int it = 0; it < steps; it++) {
        double t0 = now();
        for (auto &s : S) { Frame::Ptr f = s.fe->CreateFrame(); s.fe->begin_AddFrame(f, 613, 185); if (s.fe->wants_track()) s.fe->prepare_TrackLastFrame(s.lk); }
        double t1 = now(); t[0] += t1 - t0;
        for (auto &s : S) for (size_t i = 0; i < s.lk.status.size(); i++) { s.lk.next_xy[2 * i] = s.lk.prev_xy[2 * i] + 0.01f; s.lk.next_xy[2 * i + 1] = s.lk.prev_xy[2 * i + 1]; s.lk.status[i] = 1; }
        t0 = now();
        for (auto &s : S) { s.fe->finish_TrackLastFrame(s.lk); s.fe->prepare_EstimateCurrentPose(s.pose); }
        t1 = now(); t[1] += t1 - t0;
        // fake pose result; a keyframe is forced by declaring half of the edges outliers (inliers < 80)
        for (size_t b = 0; b < S.size(); b++) {
            St &s = S[b];
            for (int i = 0; i < 7; i++) s.pose.T[i] = s.pose.T0[i];
            bool kf = ((it + (int)b) % kf_period) == 0;
            if (kf) for (size_t i = 0; i < s.pose.outlier.size(); i++) s.pose.outlier[i] = (i % 2 == 0 && i > 100) || i >= 79 * 2 ? 1 : (i % 2);
        }
        t0 = now();
        for (auto &s : S) { s.fe->finish_EstimateCurrentPose(s.pose); if (s.fe->wants_detect()) s.fe->prepare_DetectFeatures(s.det); }
        t1 = now(); t[2] += t1 - t0;
        // keyframe branch with faked detections / right matches / triangulation / BA
        for (auto &s : S) {
            if (!s.fe->wants_detect()) continue;
            nkf++;
            int nd = std::min(150, 190 - (int)s.fe->current_frame_->feature_left_.size());
            s.det.out_n = nd > 0 ? nd : 0;
            for (int i = 0; i < s.det.out_n; i++) { s.det.out_xy[2 * i] = ux(rng); s.det.out_xy[2 * i + 1] = uy(rng); s.det.out_resp[i] = 1; }
            double a0 = now();
            s.fe->finish_DetectFeatures(s.det); s.fe->prepare_FindFeaturesInRight(s.lk);
            double a1 = now(); tk[0] += a1 - a0;
            for (size_t i = 0; i < s.lk.status.size(); i++) { s.lk.next_xy[2 * i] = s.lk.prev_xy[2 * i] - 10; s.lk.next_xy[2 * i + 1] = s.lk.prev_xy[2 * i + 1]; s.lk.status[i] = 1; }
            a0 = now();
            s.fe->finish_FindFeaturesInRight(s.lk); s.fe->prepare_Triangulate(s.tri);
            a1 = now(); tk[1] += a1 - a0;
            for (size_t k = 0; k < s.tri.feat_index.size(); k++) {
                double z = 0.537 * 353.5 / 10.0, x = (s.tri.left_xy[2 * k] - K[2]) * z / K[0], y = (s.tri.left_xy[2 * k + 1] - K[3]) * z / K[1];
                s.tri.xyz[3 * k] = x; s.tri.xyz[3 * k + 1] = y; s.tri.xyz[3 * k + 2] = z; s.tri.ok[k] = 1;
            }
            a0 = now();
            bool run = s.fe->finish_Triangulate(s.tri) && s.fe->wants_backend() && s.be->prepare_Optimize(s.ba);
            a1 = now(); tk[2] += a1 - a0;
            if (run) { a0 = now(); s.be->finish_Optimize(s.ba); a1 = now(); tk[3] += a1 - a0; }
        }
        t0 = now();
        for (auto &s : S) s.fe->end_AddFrame();
        t1 = now(); t[3] += t1 - t0;
    }
    double n = (double)B * steps;
    printf("per stream-frame [us]: begin+prep_track %.2f  fin_track+prep_pose %.2f  fin_pose %.2f  end %.2f   total %.2f\n",
           1e6 * t[0] / n, 1e6 * t[1] / n, 1e6 * t[2] / n, 1e6 * t[3] / n, 1e6 * (t[0] + t[1] + t[2] + t[3]) / n);
    if (nkf) printf("per keyframe [us] (%ld keyframes): fin_detect+prep_right %.1f  fin_right+prep_tri %.1f  fin_tri+prep_ba %.1f  fin_ba %.1f\n", nkf,
                    1e6 * tk[0] / nkf, 1e6 * tk[1] / nkf, 1e6 * tk[2] / nkf, 1e6 * tk[3] / nkf);
    return 0;
}
