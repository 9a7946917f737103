// CPU-only micro-benchmark of the host bookkeeping on the per-frame (non-keyframe) path, with faked GPU results.
#include <chrono>
#include <cstdio>
#include <random>
#include "slam.h"
using namespace slam;
static double now() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }
struct St { Map::Ptr map; std::shared_ptr<Frontend> fe; LkRequest lk; PoseRequest pose; DetectRequest det; TriRequest tri; };
int main(int argc, char **argv)
{
    int B = argc > 1 ? atoi(argv[1]) : 4096, steps = argc > 2 ? atoi(argv[2]) : 30;
    Config cfg; cfg.backend_on = 0;
    double K[4] = {353.5, 353.5, 300.9, 91.6};
    auto cl = std::make_shared<Camera>(K[0], K[1], K[2], K[3], 0.0, SE3());
    auto cr = std::make_shared<Camera>(K[0], K[1], K[2], K[3], 0.537, SE3::fromTranslation(Vec3(-0.537, 0, 0)));
    std::vector<St> S(B);
    std::mt19937 rng(1);
    std::uniform_real_distribution<float> ux(20, 590), uy(20, 160);
    for (auto &s : S) {
        s.map = std::make_shared<Map>(10); s.fe = std::make_shared<Frontend>(cfg); s.fe->SetMap(s.map); s.fe->SetCameras(cl, cr);
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
    double t[4] = {0, 0, 0, 0};
    for (int it = 0; it < steps; it++) {
        double t0 = now();
        for (auto &s : S) { Frame::Ptr f = s.fe->CreateFrame(); s.fe->begin_AddFrame(f, 613, 185); if (s.fe->wants_track()) s.fe->prepare_TrackLastFrame(s.lk); }
        double t1 = now(); t[0] += t1 - t0;
        for (auto &s : S) for (size_t i = 0; i < s.lk.status.size(); i++) { s.lk.next_xy[2 * i] = s.lk.prev_xy[2 * i] + 0.01f; s.lk.next_xy[2 * i + 1] = s.lk.prev_xy[2 * i + 1]; s.lk.status[i] = 1; }
        t0 = now();
        for (auto &s : S) { s.fe->finish_TrackLastFrame(s.lk); s.fe->prepare_EstimateCurrentPose(s.pose); }
        t1 = now(); t[1] += t1 - t0;
        for (auto &s : S) { for (int i = 0; i < 7; i++) s.pose.T[i] = s.pose.T0[i]; s.pose.n_inlier = (int)s.pose.feat_index.size(); }
        t0 = now();
        for (auto &s : S) { s.fe->finish_EstimateCurrentPose(s.pose); }
        t1 = now(); t[2] += t1 - t0;
        t0 = now();
        for (auto &s : S) s.fe->end_AddFrame();
        t1 = now(); t[3] += t1 - t0;
    }
    double n = (double)B * steps;
    printf("per stream-frame [us]: begin+prep_track %.2f  fin_track+prep_pose %.2f  fin_pose %.2f  end %.2f   total %.2f\n",
           1e6 * t[0] / n, 1e6 * t[1] / n, 1e6 * t[2] / n, 1e6 * t[3] / n, 1e6 * (t[0] + t[1] + t[2] + t[3]) / n);
    return 0;
}
