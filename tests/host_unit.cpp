// CPU-only unit tests of the host mirrors (stereovision-slam_b200/host/slam.{h,cpp}): containers, SE3 algebra, map window.
// Built and run by tests/test_host_cpp.py with g++ (no CUDA needed).
#include <cassert>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include "slam.h"
using namespace slam;

#define CHECK(c) do { if (!(c)) { std::fprintf(stderr, "CHECK failed %s:%d: %s\n", __FILE__, __LINE__, #c); std::exit(1); } } while (0)

static void test_obslist()
{
    Frame f[3];
    ObsList l;
    CHECK(l.empty());
    for (int i = 0; i < 11; i++) l.push_back(Observation{&f[i % 3], i % 2 == 0, i});     // grows past the in-place capacity
    CHECK(l.size() == 11);
    int k = 0;
    for (const Observation &o : l) { CHECK(o.index == k && o.frame == &f[k % 3]); k++; }
    l.erase(l.begin() + 4);                                                               // order of the rest is kept
    CHECK(l.size() == 10);
    int want[10] = {0, 1, 2, 3, 5, 6, 7, 8, 9, 10};
    k = 0;
    for (const Observation &o : l) CHECK(o.index == want[k++]);
    while (!l.empty()) l.erase(l.begin());
    CHECK(l.size() == 0);
    l.push_back(Observation{&f[0], true, 7});
    CHECK(l.size() == 1 && l.begin()->index == 7);
}

static void test_landmark_map()
{
    MapPoint mp[6];
    LandmarkMap m;
    for (int i : {0, 2, 5}) { mp[i].id_ = i; m.insert_or_assign(i, &mp[i]); }             // appends
    m.insert_or_assign(3, &mp[3]); m.insert_or_assign(1, &mp[1]);                          // out-of-order inserts keep ascending ids
    m.insert_or_assign(2, &mp[4]);                                                         // assign
    unsigned long ids[5] = {0, 1, 2, 3, 5};
    int k = 0;
    for (const auto &kv : m) CHECK(kv.first == ids[k++]);
    CHECK(m.size() == 5);
    for (const auto &kv : m) if (kv.first == 2) CHECK(kv.second == &mp[4]);
    m.erase_if([](const LandmarkMap::value_type &kv) { return kv.first % 2 == 1; });
    CHECK(m.size() == 2);
    k = 0;
    unsigned long ev[2] = {0, 2};
    for (const auto &kv : m) CHECK(kv.first == ev[k++]);
}

static void test_se3()
{
    double tg[6] = {0.3, -0.2, 0.5, 0.02, -0.03, 0.04};
    SE3 T = SE3::exp(tg), I = T * T.inverse();
    for (int i = 0; i < 3; i++) CHECK(std::fabs(I.d[i]) < 1e-15 && std::fabs(I.d[4 + i]) < 1e-15);
    CHECK(std::fabs(I.d[3] - 1.0) < 1e-15);
    double back[6];
    T.log(back);
    for (int i = 0; i < 6; i++) CHECK(std::fabs(back[i] - tg[i]) < 1e-12);
    // a pure translation: the short-cut in Camera::world2camera is bit-identical to the full action
    SE3 tr = SE3::fromTranslation(Vec3(-0.537, 0, 0));
    CHECK(tr.rotation_is_identity() && !T.rotation_is_identity());
    Camera cam(353.5, 353.5, 300.9, 91.6, 0.537, tr), cam_rot(353.5, 353.5, 300.9, 91.6, 0.537, T);
    CHECK(cam.pure_translation_ && !cam_rot.pure_translation_);
    Vec3 p(1.25, -0.75, 12.5);
    Vec3 a = cam.world2camera(p, T), b = tr * (T * p);
    CHECK(a.x == b.x && a.y == b.y && a.z == b.z);
    Vec3 c = cam_rot.world2camera(p, T), d = T * (T * p);
    CHECK(c.x == d.x && c.y == d.y && c.z == d.z);
    Vec2 px = cam.world2pixel(p, SE3());
    CHECK(std::fabs(px.x - (353.5 * (1.25 - 0.537) / 12.5 + 300.9)) < 1e-12);
}

static void test_map_window()
{
    Map map(3);
    Frame::Ptr kf[6];
    std::vector<MapPoint *> pts;
    for (int k = 0; k < 6; k++) {
        kf[k] = std::make_shared<Frame>();
        kf[k]->id_ = 10 * k; kf[k]->keyframe_id_ = k; kf[k]->is_keyframe_ = true;
        kf[k]->SetPose(SE3::fromTranslation(Vec3(0, 0, -1.0 * k)));                        // 1 m apart
        // every keyframe observes two new landmarks (left + right) and the previous keyframe's landmarks (left)
        for (int j = 0; j < 2; j++) {
            MapPoint *mp = map.CreateNewMappoint();
            mp->SetPos(Vec3(j, 0, 5.0 + k));
            Feature fl, fr; fl.map_point_ = (long)mp->id_; fr.map_point_ = (long)mp->id_; fr.is_on_left_image_ = false;
            kf[k]->feature_left_.push_back(fl); kf[k]->feature_right_.push_back(fr);
            int idx = (int)kf[k]->feature_left_.size() - 1;
            mp->AddObservation(Observation{kf[k].get(), true, idx});
            mp->AddObservation(Observation{kf[k].get(), false, idx});
            map.InsertMapPoint(mp);
            pts.push_back(mp);
        }
        if (k > 0) for (int j = 0; j < 2; j++) {
            MapPoint *mp = pts[2 * (k - 1) + j];
            Feature fl; fl.map_point_ = (long)mp->id_;
            kf[k]->feature_left_.push_back(fl);
            Feature none; none.valid = false; none.is_on_left_image_ = false;
            kf[k]->feature_right_.push_back(none);
            mp->AddObservation(Observation{kf[k].get(), true, (int)kf[k]->feature_left_.size() - 1});
        }
        map.InsertKeyFrame(kf[k]);
        CHECK((int)map.GetActiveKeyFrames().size() == std::min(k + 1, 3));
    }
    // src/map.cpp:76-181: no keyframe is closer than 0.2 to the newest, so the FARTHEST one is evicted each time
    unsigned long want_kf[3] = {3, 4, 5};
    int i = 0;
    for (const auto &kv : map.GetActiveKeyFrames()) CHECK(kv.first == want_kf[i++]);
    CHECK(map.GetAllKeyFrames().size() == 6 && map.GetAllMapPoints().size() == 12);
    // landmarks only seen by evicted keyframes lost all observations and left the active set (CleanMap)
    for (const auto &kv : map.GetActiveMapPoints()) CHECK(kv.second->observed_times_ > 0 && kv.first >= 4);
    CHECK(map.GetActiveMapPoints().size() == 8);       // landmarks 4..11
    CHECK(pts[0]->observed_times_ == 0 && pts[4]->observed_times_ == 1 && pts[6]->observed_times_ == 3);
    CHECK(map.GetMapPoint(7) == pts[7] && map.GetMapPoint(-1) == nullptr);
}

int main()
{
    test_obslist();
    test_landmark_map();
    test_se3();
    test_map_window();
    std::puts("host unit tests ok");
    return 0;
}
