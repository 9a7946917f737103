// Data formats either side of the hot path: the KITTI calibration reader of
// Dataset::initialize (reference src/dataset.cpp:24-80) and the two result files of
// VisualOdometry::saveSLAMOutputInFile (reference src/visual_odometry.cpp:198-310): keyframes.txt and landmarks.pcd.
// Pure host code, no device needed.
#include <cmath>
#include <cstdio>
#include <cstring>
#include <string>
#include "../../include/svslam.h"

static void quat_to_R(const double *q, double R[9])
{
    double x = q[0], y = q[1], z = q[2], w = q[3];
    R[0] = 1 - 2 * (y * y + z * z); R[1] = 2 * (x * y - z * w); R[2] = 2 * (x * z + y * w);
    R[3] = 2 * (x * y + z * w); R[4] = 1 - 2 * (x * x + z * z); R[5] = 2 * (y * z - x * w);
    R[6] = 2 * (x * z - y * w); R[7] = 2 * (y * z + x * w); R[8] = 1 - 2 * (x * x + y * y);
}

extern "C" {

// calib.txt: four lines "P<i>: p0 ... p11" (row-major 3x4 projection matrices of the rectified cameras 0..3).
// Like the reference: K = P[:, :3], t = K^-1 P[:, 3], baseline = |t|, then K *= 0.5 when `half` (the reference always
// halves, src/dataset.cpp:73, because it always processes half-resolution images).
int svs_kitti_read_calib(const char *calib_path, int half, double K_out[16] /* 4 x (fx fy cx cy) */, double t_out[12] /* 4 x xyz */,
                         double baseline_out[4])
{
    if (!calib_path || !K_out || !t_out || !baseline_out) return SVS_ERR_ARG;
    FILE *f = fopen(calib_path, "r");
    if (!f) return SVS_ERR_ARG;
    int rc = SVS_OK;
    for (int i = 0; i < 4 && rc == SVS_OK; i++) {
        char name[64];
        double p[12];
        if (fscanf(f, "%63s", name) != 1) { rc = SVS_ERR_ARG; break; }
        for (int j = 0; j < 12; j++) if (fscanf(f, "%lf", &p[j]) != 1) { rc = SVS_ERR_ARG; break; }
        if (rc != SVS_OK) break;
        // K = [[p0 p1 p2], [p4 p5 p6], [p8 p9 p10]] is upper triangular for a rectified camera; general 3x3 inverse anyway
        double a = p[0], b = p[1], c = p[2], d = p[4], e = p[5], g = p[6], h = p[8], k = p[9], l = p[10];
        double det = a * (e * l - g * k) - b * (d * l - g * h) + c * (d * k - e * h);
        if (det == 0.0) { rc = SVS_ERR_ARG; break; }
        double inv[9] = {(e * l - g * k) / det, (c * k - b * l) / det, (b * g - c * e) / det,
                         (g * h - d * l) / det, (a * l - c * h) / det, (c * d - a * g) / det,
                         (d * k - e * h) / det, (b * h - a * k) / det, (a * e - b * d) / det};
        double tv[3] = {p[3], p[7], p[11]};
        for (int r = 0; r < 3; r++) t_out[3 * i + r] = inv[3 * r] * tv[0] + inv[3 * r + 1] * tv[1] + inv[3 * r + 2] * tv[2];
        baseline_out[i] = std::sqrt(t_out[3 * i] * t_out[3 * i] + t_out[3 * i + 1] * t_out[3 * i + 1] + t_out[3 * i + 2] * t_out[3 * i + 2]);
        double s = half ? 0.5 : 1.0;
        K_out[4 * i] = a * s; K_out[4 * i + 1] = e * s; K_out[4 * i + 2] = c * s; K_out[4 * i + 3] = g * s;
    }
    fclose(f);
    return rc;
}

// keyframes.txt (src/visual_odometry.cpp:262-305): dataset directory, left camera index, then one line per keyframe in
// ascending keyframe id: "<frame id> r00 r01 r02 tx r10 r11 r12 ty r20 r21 r22 tz" (T_cw, default ostream precision = %g).
int svs_write_keyframes_txt(const char *path, const char *dataset_dir, int left_cam_index, int n, const int64_t *frame_ids,
                            const double *poses /* 7n: qx qy qz qw tx ty tz */)
{
    if (!path || !dataset_dir || n < 0 || (n > 0 && (!frame_ids || !poses))) return SVS_ERR_ARG;
    FILE *f = fopen(path, "w");
    if (!f) return SVS_ERR_ARG;
    fprintf(f, "%s\n%d\n", dataset_dir, left_cam_index);
    for (int i = 0; i < n; i++) {
        double R[9];
        const double *T = poses + 7 * (size_t)i;
        quat_to_R(T, R);
        fprintf(f, "%lld ", (long long)frame_ids[i]);
        for (int r = 0; r < 3; r++) {
            for (int c = 0; c < 4; c++) {
                double v = c < 3 ? R[3 * r + c] : T[4 + r];
                fprintf(f, "%g%s", v, (r * 4 + c < 11) ? " " : "\n");
            }
        }
    }
    return fclose(f) == 0 ? SVS_OK : SVS_ERR_ARG;
}

// landmarks.pcd (src/visual_odometry.cpp:226-247): what pcl::io::savePCDFileASCII writes for a PointCloud<PointXYZ>
// with height 1 (x y z as float32, precision 8).
int svs_write_landmarks_pcd(const char *path, int n, const double *xyz /* 3n */)
{
    if (!path || n < 0 || (n > 0 && !xyz)) return SVS_ERR_ARG;
    FILE *f = fopen(path, "w");
    if (!f) return SVS_ERR_ARG;
    fprintf(f, "# .PCD v0.7 - Point Cloud Data file format\nVERSION 0.7\nFIELDS x y z\nSIZE 4 4 4\nTYPE F F F\nCOUNT 1 1 1\n"
               "WIDTH %d\nHEIGHT 1\nVIEWPOINT 0 0 0 1 0 0 0\nPOINTS %d\nDATA ascii\n", n, n);
    for (int i = 0; i < n; i++)
        fprintf(f, "%.8g %.8g %.8g\n", (double)(float)xyz[3 * i], (double)(float)xyz[3 * i + 1], (double)(float)xyz[3 * i + 2]);
    return fclose(f) == 0 ? SVS_OK : SVS_ERR_ARG;
}

}  // extern "C"
