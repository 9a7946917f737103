/* CPU oracle — plain-C restatement of the non-OpenCV-image arithmetic on the hot path.
 *
 * TEST INFRASTRUCTURE ONLY: linked/loaded solely by tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline / --impl reference legs, as the checker and the timed
 * CPU baseline.  The product (libsvslam.so) never links or loads this file.
 *
 * What it restates (reference file:line -> function here):
 *   cv::calcOpticalFlowPyrLK call sites  src/frontend.cpp:105-109, :353-357   -> orc_lk_track
 *       (OpenCV is un-vendored; algorithm per SURVEY.md Appendix A.4/A.5; pinned against
 *        cv2 4.13 by tests/test_oracle_cv.py: status identical, positions <= 2e-5 px —
 *        window sums here are exact integers, OpenCV accumulates them in f32 lanes)
 *   VertexPose::oplusImpl                g2o_types.h:40-60                     -> se3_exp / se3_mul
 *   EdgeProjectionPoseOnly               g2o_types.h:117-163                   -> pose-only error/Jacobian
 *   Frontend::EstimateCurrentPose        src/frontend.cpp:408-556              -> orc_pose_only_lm
 *   EdgeProjection::computeError         g2o_types.h:200-216                   -> ba_edge_error
 *   Backend::Optimize (solver part)      src/backend.cpp:22-164                -> orc_ba_optimize
 *
 * g2o / Sophus / Eigen are un-vendored third-party dependencies (README.md:29-32; g2o "0.1"
 * with the post-2017 unique_ptr API, Sophus unversioned, Eigen 3.4.0) and cannot be built or
 * imported in this image, and the reference ships no tests or golden vectors:
 *   ***  PARITY UNPINNED for orc_pose_only_lm and orc_ba_optimize  ***
 * Their semantics follow upstream g2o as restated in SURVEY.md Appendix B (LM with additive
 * lambda, tau = 1e-5, rho-driven schedule, <= 10 trials, Huber, Schur complement, dense
 * pivoted LDLT, central-difference Jacobians with delta = 1e-9 for binary edges).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <float.h>

/* ------------------------------------------------------------------------- */
/* SE3: T = [qx qy qz qw tx ty tz]   (Sophus::SE3d = unit quaternion + translation) */
/* ------------------------------------------------------------------------- */
static void quat_rot(const double *q, const double *p, double *o)
{ /* Eigen QuaternionBase::_transformVector: p + w*uv + qv x uv, uv = 2 (qv x p) */
    double ux = 2.0 * (q[1] * p[2] - q[2] * p[1]);
    double uy = 2.0 * (q[2] * p[0] - q[0] * p[2]);
    double uz = 2.0 * (q[0] * p[1] - q[1] * p[0]);
    o[0] = p[0] + q[3] * ux + (q[1] * uz - q[2] * uy);
    o[1] = p[1] + q[3] * uy + (q[2] * ux - q[0] * uz);
    o[2] = p[2] + q[3] * uz + (q[0] * uy - q[1] * ux);
}
void orc_se3_act(const double *T, const double *p, double *o)
{
    double r[3];
    quat_rot(T, p, r);
    o[0] = r[0] + T[4]; o[1] = r[1] + T[5]; o[2] = r[2] + T[6];
}
/* Camera::world2pixel (src/camera.cpp:74-80) for n points: pixel = K * (ext * (T * p)), the same operations in the same
 * order as one orc_se3_act per transform followed by fx * x / z + cx (batched so that the Python pipeline does not pay one
 * foreign call per feature) */
void orc_world2pixel_batch(const double *T, const double *ext, const double *K, int n, const double *pw, double *uv)
{
    for (int i = 0; i < n; i++) {
        double a[3], c[3];
        orc_se3_act(T, pw + 3 * i, a);
        orc_se3_act(ext, a, c);
        uv[2 * i] = K[0] * c[0] / c[2] + K[2];
        uv[2 * i + 1] = K[1] * c[1] / c[2] + K[3];
    }
}
void orc_se3_mul(const double *A, const double *B, double *C)
{ /* C = A * B ; Sophus SO3 product with first-order renormalisation */
    double ax = A[0], ay = A[1], az = A[2], aw = A[3];
    double bx = B[0], by = B[1], bz = B[2], bw = B[3];
    double q[4];
    q[3] = aw * bw - ax * bx - ay * by - az * bz;
    q[0] = aw * bx + ax * bw + ay * bz - az * by;
    q[1] = aw * by + ay * bw + az * bx - ax * bz;
    q[2] = aw * bz + az * bw + ax * by - ay * bx;
    double n2 = q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3];
    if (n2 != 1.0) {
        double s = 2.0 / (1.0 + n2);
        q[0] *= s; q[1] *= s; q[2] *= s; q[3] *= s;
    }
    double t[3];
    quat_rot(A, B + 4, t);
    C[0] = q[0]; C[1] = q[1]; C[2] = q[2]; C[3] = q[3];
    C[4] = t[0] + A[4]; C[5] = t[1] + A[5]; C[6] = t[2] + A[6];
}
void orc_se3_inv(const double *T, double *O)
{
    double qi[4] = {-T[0], -T[1], -T[2], T[3]};
    double t[3];
    quat_rot(qi, T + 4, t);
    O[0] = qi[0]; O[1] = qi[1]; O[2] = qi[2]; O[3] = qi[3];
    O[4] = -t[0]; O[5] = -t[1]; O[6] = -t[2];
}
void orc_se3_exp(const double *a, double *T)
{ /* a = (upsilon[3], omega[3]); Sophus::SE3d::exp */
    const double *w = a + 3;
    double th2 = w[0] * w[0] + w[1] * w[1] + w[2] * w[2];
    double th = sqrt(th2), imag, real;
    if (th2 < 1e-10 * 1e-10) {
        double th4 = th2 * th2;
        imag = 0.5 - th2 / 48.0 + th4 / 3840.0;
        real = 1.0 - th2 / 8.0 + th4 / 384.0;
    } else {
        double h = 0.5 * th;
        imag = sin(h) / th;
        real = cos(h);
    }
    T[0] = imag * w[0]; T[1] = imag * w[1]; T[2] = imag * w[2]; T[3] = real;
    /* V = I + A*W + B*W^2 */
    double A, B;
    if (th < 1e-10) { A = 0.5; B = 1.0 / 6.0; }
    else { A = (1.0 - cos(th)) / th2; B = (th - sin(th)) / (th2 * th); }
    const double *u = a;
    double wxu[3] = {w[1] * u[2] - w[2] * u[1], w[2] * u[0] - w[0] * u[2], w[0] * u[1] - w[1] * u[0]};
    double wxwxu[3] = {w[1] * wxu[2] - w[2] * wxu[1], w[2] * wxu[0] - w[0] * wxu[2], w[0] * wxu[1] - w[1] * wxu[0]};
    for (int i = 0; i < 3; i++) T[4 + i] = u[i] + A * wxu[i] + B * wxwxu[i];
}
void orc_se3_log(const double *T, double *a)
{ /* Sophus::SE3d::log -> (upsilon, omega) */
    double n2 = T[0] * T[0] + T[1] * T[1] + T[2] * T[2], w = T[3], f, th;
    if (n2 < 1e-10 * 1e-10) {
        f = 2.0 / w - (2.0 / 3.0) * n2 / (w * w * w);
        th = 2.0 * n2 / w; /* not used beyond the small-angle branch */
        th = f * sqrt(n2);
    } else {
        double n = sqrt(n2);
        double at = (w < 0) ? atan2(-n, -w) : atan2(n, w);
        f = 2.0 * at / n;
        th = f * n;
    }
    double om[3] = {f * T[0], f * T[1], f * T[2]};
    double c;
    if (fabs(th) < 1e-10) c = 1.0 / 12.0;
    else { double h = 0.5 * th; c = (1.0 - th * cos(h) / (2.0 * sin(h))) / (th * th); }
    const double *t = T + 4;
    double wxt[3] = {om[1] * t[2] - om[2] * t[1], om[2] * t[0] - om[0] * t[2], om[0] * t[1] - om[1] * t[0]};
    double wxwxt[3] = {om[1] * wxt[2] - om[2] * wxt[1], om[2] * wxt[0] - om[0] * wxt[2], om[0] * wxt[1] - om[1] * wxt[0]};
    for (int i = 0; i < 3; i++) { a[i] = t[i] - 0.5 * wxt[i] + c * wxwxt[i]; a[3 + i] = om[i]; }
}
static void se3_oplus(double *T, const double *upd)
{ /* VertexPose::oplusImpl: T <- exp(upd) * T  (g2o_types.h:59) */
    double E[7], R[7];
    orc_se3_exp(upd, E);
    orc_se3_mul(E, T, R);
    memcpy(T, R, sizeof(R));
}

/* ------------------------------------------------------------------------- */
/* Dense LDLT with diagonal pivoting (Eigen::LDLT as used by g2o::LinearSolverDense) */
/* returns 1 when "isPositive" (no negative pivot) and the solve was done          */
/* ------------------------------------------------------------------------- */
int orc_ldlt_solve(int n, double *A /* n*n row-major, symmetric, destroyed */, const double *b, double *x)
{
    int *tr = (int *)malloc(sizeof(int) * n);
    double *tmp = (double *)malloc(sizeof(double) * n);
    int sign = 0; /* 0 zero, 1 possemidef, -1 negsemidef, 2 indefinite */
    int ok = 1;
#define L(i, j) A[(size_t)(i) * n + (j)]
    for (int k = 0; k < n; k++) {
        int piv = k;
        double best = fabs(L(k, k));
        for (int i = k + 1; i < n; i++) if (fabs(L(i, i)) > best) { best = fabs(L(i, i)); piv = i; }
        tr[k] = piv;
        if (piv != k) { /* symmetric swap on the lower triangle */
            int s = n - piv - 1;
            for (int j = 0; j < k; j++) { double t = L(k, j); L(k, j) = L(piv, j); L(piv, j) = t; }
            for (int i = 0; i < s; i++) { double t = L(piv + 1 + i, k); L(piv + 1 + i, k) = L(piv + 1 + i, piv); L(piv + 1 + i, piv) = t; }
            { double t = L(k, k); L(k, k) = L(piv, piv); L(piv, piv) = t; }
            for (int i = k + 1; i < piv; i++) { double t = L(i, k); L(i, k) = L(piv, i); L(piv, i) = t; }
        }
        int rs = n - k - 1;
        if (k > 0) {
            for (int j = 0; j < k; j++) tmp[j] = L(j, j) * L(k, j);
            double acc = 0;
            for (int j = 0; j < k; j++) acc += L(k, j) * tmp[j];
            L(k, k) -= acc;
            for (int i = 0; i < rs; i++) {
                double a2 = 0;
                for (int j = 0; j < k; j++) a2 += L(k + 1 + i, j) * tmp[j];
                L(k + 1 + i, k) -= a2;
            }
        }
        double akk = L(k, k);
        int valid = fabs(akk) > 0.0;
        if (k == 0 && !valid) { sign = 0; for (int j = 0; j < n; j++) tr[j] = j; break; }
        if (rs > 0 && valid) for (int i = 0; i < rs; i++) L(k + 1 + i, k) /= akk;
        if (sign == 1) { if (akk < 0) sign = 2; }
        else if (sign == -1) { if (akk > 0) sign = 2; }
        else if (sign == 0) { if (akk > 0) sign = 1; else if (akk < 0) sign = -1; }
    }
    if (!(sign == 1 || sign == 0)) ok = 0;
    if (ok) {
        for (int i = 0; i < n; i++) x[i] = b[i];
        for (int k = 0; k < n; k++) if (tr[k] != k) { double t = x[k]; x[k] = x[tr[k]]; x[tr[k]] = t; }
        for (int i = 0; i < n; i++) { double a2 = x[i]; for (int j = 0; j < i; j++) a2 -= L(i, j) * x[j]; x[i] = a2; }
        for (int i = 0; i < n; i++) { double d = L(i, i); x[i] = (fabs(d) > DBL_MIN) ? x[i] / d : 0.0; }
        for (int i = n - 1; i >= 0; i--) { double a2 = x[i]; for (int j = i + 1; j < n; j++) a2 -= L(j, i) * x[j]; x[i] = a2; }
        for (int k = n - 1; k >= 0; k--) if (tr[k] != k) { double t = x[k]; x[k] = x[tr[k]]; x[tr[k]] = t; }
    }
#undef L
    free(tr); free(tmp);
    return ok;
}

/* Huber (g2o::RobustKernelHuber::robustify) on e2 = chi2 */
static void huber(double e2, double delta, double *rho0, double *rho1)
{
    double d2 = delta * delta;
    if (e2 <= d2) { *rho0 = e2; *rho1 = 1.0; }
    else { double s = sqrt(e2); *rho0 = 2 * s * delta - d2; *rho1 = delta / s; }
}

/* ------------------------------------------------------------------------- */
/* Pose-only LM  (Frontend::EstimateCurrentPose, src/frontend.cpp:408-556)    */
/* ------------------------------------------------------------------------- */
static void po_error(const double *T, const double *K, const double *pw, const double *uv, double *e)
{ /* EdgeProjectionPoseOnly::computeError g2o_types.h:117-130 */
    double pc[3];
    orc_se3_act(T, pw, pc);
    double px = K[0] * pc[0] + K[2] * pc[2], py = K[1] * pc[1] + K[3] * pc[2], pz = pc[2];
    e[0] = uv[0] - px / pz;
    e[1] = uv[1] - py / pz;
}
static void po_jac(const double *T, const double *K, const double *pw, double *J /* 2x6 row major */)
{ /* EdgeProjectionPoseOnly::linearizeOplus g2o_types.h:132-163 */
    double pc[3];
    orc_se3_act(T, pw, pc);
    double fx = K[0], fy = K[1], X = pc[0], Y = pc[1], Z = pc[2];
    double Zi = 1.0 / (Z + 1e-18), Zi2 = Zi * Zi;
    J[0] = -fx * Zi; J[1] = 0; J[2] = fx * X * Zi2; J[3] = fx * X * Y * Zi2; J[4] = -fx - fx * X * X * Zi2; J[5] = fx * Y * Zi;
    J[6] = 0; J[7] = -fy * Zi; J[8] = fy * Y * Zi2; J[9] = fy + fy * Y * Y * Zi2; J[10] = -fy * X * Y * Zi2; J[11] = -fy * X * Zi;
}

typedef struct { int iterations, trials, linearizations, solves; double lambda, chi2; } orc_lm_stats;

/* one g2o optimize(max_iter) over the active (level-0) edges; err[] holds per-edge _error */
static void po_optimize(double *T, const double *K, int m, const double *pw, const double *uv,
                        const uint8_t *active, const uint8_t *robust, double huber_delta,
                        int max_iter, double *err, orc_lm_stats *st)
{
    int nact = 0;
    for (int i = 0; i < m; i++) nact += active[i] ? 1 : 0;
    if (nact == 0) return; /* optimize() returns -1: "0 vertices to optimize" */
    double lambda = 0, ni = 2;
    for (int it = 0; it < max_iter; it++) {
        double H[36], b[6], cur = 0;
        memset(H, 0, sizeof(H)); memset(b, 0, sizeof(b));
        for (int i = 0; i < m; i++) if (active[i]) po_error(T, K, pw + 3 * i, uv + 2 * i, err + 2 * i);
        for (int i = 0; i < m; i++) {
            if (!active[i]) continue;
            double *e = err + 2 * i, e2 = e[0] * e[0] + e[1] * e[1], r0 = e2, r1 = 1.0;
            if (robust[i]) huber(e2, huber_delta, &r0, &r1);
            cur += r0;
        }
        for (int i = 0; i < m; i++) {
            if (!active[i]) continue;
            double J[12], *e = err + 2 * i, e2 = e[0] * e[0] + e[1] * e[1], r0 = e2, r1 = 1.0;
            po_jac(T, K, pw + 3 * i, J);
            if (robust[i]) huber(e2, huber_delta, &r0, &r1);
            for (int a = 0; a < 6; a++) {
                b[a] -= r1 * (J[a] * e[0] + J[6 + a] * e[1]);
                for (int c = 0; c < 6; c++) H[a * 6 + c] += r1 * (J[a] * J[c] + J[6 + a] * J[6 + c]);
            }
        }
        st->linearizations++;
        if (it == 0) {
            double md = 0;
            for (int a = 0; a < 6; a++) md = fmax(fabs(H[a * 7]), md);
            lambda = 1e-5 * md; ni = 2;
        }
        double rho = 0;
        int q = 0;
        do {
            double Tb[7], Hd[36], x[6], tmp = 0;
            memcpy(Tb, T, sizeof(Tb));
            memcpy(Hd, H, sizeof(Hd));
            for (int a = 0; a < 6; a++) Hd[a * 7] += lambda;
            int ok = orc_ldlt_solve(6, Hd, b, x);
            st->solves++;
            if (!ok) memset(x, 0, sizeof(x)); /* g2o: x is stale on failure; update still applied — keep zero */
            se3_oplus(T, x);
            for (int i = 0; i < m; i++) if (active[i]) po_error(T, K, pw + 3 * i, uv + 2 * i, err + 2 * i);
            for (int i = 0; i < m; i++) {
                if (!active[i]) continue;
                double *e = err + 2 * i, e2 = e[0] * e[0] + e[1] * e[1], r0 = e2, r1 = 1.0;
                if (robust[i]) huber(e2, huber_delta, &r0, &r1);
                tmp += r0;
            }
            if (!ok) tmp = DBL_MAX;
            rho = cur - tmp;
            double scale = 0;
            for (int a = 0; a < 6; a++) scale += x[a] * (lambda * x[a] + b[a]);
            scale += 1e-3;
            rho /= scale;
            if (rho > 0 && isfinite(tmp)) {
                double alpha = 1.0 - pow(2 * rho - 1, 3);
                alpha = fmin(alpha, 2.0 / 3.0);
                double sf = fmax(1.0 / 3.0, alpha);
                lambda *= sf; ni = 2; cur = tmp;
            } else {
                lambda *= ni; ni *= 2;
                memcpy(T, Tb, sizeof(Tb));
            }
            q++;
            st->trials++;
        } while (rho < 0 && q < 10);
        st->iterations++;
        st->lambda = lambda; st->chi2 = cur;
        if (q == 10 || rho == 0) break;
    }
}

int orc_pose_only_lm(const double *pts_w, const double *uv, int m, const double *K /* fx fy cx cy */,
                     const double *T0, double chi2_th, int rounds, int iters,
                     double *T_out, uint8_t *outlier_out, int *n_inlier, orc_lm_stats *st)
{
    orc_lm_stats local; if (!st) st = &local;
    memset(st, 0, sizeof(*st));
    double *err = (double *)calloc((size_t)2 * (m > 0 ? m : 1), sizeof(double));
    uint8_t *active = (uint8_t *)malloc(m > 0 ? m : 1), *robust = (uint8_t *)malloc(m > 0 ? m : 1);
    uint8_t *outl = (uint8_t *)calloc(m > 0 ? m : 1, 1);
    for (int i = 0; i < m; i++) { active[i] = 1; robust[i] = 1; }
    double T[7];
    memcpy(T, T0, sizeof(T));
    int cnt_out = 0;
    for (int r = 0; r < rounds; r++) {
        memcpy(T, T0, sizeof(T));                         /* frontend.cpp:485 */
        po_optimize(T, K, m, pts_w, uv, active, robust, 1.0, iters, err, st); /* Huber delta 1.0 (default) */
        cnt_out = 0;
        for (int i = 0; i < m; i++) {
            if (outl[i]) po_error(T, K, pts_w + 3 * i, uv + 2 * i, err + 2 * i); /* frontend.cpp:498-501 */
            double c2 = err[2 * i] * err[2 * i] + err[2 * i + 1] * err[2 * i + 1];
            if (c2 > chi2_th) { outl[i] = 1; active[i] = 0; cnt_out++; }
            else { outl[i] = 0; active[i] = 1; }
            if (r == 2) robust[i] = 0;                     /* frontend.cpp:518-523 */
        }
    }
    memcpy(T_out, T, sizeof(T));
    for (int i = 0; i < m; i++) outlier_out[i] = outl[i];
    *n_inlier = m - cnt_out;
    free(err); free(active); free(robust); free(outl);
    return 0;
}

/* ------------------------------------------------------------------------- */
/* Bundle adjustment (Backend::Optimize solver part, src/backend.cpp:22-164)  */
/* ------------------------------------------------------------------------- */
typedef struct {
    int N, L, E;
    double *poses, *lms;
    const int32_t *ekf, *elm; const uint8_t *ecam; const double *euv;
    const double *K[2]; const double *ext[2];
} ba_prob;

static void ba_edge_error(const ba_prob *P, int e, const double *T, const double *p, double *err)
{ /* EdgeProjection::computeError g2o_types.h:200-216 */
    double a[3], c[3];
    int cam = P->ecam[e];
    orc_se3_act(T, p, a);
    orc_se3_act(P->ext[cam], a, c);
    const double *K = P->K[cam];
    double px = K[0] * c[0] + K[2] * c[2], py = K[1] * c[1] + K[3] * c[2], pz = c[2];
    err[0] = P->euv[2 * e] - px / pz;
    err[1] = P->euv[2 * e + 1] - py / pz;
}
static void quat_to_R(const double *q, double *R)
{
    double x = q[0], y = q[1], z = q[2], w = q[3];
    R[0] = 1 - 2 * (y * y + z * z); R[1] = 2 * (x * y - z * w); R[2] = 2 * (x * z + y * w);
    R[3] = 2 * (x * y + z * w); R[4] = 1 - 2 * (x * x + z * z); R[5] = 2 * (y * z - x * w);
    R[6] = 2 * (x * z - y * w); R[7] = 2 * (y * z + x * w); R[8] = 1 - 2 * (x * x + y * y);
}
static void ba_edge_jac(const ba_prob *P, int e, int mode, double *Jp /*2x6*/, double *Jl /*2x3*/)
{
    double *T = P->poses + 7 * P->ekf[e], *p = P->lms + 3 * P->elm[e];
    if (mode == 1) { /* g2o BaseBinaryEdge numeric central differences, delta = 1e-9 */
        const double delta = 1e-9, scalar = 1.0 / (2 * delta);
        for (int d = 0; d < 6; d++) {
            double add[6] = {0, 0, 0, 0, 0, 0}, Tp[7], e1[2], e2[2];
            memcpy(Tp, T, sizeof(Tp)); add[d] = delta; se3_oplus(Tp, add); ba_edge_error(P, e, Tp, p, e1);
            memcpy(Tp, T, sizeof(Tp)); add[d] = -delta; se3_oplus(Tp, add); ba_edge_error(P, e, Tp, p, e2);
            Jp[d] = scalar * (e1[0] - e2[0]); Jp[6 + d] = scalar * (e1[1] - e2[1]);
        }
        for (int d = 0; d < 3; d++) {
            double pp[3], e1[2], e2[2];
            memcpy(pp, p, sizeof(pp)); pp[d] += delta; ba_edge_error(P, e, T, pp, e1);
            memcpy(pp, p, sizeof(pp)); pp[d] += -delta; ba_edge_error(P, e, T, pp, e2);
            Jl[d] = scalar * (e1[0] - e2[0]); Jl[3 + d] = scalar * (e1[1] - e2[1]);
        }
        return;
    }
    /* analytic: e = uv - proj(K (Re (T p) + te)) with T <- exp(d) T */
    int cam = P->ecam[e];
    const double *K = P->K[cam];
    double a[3], c[3], Re[9], R[9];
    orc_se3_act(T, p, a);
    orc_se3_act(P->ext[cam], a, c);
    quat_to_R(P->ext[cam], Re);
    quat_to_R(T, R);
    double fx = K[0], fy = K[1], X = c[0], Y = c[1], Z = c[2], Zi = 1.0 / (Z + 1e-18), Zi2 = Zi * Zi;
    /* de/dc (2x3) */
    double D[6] = {-fx * Zi, 0, fx * X * Zi2, 0, -fy * Zi, fy * Y * Zi2};
    /* dc/d(delta) = Re [I | -[a]x] ; dc/dp = Re R */
    double M[18]; /* 3x6 */
    double ax[9] = {0, -a[2], a[1], a[2], 0, -a[0], -a[1], a[0], 0};
    for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) {
        M[i * 6 + j] = Re[i * 3 + j];
        double s = 0;
        for (int k = 0; k < 3; k++) s += Re[i * 3 + k] * ax[k * 3 + j];
        M[i * 6 + 3 + j] = -s;
    }
    for (int i = 0; i < 2; i++) for (int j = 0; j < 6; j++) {
        double s = 0;
        for (int k = 0; k < 3; k++) s += D[i * 3 + k] * M[k * 6 + j];
        Jp[i * 6 + j] = s;
    }
    double RR[9];
    for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) {
        double s = 0;
        for (int k = 0; k < 3; k++) s += Re[i * 3 + k] * R[k * 3 + j];
        RR[i * 3 + j] = s;
    }
    for (int i = 0; i < 2; i++) for (int j = 0; j < 3; j++) {
        double s = 0;
        for (int k = 0; k < 3; k++) s += D[i * 3 + k] * RR[k * 3 + j];
        Jl[i * 3 + j] = s;
    }
}

static int inv3(const double *A, double *I)
{
    double d = A[0] * (A[4] * A[8] - A[5] * A[7]) - A[1] * (A[3] * A[8] - A[5] * A[6]) + A[2] * (A[3] * A[7] - A[4] * A[6]);
    double id = 1.0 / d;
    I[0] = (A[4] * A[8] - A[5] * A[7]) * id; I[1] = (A[2] * A[7] - A[1] * A[8]) * id; I[2] = (A[1] * A[5] - A[2] * A[4]) * id;
    I[3] = (A[5] * A[6] - A[3] * A[8]) * id; I[4] = (A[0] * A[8] - A[2] * A[6]) * id; I[5] = (A[2] * A[3] - A[0] * A[5]) * id;
    I[6] = (A[3] * A[7] - A[4] * A[6]) * id; I[7] = (A[1] * A[6] - A[0] * A[7]) * id; I[8] = (A[0] * A[4] - A[1] * A[3]) * id;
    return isfinite(id);
}

typedef struct { int iterations, trials, linearizations, solves; double lambda, chi2, chi2_init; } orc_ba_stats;

int orc_ba_optimize(int n_kf, double *poses, int n_lm, double *lms, int n_edge,
                    const int32_t *edge_kf, const int32_t *edge_lm, const uint8_t *edge_cam,
                    const double *edge_uv, const double *K_left, const double *K_right,
                    const double *ext_left, const double *ext_right, double huber_delta,
                    int max_iter, int jac_mode, double *edge_chi2_out, orc_ba_stats *st)
{
    orc_ba_stats local; if (!st) st = &local;
    memset(st, 0, sizeof(*st));
    ba_prob P = {n_kf, n_lm, n_edge, poses, lms, edge_kf, edge_lm, edge_cam, edge_uv, {K_left, K_right}, {ext_left, ext_right}};
    int N = n_kf, L = n_lm, E = n_edge;
    if (E == 0) return 0;
    /* active vertices = those with >= 1 edge; Hessian indices in ascending order */
    int *pidx = (int *)malloc(sizeof(int) * N), *lidx = (int *)malloc(sizeof(int) * (L > 0 ? L : 1));
    for (int i = 0; i < N; i++) pidx[i] = -1;
    for (int i = 0; i < L; i++) lidx[i] = -1;
    for (int e = 0; e < E; e++) { pidx[edge_kf[e]] = 0; lidx[edge_lm[e]] = 0; }
    int NA = 0, LA = 0;
    for (int i = 0; i < N; i++) if (pidx[i] == 0) pidx[i] = NA++;
    for (int i = 0; i < L; i++) if (lidx[i] == 0) lidx[i] = LA++;
    int np = 6 * NA;
    double *Hpp = (double *)malloc(sizeof(double) * np * np), *bp = (double *)malloc(sizeof(double) * np);
    double *Hll = (double *)malloc(sizeof(double) * 9 * LA), *bl = (double *)malloc(sizeof(double) * 3 * LA);
    double *Hpl = (double *)malloc(sizeof(double) * 18 * E);
    double *err = (double *)malloc(sizeof(double) * 2 * E);
    double *S = (double *)malloc(sizeof(double) * np * np), *g = (double *)malloc(sizeof(double) * np);
    double *xp = (double *)malloc(sizeof(double) * np), *xl = (double *)malloc(sizeof(double) * 3 * LA);
    double *Dinv = (double *)malloc(sizeof(double) * 9 * LA);
    double *pb = (double *)malloc(sizeof(double) * 7 * N), *lb = (double *)malloc(sizeof(double) * 3 * (L > 0 ? L : 1));
    /* edges grouped by landmark (creation order preserved inside a landmark) */
    int *lstart = (int *)calloc(LA + 1, sizeof(int)), *lorder = (int *)malloc(sizeof(int) * E);
    for (int e = 0; e < E; e++) lstart[lidx[edge_lm[e]] + 1]++;
    for (int i = 0; i < LA; i++) lstart[i + 1] += lstart[i];
    { int *fill = (int *)calloc(LA, sizeof(int));
      for (int e = 0; e < E; e++) { int l = lidx[edge_lm[e]]; lorder[lstart[l] + fill[l]++] = e; }
      free(fill); }

    double lambda = 0, ni = 2;
    for (int it = 0; it < max_iter; it++) {
        double cur = 0;
        for (int e = 0; e < E; e++) ba_edge_error(&P, e, poses + 7 * edge_kf[e], lms + 3 * edge_lm[e], err + 2 * e);
        for (int e = 0; e < E; e++) {
            double e2 = err[2 * e] * err[2 * e] + err[2 * e + 1] * err[2 * e + 1], r0, r1;
            huber(e2, huber_delta, &r0, &r1);
            cur += r0;
        }
        if (it == 0) st->chi2_init = cur;
        memset(Hpp, 0, sizeof(double) * np * np); memset(bp, 0, sizeof(double) * np);
        memset(Hll, 0, sizeof(double) * 9 * LA); memset(bl, 0, sizeof(double) * 3 * LA);
        for (int e = 0; e < E; e++) {
            double Jp[12], Jl[6], *er = err + 2 * e, e2 = er[0] * er[0] + er[1] * er[1], r0, r1;
            ba_edge_jac(&P, e, jac_mode, Jp, Jl);
            huber(e2, huber_delta, &r0, &r1);
            int pi = pidx[edge_kf[e]], li = lidx[edge_lm[e]];
            for (int a = 0; a < 6; a++) {
                bp[6 * pi + a] -= r1 * (Jp[a] * er[0] + Jp[6 + a] * er[1]);
                for (int c = 0; c < 6; c++) Hpp[(size_t)(6 * pi + a) * np + 6 * pi + c] += r1 * (Jp[a] * Jp[c] + Jp[6 + a] * Jp[6 + c]);
                for (int c = 0; c < 3; c++) Hpl[18 * e + a * 3 + c] = r1 * (Jp[a] * Jl[c] + Jp[6 + a] * Jl[3 + c]);
            }
            for (int a = 0; a < 3; a++) {
                bl[3 * li + a] -= r1 * (Jl[a] * er[0] + Jl[3 + a] * er[1]);
                for (int c = 0; c < 3; c++) Hll[9 * li + a * 3 + c] += r1 * (Jl[a] * Jl[c] + Jl[3 + a] * Jl[3 + c]);
            }
        }
        st->linearizations++;
        if (it == 0) {
            double md = 0;
            for (int a = 0; a < np; a++) md = fmax(fabs(Hpp[(size_t)a * np + a]), md);
            for (int l = 0; l < LA; l++) for (int a = 0; a < 3; a++) md = fmax(fabs(Hll[9 * l + a * 4]), md);
            lambda = 1e-5 * md; ni = 2;
        }
        double rho = 0;
        int q = 0;
        do {
            memcpy(pb, poses, sizeof(double) * 7 * N); memcpy(lb, lms, sizeof(double) * 3 * L);
            /* Schur: S = Hpp + lambda I - sum_l W Dinv W^T ; g = bp - sum_l W Dinv bl */
            memcpy(S, Hpp, sizeof(double) * np * np); memcpy(g, bp, sizeof(double) * np);
            for (int a = 0; a < np; a++) S[(size_t)a * np + a] += lambda;
            int ok = 1;
            for (int l = 0; l < LA; l++) {
                double D[9];
                memcpy(D, Hll + 9 * l, sizeof(D));
                D[0] += lambda; D[4] += lambda; D[8] += lambda;
                if (!inv3(D, Dinv + 9 * l)) ok = 0;
                const double *Di = Dinv + 9 * l;
                for (int s1 = lstart[l]; s1 < lstart[l + 1]; s1++) {
                    int e1 = lorder[s1], p1 = pidx[edge_kf[e1]];
                    double WD[18]; /* W_e1 * Dinv (6x3) */
                    for (int a = 0; a < 6; a++) for (int c = 0; c < 3; c++) {
                        double s = 0;
                        for (int k = 0; k < 3; k++) s += Hpl[18 * e1 + a * 3 + k] * Di[k * 3 + c];
                        WD[a * 3 + c] = s;
                    }
                    for (int a = 0; a < 6; a++) {
                        double s = 0;
                        for (int k = 0; k < 3; k++) s += WD[a * 3 + k] * bl[3 * l + k];
                        g[6 * p1 + a] -= s;
                    }
                    for (int s2 = lstart[l]; s2 < lstart[l + 1]; s2++) {
                        int e2 = lorder[s2], p2 = pidx[edge_kf[e2]];
                        for (int a = 0; a < 6; a++) for (int c = 0; c < 6; c++) {
                            double s = 0;
                            for (int k = 0; k < 3; k++) s += WD[a * 3 + k] * Hpl[18 * e2 + c * 3 + k];
                            S[(size_t)(6 * p1 + a) * np + 6 * p2 + c] -= s;
                        }
                    }
                }
            }
            if (ok) ok = orc_ldlt_solve(np, S, g, xp);
            st->solves++;
            if (!ok) { memset(xp, 0, sizeof(double) * np); memset(xl, 0, sizeof(double) * 3 * LA); }
            else {
                for (int l = 0; l < LA; l++) {
                    double c[3] = {bl[3 * l], bl[3 * l + 1], bl[3 * l + 2]};
                    for (int s1 = lstart[l]; s1 < lstart[l + 1]; s1++) {
                        int e1 = lorder[s1], p1 = pidx[edge_kf[e1]];
                        for (int k = 0; k < 3; k++) {
                            double s = 0;
                            for (int a = 0; a < 6; a++) s += Hpl[18 * e1 + a * 3 + k] * xp[6 * p1 + a];
                            c[k] -= s;
                        }
                    }
                    const double *Di = Dinv + 9 * l;
                    for (int a = 0; a < 3; a++) xl[3 * l + a] = Di[a * 3] * c[0] + Di[a * 3 + 1] * c[1] + Di[a * 3 + 2] * c[2];
                }
            }
            for (int i = 0; i < N; i++) if (pidx[i] >= 0) se3_oplus(poses + 7 * i, xp + 6 * pidx[i]);
            for (int i = 0; i < L; i++) if (lidx[i] >= 0) for (int a = 0; a < 3; a++) lms[3 * i + a] += xl[3 * lidx[i] + a];
            double tmp = 0;
            for (int e = 0; e < E; e++) ba_edge_error(&P, e, poses + 7 * edge_kf[e], lms + 3 * edge_lm[e], err + 2 * e);
            for (int e = 0; e < E; e++) {
                double e2 = err[2 * e] * err[2 * e] + err[2 * e + 1] * err[2 * e + 1], r0, r1;
                huber(e2, huber_delta, &r0, &r1);
                tmp += r0;
            }
            if (!ok) tmp = DBL_MAX;
            rho = cur - tmp;
            double scale = 0;
            for (int a = 0; a < np; a++) scale += xp[a] * (lambda * xp[a] + bp[a]);
            for (int a = 0; a < 3 * LA; a++) scale += xl[a] * (lambda * xl[a] + bl[a]);
            scale += 1e-3;
            rho /= scale;
            if (rho > 0 && isfinite(tmp)) {
                double alpha = 1.0 - pow(2 * rho - 1, 3);
                alpha = fmin(alpha, 2.0 / 3.0);
                double sf = fmax(1.0 / 3.0, alpha);
                lambda *= sf; ni = 2; cur = tmp;
            } else {
                lambda *= ni; ni *= 2;
                memcpy(poses, pb, sizeof(double) * 7 * N); memcpy(lms, lb, sizeof(double) * 3 * L);
            }
            q++;
            st->trials++;
        } while (rho < 0 && q < 10);
        st->iterations++;
        st->lambda = lambda; st->chi2 = cur;
        if (q == 10 || rho == 0) break;
    }
    /* per-edge chi2 as g2o leaves it: from the last computeActiveErrors (possibly a rejected trial) */
    for (int e = 0; e < E; e++) edge_chi2_out[e] = err[2 * e] * err[2 * e] + err[2 * e + 1] * err[2 * e + 1];
    free(pidx); free(lidx); free(Hpp); free(bp); free(Hll); free(bl); free(Hpl); free(err); free(S); free(g);
    free(xp); free(xl); free(Dinv); free(pb); free(lb); free(lstart); free(lorder);
    return 0;
}

/* ------------------------------------------------------------------------- */
/* Pose-graph optimisation: LoopClosure::PoseGraphOptimization, src/loopclosure.cpp:641-799          */
/* VertexPose per keyframe (left-multiplicative update, g2o_types.h:40-60), the vertex of keyframe 0 fixed (:693-696),          */
/* EdgePoseGraph (g2o_types.h:231-267): error = log(M^-1 * v0 * v1^-1) (6-vector, translation first), information I6, no robust */
/* kernel, no linearizeOplus -> g2o numeric central differences (delta 1e-9) for the NON-fixed vertices;                         */
/* BlockSolver<6,6> + LinearSolverDense -> dense pivoted LDLT of the whole 6K x 6K system; optimize(22) (:745-746).              */
/* jac_mode 1 = numeric (the reference), 0 = closed-form Jacobians (SE3 left-Jacobian inverse) for a noise-free comparison.     */
/* ------------------------------------------------------------------------- */
static void pg_error(const double *M, const double *A, const double *B, double *e)
{
    double Mi[7], Bi[7], t1[7], t2[7];
    orc_se3_inv(M, Mi); orc_se3_inv(B, Bi);
    orc_se3_mul(Mi, A, t1); orc_se3_mul(t1, Bi, t2);
    orc_se3_log(t2, e);
}
static void mat3_mul(const double *A, const double *B, double *C)
{ for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) { double s = 0; for (int k = 0; k < 3; k++) s += A[i * 3 + k] * B[k * 3 + j]; C[i * 3 + j] = s; } }
static void hat3(const double *v, double *M) { M[0] = 0; M[1] = -v[2]; M[2] = v[1]; M[3] = v[2]; M[4] = 0; M[5] = -v[0]; M[6] = -v[1]; M[7] = v[0]; M[8] = 0; }
/* inverse of the SE3 left Jacobian at xi = (rho, phi) (Barfoot, State Estimation for Robotics, eqs. 7.85-7.95), 6x6 row-major */
void orc_se3_left_jac_inv(const double *xi, double *Ji)
{
    const double *rho = xi, *phi = xi + 3;
    double th2 = phi[0] * phi[0] + phi[1] * phi[1] + phi[2] * phi[2], th = sqrt(th2);
    double P[9], R[9], PP[9], PR[9], RP[9], PRP[9], PPR[9], RPP[9], PRPP[9], PPRP[9];
    hat3(phi, P); hat3(rho, R);
    mat3_mul(P, P, PP); mat3_mul(P, R, PR); mat3_mul(R, P, RP); mat3_mul(PR, P, PRP); mat3_mul(PP, R, PPR); mat3_mul(RP, P, RPP);
    mat3_mul(PRP, P, PRPP); mat3_mul(PP, RP, PPRP);
    double Jinv[9], Q[9];
    double c1, c2, c3, a;    /* Q coefficients; a = coefficient of phi^ phi^ in J^-1 */
    if (th < 1e-5) {
        c1 = 1.0 / 6.0 - th2 / 120.0; c2 = 1.0 / 24.0 - th2 / 720.0; c3 = 1.0 / 120.0 - th2 / 2520.0;
        a = 1.0 / 12.0 + th2 / 720.0;
    } else {
        double s = sin(th), c = cos(th), th3 = th2 * th, th4 = th2 * th2, th5 = th4 * th;
        c1 = (th - s) / th3; c2 = (1.0 - 0.5 * th2 - c) / th4; c3 = 0.5 * ((1.0 - 0.5 * th2 - c) / th4 - 3.0 * (th - s - th3 / 6.0) / th5);
        a = (1.0 - 0.5 * th * s / (1.0 - c)) / th2;      /* 1/th^2 (1 - th/2 cot(th/2)) */
    }
    for (int i = 0; i < 9; i++) {
        double I = (i % 4 == 0) ? 1.0 : 0.0;
        Jinv[i] = I - 0.5 * P[i] + a * PP[i];
        Q[i] = 0.5 * R[i] + c1 * (PR[i] + RP[i] + PRP[i]) - c2 * (PPR[i] + RPP[i] - 3.0 * PRP[i]) - c3 * (PRPP[i] + PPRP[i]);
    }
    double JQ[9], JQJ[9];
    mat3_mul(Jinv, Q, JQ); mat3_mul(JQ, Jinv, JQJ);
    for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) {
        Ji[i * 6 + j] = Jinv[i * 3 + j]; Ji[i * 6 + 3 + j] = -JQJ[i * 3 + j];
        Ji[(3 + i) * 6 + j] = 0.0; Ji[(3 + i) * 6 + 3 + j] = Jinv[i * 3 + j];
    }
}
/* Ja = d e / d delta_a, Jb = d e / d delta_b for the left-multiplicative vertex updates, 6x6 row-major each */
static void pg_jac(const double *M, const double *A, const double *B, int mode, double *Ja, double *Jb)
{
    if (mode == 1) {
        const double delta = 1e-9, scalar = 1.0 / (2 * delta);
        for (int v = 0; v < 2; v++)
            for (int d = 0; d < 6; d++) {
                double add[6] = {0, 0, 0, 0, 0, 0}, Tp[7], e1[6], e2[6];
                const double *T = v ? B : A;
                add[d] = delta; memcpy(Tp, T, sizeof(Tp)); se3_oplus(Tp, add); pg_error(M, v ? A : Tp, v ? Tp : B, e1);
                add[d] = -delta; memcpy(Tp, T, sizeof(Tp)); se3_oplus(Tp, add); pg_error(M, v ? A : Tp, v ? Tp : B, e2);
                for (int r = 0; r < 6; r++) (v ? Jb : Ja)[r * 6 + d] = scalar * (e1[r] - e2[r]);
            }
        return;
    }
    /* e(da) = log(M^-1 exp(da) A B^-1) = log(exp(Ad(M^-1) da) E) ~ e0 + Jl^-1(e0) Ad(M^-1) da
       e(db) = log(M^-1 A B^-1 exp(-db)) = log(E exp(-db))         ~ e0 - Jr^-1(e0) db,  Jr^-1(e) = Jl^-1(-e) */
    double e0[6], me0[6], Jl[36], Jr[36], Mi[7], Rm[9], tx[9], tR[9], Ad[36];
    pg_error(M, A, B, e0);
    for (int i = 0; i < 6; i++) me0[i] = -e0[i];
    orc_se3_left_jac_inv(e0, Jl); orc_se3_left_jac_inv(me0, Jr);
    orc_se3_inv(M, Mi); quat_to_R(Mi, Rm); hat3(Mi + 4, tx); mat3_mul(tx, Rm, tR);
    for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) {
        Ad[i * 6 + j] = Rm[i * 3 + j]; Ad[i * 6 + 3 + j] = tR[i * 3 + j]; Ad[(3 + i) * 6 + j] = 0.0; Ad[(3 + i) * 6 + 3 + j] = Rm[i * 3 + j];
    }
    for (int i = 0; i < 6; i++) for (int j = 0; j < 6; j++) {
        double s = 0;
        for (int k = 0; k < 6; k++) s += Jl[i * 6 + k] * Ad[k * 6 + j];
        Ja[i * 6 + j] = s; Jb[i * 6 + j] = -Jr[i * 6 + j];
    }
}
void orc_pg_edge_jac(const double *M, const double *A, const double *B, int mode, double *Ja, double *Jb) { pg_jac(M, A, B, mode, Ja, Jb); }

int orc_pose_graph_optimize(int n_kf, double *poses /* 7*n in/out */, const uint8_t *fixed /* n */, int n_edge, const int32_t *edge_a,
                            const int32_t *edge_b, const double *meas /* 7*n_edge */, int max_iter, int jac_mode, orc_ba_stats *st)
{
    orc_ba_stats local; if (!st) st = &local;
    memset(st, 0, sizeof(*st));
    int N = n_kf, E = n_edge;
    if (E == 0) return 0;
    /* active = non-fixed vertices with >= 1 edge, ascending id */
    int *idx = (int *)malloc(sizeof(int) * N);
    for (int i = 0; i < N; i++) idx[i] = -1;
    for (int e = 0; e < E; e++) { if (!fixed[edge_a[e]]) idx[edge_a[e]] = 0; if (!fixed[edge_b[e]]) idx[edge_b[e]] = 0; }
    int NA = 0;
    for (int i = 0; i < N; i++) if (idx[i] == 0) idx[i] = NA++;
    int n = 6 * NA;
    if (n == 0) { free(idx); return 0; }
    double *H = (double *)malloc(sizeof(double) * n * n), *b = (double *)malloc(sizeof(double) * n), *S = (double *)malloc(sizeof(double) * n * n);
    double *x = (double *)malloc(sizeof(double) * n), *pb = (double *)malloc(sizeof(double) * 7 * N), *err = (double *)malloc(sizeof(double) * 6 * E);
    double lambda = 0, ni = 2, cur = 0;
    for (int it = 0; it < max_iter; it++) {
        cur = 0;
        for (int e = 0; e < E; e++) {
            pg_error(meas + 7 * e, poses + 7 * edge_a[e], poses + 7 * edge_b[e], err + 6 * e);
            for (int r = 0; r < 6; r++) cur += err[6 * e + r] * err[6 * e + r];
        }
        if (it == 0) st->chi2_init = cur;
        memset(H, 0, sizeof(double) * n * n); memset(b, 0, sizeof(double) * n);
        for (int e = 0; e < E; e++) {
            double Ja[36], Jb[36];
            pg_jac(meas + 7 * e, poses + 7 * edge_a[e], poses + 7 * edge_b[e], jac_mode, Ja, Jb);
            int ia = idx[edge_a[e]], ib = idx[edge_b[e]];
            const double *J[2] = {Ja, Jb};
            int id[2] = {ia, ib};
            for (int u = 0; u < 2; u++) {
                if (id[u] < 0) continue;
                for (int r = 0; r < 6; r++) {
                    double s = 0;
                    for (int k = 0; k < 6; k++) s += J[u][k * 6 + r] * err[6 * e + k];
                    b[6 * id[u] + r] -= s;
                }
                for (int v = 0; v < 2; v++) {
                    if (id[v] < 0) continue;
                    for (int r = 0; r < 6; r++) for (int c = 0; c < 6; c++) {
                        double s = 0;
                        for (int k = 0; k < 6; k++) s += J[u][k * 6 + r] * J[v][k * 6 + c];
                        H[(size_t)(6 * id[u] + r) * n + 6 * id[v] + c] += s;
                    }
                }
            }
        }
        st->linearizations++;
        if (it == 0) {
            double md = 0;
            for (int i = 0; i < n; i++) md = fmax(md, fabs(H[(size_t)i * n + i]));
            lambda = 1e-5 * md; ni = 2;
        }
        double rho = 0;
        int q = 0;
        do {
            memcpy(pb, poses, sizeof(double) * 7 * N);
            memcpy(S, H, sizeof(double) * n * n);
            for (int i = 0; i < n; i++) S[(size_t)i * n + i] += lambda;
            int ok = orc_ldlt_solve(n, S, b, x);
            st->solves++;
            if (!ok) memset(x, 0, sizeof(double) * n);
            for (int i = 0; i < N; i++) if (idx[i] >= 0) se3_oplus(poses + 7 * i, x + 6 * idx[i]);
            double tmp = 0;
            for (int e = 0; e < E; e++) {
                pg_error(meas + 7 * e, poses + 7 * edge_a[e], poses + 7 * edge_b[e], err + 6 * e);
                for (int r = 0; r < 6; r++) tmp += err[6 * e + r] * err[6 * e + r];
            }
            if (!ok) tmp = DBL_MAX;
            rho = cur - tmp;
            double scale = 1e-3;
            for (int a = 0; a < n; a++) scale += x[a] * (lambda * x[a] + b[a]);
            rho /= scale;
            if (rho > 0 && isfinite(tmp)) {
                double alpha = 1.0 - pow(2 * rho - 1, 3);
                alpha = fmin(alpha, 2.0 / 3.0);
                lambda *= fmax(1.0 / 3.0, alpha); ni = 2; cur = tmp;
            } else {
                lambda *= ni; ni *= 2;
                memcpy(poses, pb, sizeof(double) * 7 * N);
            }
            q++; st->trials++;
        } while (rho < 0 && q < 10);
        st->iterations++;
        st->lambda = lambda; st->chi2 = cur;
        if (q == 10 || rho == 0) break;
    }
    free(idx); free(H); free(b); free(S); free(x); free(pb); free(err);
    return 0;
}

/* ------------------------------------------------------------------------- */
/* Pyramidal LK (cv::calcOpticalFlowPyrLK as called at frontend.cpp:105,353) */
/* ------------------------------------------------------------------------- */
static inline int refl101(int i, int n)
{
    if (n == 1) return 0;
    while (i < 0 || i >= n) { if (i < 0) i = -i; else i = 2 * (n - 1) - i; }
    return i;
}
typedef struct { const uint8_t *p; int w, h, stride; } orc_img;
static inline int px(const orc_img *im, int x, int y) { return im->p[(size_t)refl101(y, im->h) * im->stride + refl101(x, im->w)]; }
/* Scharr derivative (calcScharrDeriv); zero outside the image (constant border) */
static inline void scharr(const orc_img *im, int x, int y, int *dx, int *dy)
{
    if (x < 0 || y < 0 || x >= im->w || y >= im->h) { *dx = 0; *dy = 0; return; }
    int v[3][3];
    for (int j = 0; j < 3; j++) for (int i = 0; i < 3; i++) v[j][i] = px(im, x + i - 1, y + j - 1);
    int s0[3], s1[3];
    for (int i = 0; i < 3; i++) { s0[i] = (v[0][i] + v[2][i]) * 3 + v[1][i] * 10; s1[i] = v[2][i] - v[0][i]; }
    *dx = s0[2] - s0[0];
    *dy = (s1[0] + s1[2]) * 3 + s1[1] * 10;
}
static inline int descale(int x, int n) { return (x + (1 << (n - 1))) >> n; }
static inline int cv_round_f(float v) { return (int)lrintf(v); }
static inline int cv_floor_f(float v) { int i = (int)v; return i - (v < (float)i); }

/* levels[l] for l = 0..n_levels-1 (n_levels = maxLevel+1 actually built) */
int orc_lk_track(const orc_img *prev, const orc_img *next, int n_levels, const float *prev_xy,
                 float *next_xy /* in: initial flow, out */, int n, int win, int max_iter, double eps,
                 uint8_t *status, int *iters_out /* optional, total iterations per point */)
{
    if (max_iter > 100) max_iter = 100; if (max_iter < 0) max_iter = 0;
    if (eps < 0) eps = 0; if (eps > 10) eps = 10;
    double eps2 = eps * eps;
    const float half = (win - 1) * 0.5f;
    const int W_BITS = 14;
    const float FLT_SCALE = 1.f / (1 << 20);
    int *Iw = (int *)malloc(sizeof(int) * win * win * 3);
    for (int i = 0; i < n; i++) { status[i] = 1; if (iters_out) iters_out[i] = 0; }
    for (int level = n_levels - 1; level >= 0; level--) {
        const orc_img *I = &prev[level], *J = &next[level];
        float lscale = (float)(1. / (1 << level));
        for (int pt = 0; pt < n; pt++) {
            float ppx = prev_xy[2 * pt] * lscale, ppy = prev_xy[2 * pt + 1] * lscale, nx, ny;
            if (level == n_levels - 1) { nx = next_xy[2 * pt] * lscale; ny = next_xy[2 * pt + 1] * lscale; }
            else { nx = next_xy[2 * pt] * 2.f; ny = next_xy[2 * pt + 1] * 2.f; }
            next_xy[2 * pt] = nx; next_xy[2 * pt + 1] = ny;
            ppx -= half; ppy -= half;
            int ix = cv_floor_f(ppx), iy = cv_floor_f(ppy);
            if (ix < -win || ix >= I->w || iy < -win || iy >= I->h) { if (level == 0) status[pt] = 0; continue; }
            float a = ppx - ix, b = ppy - iy;
            int w00 = cv_round_f((1.f - a) * (1.f - b) * (1 << W_BITS));
            int w01 = cv_round_f(a * (1.f - b) * (1 << W_BITS));
            int w10 = cv_round_f((1.f - a) * b * (1 << W_BITS));
            int w11 = (1 << W_BITS) - w00 - w01 - w10;
            int64_t sA11 = 0, sA12 = 0, sA22 = 0;
            for (int y = 0; y < win; y++) for (int x = 0; x < win; x++) {
                int X = ix + x, Y = iy + y;
                int ival = descale(px(I, X, Y) * w00 + px(I, X + 1, Y) * w01 + px(I, X, Y + 1) * w10 + px(I, X + 1, Y + 1) * w11, W_BITS - 5);
                int d00x, d00y, d01x, d01y, d10x, d10y, d11x, d11y;
                scharr(I, X, Y, &d00x, &d00y); scharr(I, X + 1, Y, &d01x, &d01y);
                scharr(I, X, Y + 1, &d10x, &d10y); scharr(I, X + 1, Y + 1, &d11x, &d11y);
                int ixv = descale(d00x * w00 + d01x * w01 + d10x * w10 + d11x * w11, W_BITS);
                int iyv = descale(d00y * w00 + d01y * w01 + d10y * w10 + d11y * w11, W_BITS);
                int *o = Iw + 3 * (y * win + x);
                o[0] = (short)ival; o[1] = (short)ixv; o[2] = (short)iyv;
                sA11 += (int64_t)o[1] * o[1]; sA12 += (int64_t)o[1] * o[2]; sA22 += (int64_t)o[2] * o[2];
            }
            float A11 = (float)sA11 * FLT_SCALE, A12 = (float)sA12 * FLT_SCALE, A22 = (float)sA22 * FLT_SCALE;
            float D = A11 * A22 - A12 * A12;
            float minEig = (A22 + A11 - sqrtf((A11 - A22) * (A11 - A22) + 4.f * A12 * A12)) / (float)(2 * win * win);
            if ((double)minEig < 1e-4 || D < FLT_EPSILON) { if (level == 0) status[pt] = 0; continue; }
            D = 1.f / D;
            nx -= half; ny -= half;
            float pdx = 0, pdy = 0;
            for (int j = 0; j < max_iter; j++) {
                int jx = cv_floor_f(nx), jy = cv_floor_f(ny);
                if (jx < -win || jx >= J->w || jy < -win || jy >= J->h) { if (level == 0) status[pt] = 0; break; }
                if (iters_out) iters_out[pt]++;
                a = nx - jx; b = ny - jy;
                w00 = cv_round_f((1.f - a) * (1.f - b) * (1 << W_BITS));
                w01 = cv_round_f(a * (1.f - b) * (1 << W_BITS));
                w10 = cv_round_f((1.f - a) * b * (1 << W_BITS));
                w11 = (1 << W_BITS) - w00 - w01 - w10;
                int64_t sb1 = 0, sb2 = 0;
                for (int y = 0; y < win; y++) for (int x = 0; x < win; x++) {
                    int X = jx + x, Y = jy + y;
                    const int *o = Iw + 3 * (y * win + x);
                    int diff = descale(px(J, X, Y) * w00 + px(J, X + 1, Y) * w01 + px(J, X, Y + 1) * w10 + px(J, X + 1, Y + 1) * w11, W_BITS - 5) - o[0];
                    sb1 += (int64_t)diff * o[1]; sb2 += (int64_t)diff * o[2];
                }
                float b1 = (float)sb1 * FLT_SCALE, b2 = (float)sb2 * FLT_SCALE;
                float dx = (float)((A12 * b2 - A22 * b1) * D), dy = (float)((A12 * b1 - A11 * b2) * D);
                nx += dx; ny += dy;
                next_xy[2 * pt] = nx + half; next_xy[2 * pt + 1] = ny + half;
                if ((double)dx * dx + (double)dy * dy <= eps2) break;
                if (j > 0 && fabs((double)(dx + pdx)) < 0.01 && fabs((double)(dy + pdy)) < 0.01) {
                    next_xy[2 * pt] -= dx * 0.5f; next_xy[2 * pt + 1] -= dy * 0.5f;
                    break;
                }
                pdx = dx; pdy = dy;
            }
            if (status[pt] && level == 0) { /* the err pass re-checks the final position */
                float fx = next_xy[2 * pt] - half, fy = next_xy[2 * pt + 1] - half;
                int jx = cv_floor_f(fx), jy = cv_floor_f(fy);
                if (jx < -win || jx >= J->w || jy < -win || jy >= J->h) status[pt] = 0;
            }
        }
    }
    free(Iw);
    return 0;
}
