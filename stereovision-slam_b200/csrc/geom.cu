// Geometry kernels in FP64 (rows a4, a5).
//
//   k_triangulate   <- slam::triangulation (reference include/StereoVisionSLAM/algorithm.h:59-86) as called
//                      at src/frontend.cpp:174 and :286 with Camera::pixel2camera (src/camera.cpp:58-72)
//   k_pose_only_lm  <- the g2o block of Frontend::EstimateCurrentPose (src/frontend.cpp:408-527):
//                      1 VertexPose + M EdgeProjectionPoseOnly, Huber(1.0), 4 rounds x optimize(10)
//
// Both are latency-bound per problem (KBs of data); they are batched so that many streams' problems
// run concurrently: one thread per point pair, one warp per pose problem (whole LM loop in-kernel,
// H/b reduced with xor-butterfly shuffles so every lane holds the bitwise-identical 6x6 system).
#include "svs_internal.h"
#include "geom_dev.cuh"
#include <cstring>

// ---------------------------------------------------------------------------------------------
// 4x4 SVD by one-sided Jacobi (Hestenes): rotate column pairs of U = A V until orthogonal.
__global__ void __launch_bounds__(128)
k_triangulate(const float *__restrict__ lxy, const float *__restrict__ rxy, int n, double fxl, double fyl, double cxl,
              double cyl, double fxr, double fyr, double cxr, double cyr, double baseline, double *__restrict__ out_xyz,
              uint8_t *__restrict__ out_ok)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    double x1 = ((double)lxy[2 * i] - cxl) * 1.0 / fxl, y1 = ((double)lxy[2 * i + 1] - cyl) * 1.0 / fyl;
    double x2 = ((double)rxy[2 * i] - cxr) * 1.0 / fxr, y2 = ((double)rxy[2 * i + 1] - cyr) * 1.0 / fyr;
    // rows: x*m2 - m0, y*m2 - m1 for the left [I|0] and right [I|(-b,0,0)] extrinsics
    double U[4][4] = {{-1, 0, x1, 0}, {0, -1, y1, 0}, {-1, 0, x2, baseline}, {0, -1, y2, 0}};
    double V[4][4] = {{1, 0, 0, 0}, {0, 1, 0, 0}, {0, 0, 1, 0}, {0, 0, 0, 1}};
    for (int sweep = 0; sweep < 40; sweep++) {
        double off = 0;
#pragma unroll
        for (int p = 0; p < 3; p++)
#pragma unroll
            for (int q = p + 1; q < 4; q++) {
                double al = 0, be = 0, ga = 0;
#pragma unroll
                for (int r = 0; r < 4; r++) { al += U[r][p] * U[r][p]; be += U[r][q] * U[r][q]; ga += U[r][p] * U[r][q]; }
                if (ga == 0.0) continue;
                double lim = fabs(ga) / sqrt(al * be);
                off = fmax(off, lim);
                if (lim < 1e-16) continue;
                double zeta = (be - al) / (2.0 * ga);
                double t = (zeta >= 0 ? 1.0 : -1.0) / (fabs(zeta) + sqrt(1.0 + zeta * zeta));
                double cs = 1.0 / sqrt(1.0 + t * t), sn = cs * t;
#pragma unroll
                for (int r = 0; r < 4; r++) {
                    double up = U[r][p], uq = U[r][q];
                    U[r][p] = cs * up - sn * uq; U[r][q] = sn * up + cs * uq;
                    double vp = V[r][p], vq = V[r][q];
                    V[r][p] = cs * vp - sn * vq; V[r][q] = sn * vp + cs * vq;
                }
            }
        if (off < 1e-15) break;
    }
    double s[4];
#pragma unroll
    for (int p = 0; p < 4; p++) s[p] = sqrt(U[0][p] * U[0][p] + U[1][p] * U[1][p] + U[2][p] * U[2][p] + U[3][p] * U[3][p]);
    int m1 = 0;                    // smallest
#pragma unroll
    for (int p = 1; p < 4; p++) if (s[p] < s[m1]) m1 = p;
    double s3 = 1e300;             // second smallest
#pragma unroll
    for (int p = 0; p < 4; p++) if (p != m1 && s[p] < s3) s3 = s[p];
    double w = V[3][m1];
    out_xyz[3 * i] = V[0][m1] / w; out_xyz[3 * i + 1] = V[1][m1] / w; out_xyz[3 * i + 2] = V[2][m1] / w;
    out_ok[i] = (s[m1] / s3 < 1e-2) ? 1 : 0;
}

// ---------------------------------------------------------------------------------------------
// 6x6 pivoted LDLT solve (Eigen::LDLT semantics): returns false when a pivot is negative.
__device__ bool ldlt6_solve(double *A /* 36, lower used, destroyed */, const double *b, double *x)
{
    const int n = 6;
    int tr[6];
    double tmp[6];
    int sign = 0;
    for (int k = 0; k < n; k++) {
        int piv = k;
        double best = fabs(A[k * n + k]);
        for (int i = k + 1; i < n; i++) if (fabs(A[i * n + i]) > best) { best = fabs(A[i * n + i]); piv = i; }
        tr[k] = piv;
        if (piv != k) {
            int s = n - piv - 1;
            for (int j = 0; j < k; j++) { double t = A[k * n + j]; A[k * n + j] = A[piv * n + j]; A[piv * n + j] = t; }
            for (int i = 0; i < s; i++) { double t = A[(piv + 1 + i) * n + k]; A[(piv + 1 + i) * n + k] = A[(piv + 1 + i) * n + piv]; A[(piv + 1 + i) * n + piv] = t; }
            { double t = A[k * n + k]; A[k * n + k] = A[piv * n + piv]; A[piv * n + piv] = t; }
            for (int i = k + 1; i < piv; i++) { double t = A[i * n + k]; A[i * n + k] = A[piv * n + i]; A[piv * n + i] = t; }
        }
        int rs = n - k - 1;
        if (k > 0) {
            for (int j = 0; j < k; j++) tmp[j] = A[j * n + j] * A[k * n + j];
            double acc = 0;
            for (int j = 0; j < k; j++) acc += A[k * n + j] * tmp[j];
            A[k * n + k] -= acc;
            for (int i = 0; i < rs; i++) {
                double a2 = 0;
                for (int j = 0; j < k; j++) a2 += A[(k + 1 + i) * n + j] * tmp[j];
                A[(k + 1 + i) * n + k] -= a2;
            }
        }
        double akk = A[k * n + k];
        bool valid = fabs(akk) > 0.0;
        if (k == 0 && !valid) { sign = 0; for (int j = 0; j < n; j++) tr[j] = j; break; }
        if (rs > 0 && valid) for (int i = 0; i < rs; i++) A[(k + 1 + i) * n + k] /= akk;
        if (sign == 1) { if (akk < 0) sign = 2; }
        else if (sign == -1) { if (akk > 0) sign = 2; }
        else if (sign == 0) { if (akk > 0) sign = 1; else if (akk < 0) sign = -1; }
    }
    if (!(sign == 1 || sign == 0)) return false;
    for (int i = 0; i < n; i++) x[i] = b[i];
    for (int k = 0; k < n; k++) if (tr[k] != k) { double t = x[k]; x[k] = x[tr[k]]; x[tr[k]] = t; }
    for (int i = 0; i < n; i++) { double a2 = x[i]; for (int j = 0; j < i; j++) a2 -= A[i * n + j] * x[j]; x[i] = a2; }
    for (int i = 0; i < n; i++) { double d = A[i * n + i]; x[i] = (fabs(d) > DBL_MIN) ? x[i] / d : 0.0; }
    for (int i = n - 1; i >= 0; i--) { double a2 = x[i]; for (int j = i + 1; j < n; j++) a2 -= A[j * n + i] * x[j]; x[i] = a2; }
    for (int k = n - 1; k >= 0; k--) if (tr[k] != k) { double t = x[k]; x[k] = x[tr[k]]; x[tr[k]] = t; }
    return true;
}

// 6x6 LDL^T without pivoting, fully unrolled (register resident).  H is packed upper (21), lambda added to the
// diagonal.  For the symmetric positive definite H + lambda I of the pose problem this gives the same solution as
// Eigen's pivoted LDLT up to rounding; "isPositive" = no negative pivot.
__device__ __forceinline__ bool ldlt6_reg(const double *Hp, double lambda, const double *b, double *x)
{
    double A[6][6];
    {
        int k = 0;
#pragma unroll
        for (int a = 0; a < 6; a++)
#pragma unroll
            for (int c = a; c < 6; c++) { A[c][a] = Hp[k]; k++; }
    }
#pragma unroll
    for (int a = 0; a < 6; a++) A[a][a] += lambda;
    double d[6];
    bool ok = true;
#pragma unroll
    for (int j = 0; j < 6; j++) {
        double dj = A[j][j];
#pragma unroll
        for (int k = 0; k < j; k++) dj -= A[j][k] * A[j][k] * d[k];
        d[j] = dj;
        if (dj < 0.0) ok = false;
        double inv = (fabs(dj) > DBL_MIN) ? 1.0 / dj : 0.0;
#pragma unroll
        for (int i = j + 1; i < 6; i++) {
            double v = A[i][j];
#pragma unroll
            for (int k = 0; k < j; k++) v -= A[i][k] * A[j][k] * d[k];
            A[i][j] = v * inv;
        }
    }
    double y[6];
#pragma unroll
    for (int i = 0; i < 6; i++) {
        double v = b[i];
#pragma unroll
        for (int k = 0; k < i; k++) v -= A[i][k] * y[k];
        y[i] = v;
    }
#pragma unroll
    for (int i = 0; i < 6; i++) y[i] = (fabs(d[i]) > DBL_MIN) ? y[i] / d[i] : 0.0;
#pragma unroll
    for (int i = 5; i >= 0; i--) {
        double v = y[i];
#pragma unroll
        for (int k = i + 1; k < 6; k++) v -= A[k][i] * x[k];
        x[i] = v;
    }
    return ok;
}

// Sum 32 per-lane values across the warp so that EVERY lane ends with all 32 totals (bitwise identical on all lanes):
// recursive-halving reduce-scatter (31 shuffle steps) + an all-gather through shared memory.
__device__ __forceinline__ void warp_allreduce32(double *v /* 32 per lane, in/out */, double *sm /* 32 doubles per warp */, int lane)
{
#pragma unroll
    for (int o = 16, n = 32; o > 0; o >>= 1, n >>= 1) {
        const bool up = (lane & o) != 0;
#pragma unroll
        for (int i = 0; i < 16; i++) {
            if (i < n / 2) {
                double send = up ? v[i] : v[i + n / 2];
                double keep = up ? v[i + n / 2] : v[i];
                v[i] = keep + __shfl_xor_sync(0xffffffffu, send, o);
            }
        }
    }
    // lane now owns the total of value index: bits chosen by the halving order
    int idx = 0;
#pragma unroll
    for (int o = 16, n = 32; o > 0; o >>= 1, n >>= 1) if (lane & o) idx += n / 2;
    __syncwarp();
    sm[idx] = v[0];
    __syncwarp();
#pragma unroll
    for (int i = 0; i < 32; i++) v[i] = sm[i];
    __syncwarp();
}

#define PO_WARPS 4
// One warp per problem.  flags[e]: bit0 = outlier (inactive, level 1).  Stale-error semantics of g2o are kept
// by remembering the pose at which the active edges' errors were last evaluated (Teval).
__global__ void __launch_bounds__(PO_WARPS * 32)
k_pose_only_lm(int n_prob, const int32_t *__restrict__ off, const double *__restrict__ pts_w, const double *__restrict__ uv,
               const double *__restrict__ Kall, const double *__restrict__ T0all, double chi2_th, int rounds, int iters,
               double *__restrict__ T_out, uint8_t *__restrict__ outl, int32_t *__restrict__ n_inlier, svs_lm_stats *__restrict__ stats,
               const int32_t *__restrict__ end /* null: CSR (edges of problem b end at off[b + 1]) */)
{
    __shared__ double s_red[PO_WARPS][32];
    int prob = blockIdx.x * PO_WARPS + (threadIdx.x >> 5);
    if (prob >= n_prob) return;
    int lane = threadIdx.x & 31;
    double *red = s_red[threadIdx.x >> 5];
    int e0 = off[prob], e1 = end ? end[prob] : off[prob + 1];
    double K[4], T0[7], T[7], Teval[7];
#pragma unroll
    for (int i = 0; i < 4; i++) K[i] = Kall[4 * prob + i];
#pragma unroll
    for (int i = 0; i < 7; i++) { T0[i] = T0all[7 * prob + i]; T[i] = T0[i]; Teval[i] = T0[i]; }
    for (int e = e0 + lane; e < e1; e += 32) outl[e] = 0;
    __syncwarp();
    int st_it = 0, st_tr = 0, st_lin = 0, st_sol = 0;
    double st_lambda = 0, st_chi = 0;
    int cnt_out = 0;
    const double huber_delta = 1.0;

    auto robust_chi2 = [&](const double *Tx, bool robust) -> double {
        double acc = 0;
        for (int e = e0 + lane; e < e1; e += 32) {
            if (outl[e]) continue;
            double er[2], pc[3];
            gd::po_error(Tx, K, pts_w + 3 * e, uv + 2 * e, er, pc);
            double e2 = er[0] * er[0] + er[1] * er[1], r0 = e2, r1 = 1.0;
            if (robust) gd::huber(e2, huber_delta, r0, r1);
            acc += r0;
        }
        return gd::warp_sum(acc);
    };

    for (int r = 0; r < rounds; r++) {
        bool robust = r <= 2;   // the kernel is removed after round index 2 (frontend.cpp:518-523)
#pragma unroll
        for (int i = 0; i < 7; i++) T[i] = T0[i];
        int nact = 0;
        for (int e = e0 + lane; e < e1; e += 32) nact += outl[e] ? 0 : 1;
        nact = __reduce_add_sync(0xffffffffu, nact);
        bool evaluated = false;
        if (nact > 0) {
            gd::LmCtl lm = {0.0, 2.0};
            for (int it = 0; it < iters; it++) {
                // ---- computeActiveErrors + buildSystem at T
                double H[21], b[6], cur = 0;
#pragma unroll
                for (int i = 0; i < 21; i++) H[i] = 0;
#pragma unroll
                for (int i = 0; i < 6; i++) b[i] = 0;
                for (int e = e0 + lane; e < e1; e += 32) {
                    if (outl[e]) continue;
                    double er[2], pc[3], J[12];
                    gd::po_error(T, K, pts_w + 3 * e, uv + 2 * e, er, pc);
                    gd::po_jac(K, pc, J);
                    double e2 = er[0] * er[0] + er[1] * er[1], r0 = e2, r1 = 1.0;
                    if (robust) gd::huber(e2, huber_delta, r0, r1);
                    cur += r0;
                    // J = [a0 0 a2 a3 a4 a5; 0 b1 b2 b3 b4 b5] (g2o_types.h:159-162): the structural zeros are skipped,
                    // which leaves every sum unchanged (they would add +-0)
                    const double w0 = r1 * er[0], w1 = r1 * er[1];
                    b[0] -= J[0] * w0; b[1] -= J[7] * w1;
#pragma unroll
                    for (int a = 2; a < 6; a++) b[a] -= J[a] * w0 + J[6 + a] * w1;
                    const double ra0 = r1 * J[0], rb1 = r1 * J[7];
                    H[0] += ra0 * J[0];                                   // (0,0)
#pragma unroll
                    for (int c2 = 2; c2 < 6; c2++) H[c2] += ra0 * J[c2];   // (0,2..5); (0,1) stays 0
                    H[6] += rb1 * J[7];                                   // (1,1)
#pragma unroll
                    for (int c2 = 2; c2 < 6; c2++) H[5 + c2] += rb1 * J[6 + c2];   // (1,2..5)
                    {
                        int k = 11;
#pragma unroll
                        for (int a = 2; a < 6; a++) {
                            const double ra = r1 * J[a], rb = r1 * J[6 + a];
#pragma unroll
                            for (int c2 = a; c2 < 6; c2++) H[k++] += ra * J[c2] + rb * J[6 + c2];
                        }
                    }
                }
                {
                    double v[32];
#pragma unroll
                    for (int i = 0; i < 21; i++) v[i] = H[i];
#pragma unroll
                    for (int i = 0; i < 6; i++) v[21 + i] = b[i];
                    v[27] = cur;
#pragma unroll
                    for (int i = 28; i < 32; i++) v[i] = 0.0;
                    warp_allreduce32(v, red, lane);
#pragma unroll
                    for (int i = 0; i < 21; i++) H[i] = v[i];
#pragma unroll
                    for (int i = 0; i < 6; i++) b[i] = v[21 + i];
                    cur = v[27];
                }
                st_lin++;
                if (it == 0) {
                    double md = 0;
                    int k = 0;
#pragma unroll
                    for (int a = 0; a < 6; a++) { md = fmax(fabs(H[k]), md); k += 6 - a; }
                    lm.lambda = 1e-5 * md; lm.ni = 2;
                }
                double rho = 0;
                int q = 0;
                do {
                    double x[6];
                    bool ok = ldlt6_reg(H, lm.lambda, b, x);
                    st_sol++;
                    if (!ok) { for (int a = 0; a < 6; a++) x[a] = 0; }
                    double Tn[7];
                    gd::se3_oplus(T, x, Tn);
                    double tmp = robust_chi2(Tn, robust);
#pragma unroll
                    for (int i = 0; i < 7; i++) Teval[i] = Tn[i];
                    evaluated = true;
                    if (!ok) tmp = DBL_MAX;
                    rho = cur - tmp;
                    double scale = 0;
#pragma unroll
                    for (int a = 0; a < 6; a++) scale += x[a] * (lm.lambda * x[a] + b[a]);
                    scale += 1e-3;
                    rho /= scale;
                    if (gd::lm_accept(lm, rho, tmp)) {
                        cur = tmp;
#pragma unroll
                        for (int i = 0; i < 7; i++) T[i] = Tn[i];
                    }
                    q++; st_tr++;
                } while (rho < 0 && q < 10);
                st_it++;
                st_lambda = lm.lambda; st_chi = cur;
                if (q == 10 || rho == 0) break;
            }
        }
        // ---- outlier re-classification (frontend.cpp:495-516)
        cnt_out = 0;
        for (int e = e0 + lane; e < e1; e += 32) {
            // active edges carry the error of the last evaluated state, outlier edges are recomputed at T
            const double *Tx = (outl[e] || !evaluated) ? T : Teval;
            double er[2], pc[3];
            gd::po_error(Tx, K, pts_w + 3 * e, uv + 2 * e, er, pc);
            double c2 = er[0] * er[0] + er[1] * er[1];
            if (c2 > chi2_th) { outl[e] = 1; cnt_out++; }
            else outl[e] = 0;
        }
        cnt_out = __reduce_add_sync(0xffffffffu, cnt_out);
        __syncwarp();
    }
    if (lane == 0) {
#pragma unroll
        for (int i = 0; i < 7; i++) T_out[7 * prob + i] = T[i];
        n_inlier[prob] = (e1 - e0) - cnt_out;
        if (stats) {
            stats[prob].iterations = st_it; stats[prob].trials = st_tr; stats[prob].linearizations = st_lin;
            stats[prob].solves = st_sol; stats[prob].lambda = st_lambda; stats[prob].chi2 = st_chi;
        }
    }
}

// Device-pointer entry for the device-resident tracker (track.cu): problem b owns edges [off[b], end[b]).
int svs_i_pose_only_lm_dev(svs_ctx *c, int n_prob, const int32_t *off_dev, const int32_t *end_dev, const double *pts_w_dev,
                           const double *uv_dev, const double *K_dev, const double *T0_dev, double chi2_th, int rounds, int iters,
                           double *T_out_dev, uint8_t *outl_dev, int32_t *n_inlier_dev)
{
    if (n_prob <= 0) return SVS_OK;
    SVS_KERNEL(c, KID_POSE_LM, k_pose_only_lm<<<(n_prob + PO_WARPS - 1) / PO_WARPS, PO_WARPS * 32, 0, c->stream>>>(
        n_prob, off_dev, pts_w_dev, uv_dev, K_dev, T0_dev, chi2_th, rounds, iters, T_out_dev, outl_dev, n_inlier_dev, nullptr, end_dev));
    return SVS_OK;
}

// ---------------------------------------------------------------------------------------------
extern "C" {

int svs_triangulate(svs_ctx *c, const float *left_xy, const float *right_xy, int n, const double Kl[4], const double Kr[4],
                    double baseline, double *out_xyz, uint8_t *out_ok)
{
    if (!c || n < 0 || !Kl || !Kr || (n > 0 && (!left_xy || !right_xy || !out_xyz || !out_ok))) return SVS_ERR_ARG;
    if (n == 0) return SVS_OK;
    SVS_CUDA(c, cudaSetDevice(c->device));
    size_t xy_b = (size_t)n * 8, o_b = (size_t)n * 24;
    SVS_CUDA(c, c->h_in.reserve(2 * xy_b));
    SVS_CUDA(c, c->d_in2.reserve(2 * xy_b));
    SVS_CUDA(c, c->d_out.reserve(o_b + n));
    SVS_CUDA(c, c->h_out.reserve(o_b + n));
    uint8_t *hb = c->h_in.as<uint8_t>(), *db = c->d_in2.as<uint8_t>();
    memcpy(hb, left_xy, xy_b);
    memcpy(hb + xy_b, right_xy, xy_b);
    SVS_CUDA(c, cudaMemcpyAsync(db, hb, 2 * xy_b, cudaMemcpyHostToDevice, c->stream));
    uint8_t *dob = c->d_out.as<uint8_t>();
    SVS_KERNEL(c, KID_TRIANGULATE, k_triangulate<<<(n + 127) / 128, 128, 0, c->stream>>>(reinterpret_cast<float *>(db), reinterpret_cast<float *>(db + xy_b), n,
                                                          Kl[0], Kl[1], Kl[2], Kl[3], Kr[0], Kr[1], Kr[2], Kr[3], baseline,
                                                          reinterpret_cast<double *>(dob), dob + o_b));
    SVS_CUDA(c, cudaMemcpyAsync(c->h_out.p, dob, o_b + n, cudaMemcpyDeviceToHost, c->stream));
    SVS_CUDA(c, svs_i_wait(c));
    memcpy(out_xyz, c->h_out.p, o_b);
    memcpy(out_ok, c->h_out.as<uint8_t>() + o_b, n);
    return SVS_OK;
}

int svs_pose_only_lm(svs_ctx *c, int n_prob, const int32_t *off, const double *pts_w, const double *uv, const double *K,
                     const double *T0, double chi2_th, int rounds, int iters, double *T_out, uint8_t *outlier_out,
                     int32_t *n_inlier, svs_lm_stats *stats)
{
    if (!c || n_prob < 0 || !off || !K || !T0 || !T_out || !n_inlier) return SVS_ERR_ARG;
    if (n_prob == 0) return SVS_OK;
    int M = off[n_prob];
    if (M > 0 && (!pts_w || !uv || !outlier_out)) return SVS_ERR_ARG;
    SVS_CUDA(c, cudaSetDevice(c->device));
    size_t off_b = align_up((size_t)(n_prob + 1) * 4, 16), p_b = (size_t)M * 24, u_b = (size_t)M * 16, k_b = (size_t)n_prob * 32,
           t_b = (size_t)n_prob * 56;
    size_t in_b = off_b + p_b + u_b + k_b + t_b;
    size_t st_b = (size_t)n_prob * sizeof(svs_lm_stats), ni_b = align_up((size_t)n_prob * 4, 8);
    size_t out_b = t_b + st_b + ni_b + M;
    SVS_CUDA(c, c->h_in.reserve(in_b));
    SVS_CUDA(c, c->d_in2.reserve(in_b));
    SVS_CUDA(c, c->d_out.reserve(out_b));
    SVS_CUDA(c, c->h_out.reserve(out_b));
    uint8_t *hb = c->h_in.as<uint8_t>(), *db = c->d_in2.as<uint8_t>();
    memcpy(hb, off, (size_t)(n_prob + 1) * 4);
    if (M) { memcpy(hb + off_b, pts_w, p_b); memcpy(hb + off_b + p_b, uv, u_b); }
    memcpy(hb + off_b + p_b + u_b, K, k_b);
    memcpy(hb + off_b + p_b + u_b + k_b, T0, t_b);
    SVS_CUDA(c, cudaMemcpyAsync(db, hb, in_b, cudaMemcpyHostToDevice, c->stream));
    uint8_t *dob = c->d_out.as<uint8_t>();
    SVS_KERNEL(c, KID_POSE_LM, k_pose_only_lm<<<(n_prob + PO_WARPS - 1) / PO_WARPS, PO_WARPS * 32, 0, c->stream>>>(
        n_prob, reinterpret_cast<int32_t *>(db), reinterpret_cast<double *>(db + off_b), reinterpret_cast<double *>(db + off_b + p_b),
        reinterpret_cast<double *>(db + off_b + p_b + u_b), reinterpret_cast<double *>(db + off_b + p_b + u_b + k_b), chi2_th, rounds,
        iters, reinterpret_cast<double *>(dob), dob + t_b + st_b + ni_b, reinterpret_cast<int32_t *>(dob + t_b + st_b),
        reinterpret_cast<svs_lm_stats *>(dob + t_b), nullptr));
    SVS_CUDA(c, cudaMemcpyAsync(c->h_out.p, dob, out_b, cudaMemcpyDeviceToHost, c->stream));
    SVS_CUDA(c, svs_i_wait(c));
    uint8_t *ho = c->h_out.as<uint8_t>();
    memcpy(T_out, ho, t_b);
    if (stats) memcpy(stats, ho + t_b, st_b);
    memcpy(n_inlier, ho + t_b + st_b, (size_t)n_prob * 4);
    if (M) memcpy(outlier_out, ho + t_b + st_b + ni_b, M);
    return SVS_OK;
}

}  // extern "C"
