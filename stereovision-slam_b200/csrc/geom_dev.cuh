// FP64 device helpers shared by the geometry kernels: SE3 (Sophus::SE3d semantics: unit quaternion +
// translation, T = [qx qy qz qw tx ty tz]), Huber kernel (g2o::RobustKernelHuber), projection
// residuals and Jacobians of the reference's two edge types (include/StereoVisionSLAM/g2o_types.h).
#pragma once
#include <cuda_runtime.h>
#include <cfloat>

namespace gd {

__device__ __forceinline__ void quat_rot(const double *q, const double *p, double *o)
{   // Eigen QuaternionBase::_transformVector
    double ux = 2.0 * (q[1] * p[2] - q[2] * p[1]);
    double uy = 2.0 * (q[2] * p[0] - q[0] * p[2]);
    double uz = 2.0 * (q[0] * p[1] - q[1] * p[0]);
    o[0] = p[0] + q[3] * ux + (q[1] * uz - q[2] * uy);
    o[1] = p[1] + q[3] * uy + (q[2] * ux - q[0] * uz);
    o[2] = p[2] + q[3] * uz + (q[0] * uy - q[1] * ux);
}
__device__ __forceinline__ void se3_act(const double *T, const double *p, double *o)
{
    double r[3];
    quat_rot(T, p, r);
    o[0] = r[0] + T[4]; o[1] = r[1] + T[5]; o[2] = r[2] + T[6];
}
__device__ __forceinline__ void se3_mul(const double *A, const double *B, double *C)
{   // Sophus SO3 product incl. its first-order renormalisation
    double ax = A[0], ay = A[1], az = A[2], aw = A[3];
    double bx = B[0], by = B[1], bz = B[2], bw = B[3];
    double q0 = aw * bx + ax * bw + ay * bz - az * by;
    double q1 = aw * by + ay * bw + az * bx - ax * bz;
    double q2 = aw * bz + az * bw + ax * by - ay * bx;
    double q3 = aw * bw - ax * bx - ay * by - az * bz;
    double n2 = q0 * q0 + q1 * q1 + q2 * q2 + q3 * q3;
    if (n2 != 1.0) { double s = 2.0 / (1.0 + n2); q0 *= s; q1 *= s; q2 *= s; q3 *= s; }
    double t[3];
    quat_rot(A, B + 4, t);
    C[0] = q0; C[1] = q1; C[2] = q2; C[3] = q3;
    C[4] = t[0] + A[4]; C[5] = t[1] + A[5]; C[6] = t[2] + A[6];
}
__device__ __forceinline__ void se3_exp(const double *a, double *T)
{   // Sophus::SE3d::exp, tangent = (upsilon, omega)
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
    double A, B;
    if (th < 1e-10) { A = 0.5; B = 1.0 / 6.0; }
    else { A = (1.0 - cos(th)) / th2; B = (th - sin(th)) / (th2 * th); }
    const double *u = a;
    double c0 = w[1] * u[2] - w[2] * u[1], c1 = w[2] * u[0] - w[0] * u[2], c2 = w[0] * u[1] - w[1] * u[0];
    double d0 = w[1] * c2 - w[2] * c1, d1 = w[2] * c0 - w[0] * c2, d2 = w[0] * c1 - w[1] * c0;
    T[4] = u[0] + A * c0 + B * d0; T[5] = u[1] + A * c1 + B * d1; T[6] = u[2] + A * c2 + B * d2;
}
__device__ __forceinline__ void se3_inv(const double *T, double *O)
{   // Sophus::SE3d::inverse: conjugate quaternion, -R^T t
    double q[4] = {-T[0], -T[1], -T[2], T[3]}, r[3];
    quat_rot(q, T + 4, r);
    O[0] = q[0]; O[1] = q[1]; O[2] = q[2]; O[3] = q[3]; O[4] = -r[0]; O[5] = -r[1]; O[6] = -r[2];
}
__device__ __forceinline__ void se3_log(const double *T, double *a)
{   // Sophus::SE3d::log -> (upsilon, omega); same branches as the host mirror (host/slam.cpp SE3::log)
    const double n2 = T[0] * T[0] + T[1] * T[1] + T[2] * T[2], w = T[3];
    double f, th;
    if (n2 < 1e-10 * 1e-10) { f = 2.0 / w - (2.0 / 3.0) * n2 / (w * w * w); th = f * sqrt(n2); }
    else { const double n = sqrt(n2); const double at = (w < 0) ? atan2(-n, -w) : atan2(n, w); f = 2.0 * at / n; th = f * n; }
    const double om[3] = {f * T[0], f * T[1], f * T[2]};
    double c;
    if (fabs(th) < 1e-10) c = 1.0 / 12.0;
    else { const double h = 0.5 * th; c = (1.0 - th * cos(h) / (2.0 * sin(h))) / (th * th); }
    const double *t = T + 4;
    const double x0 = om[1] * t[2] - om[2] * t[1], x1 = om[2] * t[0] - om[0] * t[2], x2 = om[0] * t[1] - om[1] * t[0];
    const double y0 = om[1] * x2 - om[2] * x1, y1 = om[2] * x0 - om[0] * x2, y2 = om[0] * x1 - om[1] * x0;
    a[0] = t[0] - 0.5 * x0 + c * y0; a[1] = t[1] - 0.5 * x1 + c * y1; a[2] = t[2] - 0.5 * x2 + c * y2;
    a[3] = om[0]; a[4] = om[1]; a[5] = om[2];
}
__device__ __forceinline__ void se3_oplus(const double *T, const double *upd, double *O)
{   // VertexPose::oplusImpl (g2o_types.h:40-60): O = exp(upd) * T
    double E[7];
    se3_exp(upd, E);
    se3_mul(E, T, O);
}
__device__ __forceinline__ void quat_to_R(const double *q, double *R)
{
    double x = q[0], y = q[1], z = q[2], w = q[3];
    R[0] = 1 - 2 * (y * y + z * z); R[1] = 2 * (x * y - z * w); R[2] = 2 * (x * z + y * w);
    R[3] = 2 * (x * y + z * w); R[4] = 1 - 2 * (x * x + z * z); R[5] = 2 * (y * z - x * w);
    R[6] = 2 * (x * z - y * w); R[7] = 2 * (y * z + x * w); R[8] = 1 - 2 * (x * x + y * y);
}
__device__ __forceinline__ void huber(double e2, double delta, double &rho0, double &rho1)
{
    double d2 = delta * delta;
    if (e2 <= d2) { rho0 = e2; rho1 = 1.0; }
    else { double rs = rsqrt(e2), s = e2 * rs; rho0 = 2 * s * delta - d2; rho1 = delta * rs; }   // one rsqrt instead of sqrt + divide
}

// EdgeProjectionPoseOnly::computeError (g2o_types.h:117-130)
__device__ __forceinline__ void po_error(const double *T, const double *K, const double *pw, const double *uv, double *e, double *pc)
{
    se3_act(T, pw, pc);
    double px = K[0] * pc[0] + K[2] * pc[2], py = K[1] * pc[1] + K[3] * pc[2], iz = 1.0 / pc[2];   // one reciprocal, two products
    e[0] = uv[0] - px * iz;
    e[1] = uv[1] - py * iz;
}
// EdgeProjectionPoseOnly::linearizeOplus (g2o_types.h:132-163), J 2x6 row-major
__device__ __forceinline__ void po_jac(const double *K, const double *pc, double *J)
{
    double fx = K[0], fy = K[1], X = pc[0], Y = pc[1], Z = pc[2];
    double Zi = 1.0 / (Z + 1e-18), Zi2 = Zi * Zi;
    J[0] = -fx * Zi; J[1] = 0; J[2] = fx * X * Zi2; J[3] = fx * X * Y * Zi2; J[4] = -fx - fx * X * X * Zi2; J[5] = fx * Y * Zi;
    J[6] = 0; J[7] = -fy * Zi; J[8] = fy * Y * Zi2; J[9] = fy + fy * Y * Y * Zi2; J[10] = -fy * X * Y * Zi2; J[11] = -fy * X * Zi;
}

// EdgeProjection::computeError (g2o_types.h:200-216): e = uv - proj(K (ext (T p)))
__device__ __forceinline__ void ba_error(const double *T, const double *ext, const double *K, const double *p, const double *uv,
                                         double *e, double *a /* T p */, double *c /* ext T p */)
{
    se3_act(T, p, a);
    se3_act(ext, a, c);
    double px = K[0] * c[0] + K[2] * c[2], py = K[1] * c[1] + K[3] * c[2], iz = 1.0 / c[2];
    e[0] = uv[0] - px * iz;
    e[1] = uv[1] - py * iz;
}
// analytic Jacobians of EdgeProjection w.r.t. the left-multiplicative pose update (2x6) and the landmark (2x3)
__device__ __forceinline__ void ba_jac_analytic(const double *T, const double *ext, const double *K, const double *a, const double *c,
                                                double *Jp, double *Jl)
{
    double Re[9], R[9];
    quat_to_R(ext, Re);
    quat_to_R(T, R);
    double fx = K[0], fy = K[1], X = c[0], Y = c[1], Z = c[2], Zi = 1.0 / (Z + 1e-18), Zi2 = Zi * Zi;
    double D[6] = {-fx * Zi, 0, fx * X * Zi2, 0, -fy * Zi, fy * Y * Zi2};
    double ax[9] = {0, -a[2], a[1], a[2], 0, -a[0], -a[1], a[0], 0};
    double M[18];
#pragma unroll
    for (int i = 0; i < 3; i++)
#pragma unroll
        for (int j = 0; j < 3; j++) {
            M[i * 6 + j] = Re[i * 3 + j];
            double s = 0;
#pragma unroll
            for (int k = 0; k < 3; k++) s += Re[i * 3 + k] * ax[k * 3 + j];
            M[i * 6 + 3 + j] = -s;
        }
#pragma unroll
    for (int i = 0; i < 2; i++)
#pragma unroll
        for (int j = 0; j < 6; j++) {
            double s = 0;
#pragma unroll
            for (int k = 0; k < 3; k++) s += D[i * 3 + k] * M[k * 6 + j];
            Jp[i * 6 + j] = s;
        }
    double RR[9];
#pragma unroll
    for (int i = 0; i < 3; i++)
#pragma unroll
        for (int j = 0; j < 3; j++) {
            double s = 0;
#pragma unroll
            for (int k = 0; k < 3; k++) s += Re[i * 3 + k] * R[k * 3 + j];
            RR[i * 3 + j] = s;
        }
#pragma unroll
    for (int i = 0; i < 2; i++)
#pragma unroll
        for (int j = 0; j < 3; j++) {
            double s = 0;
#pragma unroll
            for (int k = 0; k < 3; k++) s += D[i * 3 + k] * RR[k * 3 + j];
            Jl[i * 3 + j] = s;
        }
}
// g2o's default numeric Jacobian for a binary edge: central differences, delta = 1e-9
__device__ __forceinline__ void ba_jac_numeric(const double *T, const double *ext, const double *K, const double *p, const double *uv,
                                               double *Jp, double *Jl)
{
    const double delta = 1e-9, scalar = 1.0 / (2 * delta);
    double a[3], c[3];
    for (int d = 0; d < 6; d++) {
        double add[6] = {0, 0, 0, 0, 0, 0}, Tp[7], e1[2], e2[2];
        add[d] = delta; se3_oplus(T, add, Tp); ba_error(Tp, ext, K, p, uv, e1, a, c);
        add[d] = -delta; se3_oplus(T, add, Tp); ba_error(Tp, ext, K, p, uv, e2, a, c);
        Jp[d] = scalar * (e1[0] - e2[0]); Jp[6 + d] = scalar * (e1[1] - e2[1]);
    }
    for (int d = 0; d < 3; d++) {
        double pp[3] = {p[0], p[1], p[2]}, e1[2], e2[2];
        pp[d] = p[d] + delta; ba_error(T, ext, K, pp, uv, e1, a, c);
        pp[d] = p[d]; pp[d] += -delta; ba_error(T, ext, K, pp, uv, e2, a, c);
        Jl[d] = scalar * (e1[0] - e2[0]); Jl[3 + d] = scalar * (e1[1] - e2[1]);
    }
}

__device__ __forceinline__ bool inv3(const double *A, double *I)
{
    double d = A[0] * (A[4] * A[8] - A[5] * A[7]) - A[1] * (A[3] * A[8] - A[5] * A[6]) + A[2] * (A[3] * A[7] - A[4] * A[6]);
    double id = 1.0 / d;
    I[0] = (A[4] * A[8] - A[5] * A[7]) * id; I[1] = (A[2] * A[7] - A[1] * A[8]) * id; I[2] = (A[1] * A[5] - A[2] * A[4]) * id;
    I[3] = (A[5] * A[6] - A[3] * A[8]) * id; I[4] = (A[0] * A[8] - A[2] * A[6]) * id; I[5] = (A[2] * A[3] - A[0] * A[5]) * id;
    I[6] = (A[3] * A[7] - A[4] * A[6]) * id; I[7] = (A[1] * A[6] - A[0] * A[7]) * id; I[8] = (A[0] * A[4] - A[1] * A[3]) * id;
    return isfinite(id);
}

__device__ __forceinline__ double warp_sum(double v)
{   // xor butterfly: every lane ends with the bitwise-identical total
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// LM step-acceptance bookkeeping shared by both solvers (g2o OptimizationAlgorithmLevenberg::solve)
struct LmCtl { double lambda, ni; };
__device__ __forceinline__ bool lm_accept(LmCtl &s, double rho, double tmp_chi)
{
    if (rho > 0 && isfinite(tmp_chi)) {
        double alpha = 1.0 - pow(2 * rho - 1, 3.0);
        alpha = fmin(alpha, 2.0 / 3.0);
        double sf = fmax(1.0 / 3.0, alpha);
        s.lambda *= sf; s.ni = 2;
        return true;
    }
    s.lambda *= s.ni; s.ni *= 2;
    return false;
}

}  // namespace gd
