// Pose-graph optimisation: the g2o block of LoopClosure::PoseGraphOptimization, reference
// src/loopclosure.cpp:641-746 — one VertexPose per keyframe (left-multiplicative update exp(d) * T, g2o_types.h:40-60), the
// vertex of keyframe 0 fixed (:693-696), one EdgePoseGraph (g2o_types.h:231-267) per consecutive-keyframe pair and per closed
// loop: error = log(M^-1 * v0 * v1^-1), information I6, no robust kernel, g2o's numeric central-difference Jacobians
// (the edge has no linearizeOplus), BlockSolver<6,6> + LinearSolverDense = dense pivoted LDL^T of the whole 6K x 6K system,
// Levenberg-Marquardt, optimize(22) — and the landmark move that follows it (:749-777).
//
// k_pg_lm: one persistent cooperative kernel runs the whole optimize(): per LM iteration the edges are linearised in
// parallel (one thread per edge: error + two 6x6 Jacobians), the block-sparse H is assembled in a fixed order (vertex ->
// incident edges, vertex pair -> its edges: no atomics, bitwise reproducible), and every trial permutes H + lambda I into
// Eigen's pivot order and solves it with the grid-cooperative blocked LDL^T of coop_ldlt.cuh; accept / reject, lambda schedule
// and stopping tests as in k_ba_window.  jacobian_mode 1 = g2o's numeric differences (delta 1e-9, the reference), 0 = the
// closed form (Ad(M^-1) and the inverse left Jacobian of SE3) for a noise-free comparison with the oracle.
#include "svs_internal.h"
#include "geom_dev.cuh"
#include "coop_ldlt.cuh"
#include <algorithm>
#include <cstring>
#include <vector>

namespace cg = cooperative_groups;

#define PG_T 512
enum { PC_LAMBDA = 0, PC_NI, PC_CUR, PC_RHO, PC_ACCEPT, PC_OK, PC_CHI_INIT, PC_ITERS, PC_TRIALS, PC_LINS, PC_SIGN, PC_NEW, PC_COUNT = 16 };

struct PgDev {
    int N, NA, E, n, n_pairs;
    double *poses, *poseT;                 // 7 N
    const int32_t *idx;                    // N: row block of a vertex, -1 fixed / unused
    const int32_t *act;                    // NA: vertex of row block
    const int32_t *edge_a, *edge_b;        // E
    const double *meas;                    // 7 E
    const int32_t *v_off, *v_edge;         // CSR vertex (active index) -> incident edges (edge*2 + role), creation order
    const int32_t *p_off, *p_edge, *p_u, *p_v;   // vertex pairs (active indices u < v) -> their edges (edge*2 + (a is u ? 0 : 1))
    double *J, *err;                       // 72 E (Ja | Jb), 6 E
    double *H, *b, *A, *dvec, *xs, *x, *cta, *ctl;
    int *perm;
    int jac_mode;
};

__device__ __forceinline__ void pg_error(const double *M, const double *A, const double *B, double *e)
{
    double Mi[7], Bi[7], t1[7], t2[7];
    gd::se3_inv(M, Mi); gd::se3_inv(B, Bi);
    gd::se3_mul(Mi, A, t1); gd::se3_mul(t1, Bi, t2);
    gd::se3_log(t2, e);
}
__device__ __forceinline__ void m3mul(const double *A, const double *B, double *C)
{
#pragma unroll
    for (int i = 0; i < 3; i++)
#pragma unroll
        for (int j = 0; j < 3; j++) C[i * 3 + j] = A[i * 3] * B[j] + A[i * 3 + 1] * B[3 + j] + A[i * 3 + 2] * B[6 + j];
}
__device__ __forceinline__ void hat3(const double *v, double *M) { M[0] = 0; M[1] = -v[2]; M[2] = v[1]; M[3] = v[2]; M[4] = 0; M[5] = -v[0]; M[6] = -v[1]; M[7] = v[0]; M[8] = 0; }
// inverse of the SE3 left Jacobian at xi = (rho, phi) (Barfoot 7.85-7.95), 6x6 row-major
__device__ void se3_left_jac_inv(const double *xi, double *Ji)
{
    const double *rho = xi, *phi = xi + 3;
    const double th2 = phi[0] * phi[0] + phi[1] * phi[1] + phi[2] * phi[2], th = sqrt(th2);
    double P[9], R[9], PP[9], PR[9], RP[9], PRP[9], PPR[9], RPP[9], PRPP[9], PPRP[9];
    hat3(phi, P); hat3(rho, R);
    m3mul(P, P, PP); m3mul(P, R, PR); m3mul(R, P, RP); m3mul(PR, P, PRP); m3mul(PP, R, PPR); m3mul(RP, P, RPP);
    m3mul(PRP, P, PRPP); m3mul(PP, RP, PPRP);
    double c1, c2, c3, a;
    if (th < 1e-5) {
        c1 = 1.0 / 6.0 - th2 / 120.0; c2 = 1.0 / 24.0 - th2 / 720.0; c3 = 1.0 / 120.0 - th2 / 2520.0;
        a = 1.0 / 12.0 + th2 / 720.0;
    } else {
        const double s = sin(th), c = cos(th), th3 = th2 * th, th4 = th2 * th2, th5 = th4 * th;
        c1 = (th - s) / th3; c2 = (1.0 - 0.5 * th2 - c) / th4; c3 = 0.5 * ((1.0 - 0.5 * th2 - c) / th4 - 3.0 * (th - s - th3 / 6.0) / th5);
        a = (1.0 - 0.5 * th * s / (1.0 - c)) / th2;
    }
    double Jinv[9], Q[9], JQ[9], JQJ[9];
#pragma unroll
    for (int i = 0; i < 9; i++) {
        const double I = (i % 4 == 0) ? 1.0 : 0.0;
        Jinv[i] = I - 0.5 * P[i] + a * PP[i];
        Q[i] = 0.5 * R[i] + c1 * (PR[i] + RP[i] + PRP[i]) - c2 * (PPR[i] + RPP[i] - 3.0 * PRP[i]) - c3 * (PRPP[i] + PPRP[i]);
    }
    m3mul(Jinv, Q, JQ); m3mul(JQ, Jinv, JQJ);
#pragma unroll
    for (int i = 0; i < 3; i++)
#pragma unroll
        for (int j = 0; j < 3; j++) {
            Ji[i * 6 + j] = Jinv[i * 3 + j]; Ji[i * 6 + 3 + j] = -JQJ[i * 3 + j];
            Ji[(3 + i) * 6 + j] = 0.0; Ji[(3 + i) * 6 + 3 + j] = Jinv[i * 3 + j];
        }
}
__device__ void pg_jac(const double *M, const double *A, const double *B, const double *e0, int mode, double *Ja, double *Jb)
{
    if (mode == 1) {   // g2o BaseBinaryEdge::linearizeOplus: central differences, delta = 1e-9
        const double delta = 1e-9, scalar = 1.0 / (2 * delta);
        for (int v = 0; v < 2; v++)
            for (int d = 0; d < 6; d++) {
                double add[6] = {0, 0, 0, 0, 0, 0}, Tp[7], e1[6], e2[6];
                const double *T = v ? B : A;
                add[d] = delta; gd::se3_oplus(T, add, Tp); pg_error(M, v ? A : Tp, v ? Tp : B, e1);
                add[d] = -delta; gd::se3_oplus(T, add, Tp); pg_error(M, v ? A : Tp, v ? Tp : B, e2);
                for (int r = 0; r < 6; r++) (v ? Jb : Ja)[r * 6 + d] = scalar * (e1[r] - e2[r]);
            }
        return;
    }
    double me0[6], Jl[36], Jr[36], Mi[7], Rm[9], tx[9], tR[9];
    for (int i = 0; i < 6; i++) me0[i] = -e0[i];
    se3_left_jac_inv(e0, Jl); se3_left_jac_inv(me0, Jr);
    gd::se3_inv(M, Mi); gd::quat_to_R(Mi, Rm); hat3(Mi + 4, tx); m3mul(tx, Rm, tR);
    for (int i = 0; i < 6; i++)
        for (int j = 0; j < 6; j++) {
            // Ad(M^-1) = [[R, t^ R], [0, R]]
            double s = 0;
            for (int k = 0; k < 6; k++) {
                double ad;
                if (k < 3) ad = (j < 3) ? Rm[k * 3 + j] : tR[k * 3 + j - 3];
                else ad = (j < 3) ? 0.0 : Rm[(k - 3) * 3 + j - 3];
                s += Jl[i * 6 + k] * ad;
            }
            Ja[i * 6 + j] = s; Jb[i * 6 + j] = -Jr[i * 6 + j];
        }
}

__device__ __forceinline__ double pg_block_sum(double v, double *red)
{
    const int tid = threadIdx.x;
    red[tid] = v;
    __syncthreads();
    for (int s = PG_T / 2; s > 0; s >>= 1) { if (tid < s) red[tid] += red[tid + s]; __syncthreads(); }
    const double r = red[0];
    __syncthreads();
    return r;
}
__device__ __forceinline__ double pg_column(const double *cta, int col, int n_cta, double *red)
{
    double a = 0;
    for (int i = threadIdx.x; i < n_cta; i += PG_T) a += cta[4 * i + col];
    return pg_block_sum(a, red);
}

__global__ void __launch_bounds__(PG_T, 1) k_pg_lm(PgDev D, int max_iter)
{
    cg::grid_group grid = cg::this_grid();
    extern __shared__ double smd[];
    double *red = smd, *sm_ldlt = smd + PG_T;
    const int tid = threadIdx.x, gtid = blockIdx.x * PG_T + tid, gsz = gridDim.x * PG_T;
    const int n = D.n, pitch = n + 1, n_cta = gridDim.x;
    double *ctl = D.ctl;
    unsigned ldlt_gen = 0;
    if (gtid == 0) D.perm[D.n] = 0;          // sub-grid barrier counter of the LDLT
    for (int i = gtid; i < 7 * D.N; i += gsz) D.poseT[i] = D.poses[i];
    for (long long i = gtid; i < (long long)n * n; i += gsz) D.H[i] = 0.0;
    if (gtid == 0) for (int i = 0; i < PC_COUNT; i++) ctl[i] = 0.0;
    if (gtid == 0) ctl[PC_NI] = 2;
    grid.sync();
    for (int it = 0; it < max_iter; it++) {
        // ---- linearise: error + Jacobians per edge, chi2
        {
            double acc = 0;
            for (int e = gtid; e < D.E; e += gsz) {
                const double *M = D.meas + 7 * (size_t)e, *A = D.poses + 7 * D.edge_a[e], *B = D.poses + 7 * D.edge_b[e];
                double e0[6];
                pg_error(M, A, B, e0);
                pg_jac(M, A, B, e0, D.jac_mode, D.J + 72 * (size_t)e, D.J + 72 * (size_t)e + 36);
                for (int r = 0; r < 6; r++) { D.err[6 * (size_t)e + r] = e0[r]; acc += e0[r] * e0[r]; }
            }
            const double t = pg_block_sum(acc, red);
            if (tid == 0) D.cta[4 * blockIdx.x] = t;
        }
        grid.sync();
        // ---- H (block-sparse inside the dense n x n array), b — fixed summation orders
        for (int t = gtid; t < D.NA * 42; t += gsz) {       // diagonal blocks and the gradient
            const int v = t / 42, ent = t - 42 * v;
            double s = 0;
            for (int q = D.v_off[v]; q < D.v_off[v + 1]; q++) {
                const int e = D.v_edge[q] >> 1, role = D.v_edge[q] & 1;
                const double *Jv = D.J + 72 * (size_t)e + 36 * role;
                if (ent < 36) { const int r = ent / 6, c = ent - 6 * r; for (int k = 0; k < 6; k++) s += Jv[k * 6 + r] * Jv[k * 6 + c]; }
                else { const int r = ent - 36; for (int k = 0; k < 6; k++) s -= Jv[k * 6 + r] * D.err[6 * (size_t)e + k]; }
            }
            if (ent < 36) D.H[(size_t)(6 * v + ent / 6) * n + 6 * v + ent % 6] = s;
            else D.b[6 * v + ent - 36] = s;
        }
        for (int t = gtid; t < D.n_pairs * 36; t += gsz) {  // off-diagonal blocks H_uv = sum J_u^T J_v and their transposes
            const int p = t / 36, ent = t - 36 * p, r = ent / 6, c = ent - 6 * r;
            double s = 0;
            for (int q = D.p_off[p]; q < D.p_off[p + 1]; q++) {
                const int e = D.p_edge[q] >> 1, swap = D.p_edge[q] & 1;      // swap: vertex a of the edge is v, b is u
                const double *Ju = D.J + 72 * (size_t)e + (swap ? 36 : 0), *Jv = D.J + 72 * (size_t)e + (swap ? 0 : 36);
                for (int k = 0; k < 6; k++) s += Ju[k * 6 + r] * Jv[k * 6 + c];
            }
            const int u = D.p_u[p], v = D.p_v[p];
            D.H[(size_t)(6 * u + r) * n + 6 * v + c] = s;
            D.H[(size_t)(6 * v + c) * n + 6 * u + r] = s;
        }
        grid.sync();
        if (blockIdx.x == 0) {
            const double cur = pg_column(D.cta, 0, n_cta, red);
            if (it == 0) {
                double md = 0;
                for (int i = tid; i < n; i += PG_T) md = fmax(md, fabs(D.H[(size_t)i * n + i]));
                red[tid] = md;
                __syncthreads();
                for (int s = PG_T / 2; s > 0; s >>= 1) { if (tid < s) red[tid] = fmax(red[tid], red[tid + s]); __syncthreads(); }
                if (tid == 0) { ctl[PC_LAMBDA] = 1e-5 * red[0]; ctl[PC_NI] = 2; ctl[PC_CHI_INIT] = cur; }
                __syncthreads();
            }
            if (tid == 0) { ctl[PC_CUR] = cur; ctl[PC_LINS] += 1; }
        }
        grid.sync();
        int q = 0;
        double rho = 0;
        do {
            const double lambda = ctl[PC_LAMBDA], chi_cur = ctl[PC_CUR];
            // pivot order: rank of |H_ii + lambda| in decreasing order (ties: lower index first)
            for (int i = gtid; i < n; i += gsz) {
                const double di = fabs(D.H[(size_t)i * n + i] + lambda);
                int r = 0;
                for (int j = 0; j < n; j++) { const double dj = fabs(D.H[(size_t)j * n + j] + lambda); r += (dj > di || (dj == di && j < i)) ? 1 : 0; }
                D.perm[r] = i;
            }
            if (gtid == 0) ctl[PC_SIGN] = 0;
            grid.sync();
            for (long long t = gtid; t < (long long)(n + 1) * n; t += gsz) {
                const int a = (int)(t / n), c = (int)(t - (long long)a * n);
                if (a == n) { D.A[(size_t)a * pitch + c] = D.b[D.perm[c]]; continue; }
                if (c > a) continue;
                double v = D.H[(size_t)D.perm[a] * n + D.perm[c]];
                if (a == c) v += lambda;
                D.A[(size_t)a * pitch + c] = v;
            }
            grid.sync();
            coop_ldlt_solve<PG_T>(grid, D.A, n, pitch, D.dvec, D.xs, ctl + PC_SIGN, sm_ldlt, reinterpret_cast<unsigned *>(D.perm + D.n), ldlt_gen);
            if (blockIdx.x == 0) {
                const int sign = (int)ctl[PC_SIGN];
                const bool ok = (sign == 1 || sign == 0);
                for (int i = tid; i < n; i += PG_T) D.x[D.perm[i]] = ok ? D.xs[i] : 0.0;
                __syncthreads();
                for (int v = tid; v < D.N; v += PG_T) {
                    const int a = D.idx[v];
                    if (a >= 0) gd::se3_oplus(D.poses + 7 * v, D.x + 6 * a, D.poseT + 7 * v);
                }
                double sc = 0;
                for (int i = tid; i < n; i += PG_T) sc += D.x[i] * (lambda * D.x[i] + D.b[i]);
                const double scale = pg_block_sum(sc, red);
                if (tid == 0) { ctl[PC_OK] = ok ? 1.0 : 0.0; ctl[PC_NEW] = scale; }
            }
            grid.sync();
            {
                double acc = 0;
                for (int e = gtid; e < D.E; e += gsz) {
                    double e1[6];
                    pg_error(D.meas + 7 * (size_t)e, D.poseT + 7 * D.edge_a[e], D.poseT + 7 * D.edge_b[e], e1);
                    for (int r = 0; r < 6; r++) acc += e1[r] * e1[r];
                }
                const double t = pg_block_sum(acc, red);
                if (tid == 0) D.cta[4 * blockIdx.x + 1] = t;
            }
            grid.sync();
            if (blockIdx.x == 0) {
                const double chn = pg_column(D.cta, 1, n_cta, red);
                if (tid == 0) {
                    const bool ok = ctl[PC_OK] != 0.0;
                    const double tc = ok ? chn : DBL_MAX;
                    const double r = (chi_cur - tc) / (ctl[PC_NEW] + 1e-3);
                    gd::LmCtl lm = {ctl[PC_LAMBDA], ctl[PC_NI]};
                    const int acc2 = gd::lm_accept(lm, r, tc) ? 1 : 0;
                    ctl[PC_LAMBDA] = lm.lambda; ctl[PC_NI] = lm.ni; ctl[PC_RHO] = r; ctl[PC_ACCEPT] = acc2;
                    if (acc2) ctl[PC_CUR] = tc;
                    ctl[PC_TRIALS] += 1;
                }
            }
            grid.sync();
            rho = ctl[PC_RHO];
            if (ctl[PC_ACCEPT] != 0.0)
                for (int i = gtid; i < 7 * D.N; i += gsz) D.poses[i] = D.poseT[i];
            grid.sync();
            q++;
        } while (rho < 0 && q < 10);
        if (gtid == 0) ctl[PC_ITERS] += 1;
        if (q == 10 || rho == 0) break;
    }
}

__global__ void k_pg_move_landmarks(int n_lm, double *lms, const int32_t *lm_kf, const double *old_poses, const double *new_poses)
{   // src/loopclosure.cpp:749-777: pos_w = new_pose^-1 * (old_pose * pos)
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_lm) return;
    const int k = lm_kf[i];
    if (k < 0) return;
    double s[3], inv[7], o[3];
    gd::se3_act(old_poses + 7 * (size_t)k, lms + 3 * (size_t)i, s);
    gd::se3_inv(new_poses + 7 * (size_t)k, inv);
    gd::se3_act(inv, s, o);
    lms[3 * (size_t)i] = o[0]; lms[3 * (size_t)i + 1] = o[1]; lms[3 * (size_t)i + 2] = o[2];
}

extern "C" {

int svs_pose_graph_optimize(svs_ctx *c, int n_kf, double *poses, const uint8_t *fixed, int n_edge, const int32_t *edge_a,
                            const int32_t *edge_b, const double *meas, int max_iter, int jacobian_mode, svs_ba_stats *stats)
{
    if (!c || n_kf <= 0 || !poses || !fixed || n_edge < 0 || max_iter < 0 || (n_edge > 0 && (!edge_a || !edge_b || !meas))) return SVS_ERR_ARG;
    if (stats) memset(stats, 0, sizeof(*stats));
    if (n_edge == 0) return SVS_OK;
    SVS_CUDA(c, cudaSetDevice(c->device));
    const int N = n_kf, E = n_edge;
    for (int e = 0; e < E; e++) if (edge_a[e] < 0 || edge_a[e] >= N || edge_b[e] < 0 || edge_b[e] >= N || edge_a[e] == edge_b[e]) SVS_FAIL(c, SVS_ERR_ARG, "pose graph: bad edge");
    // active = non-fixed vertices with >= 1 edge, ascending id (g2o's Hessian index order)
    std::vector<int32_t> idx(N, -1), act;
    for (int e = 0; e < E; e++) { if (!fixed[edge_a[e]]) idx[edge_a[e]] = 0; if (!fixed[edge_b[e]]) idx[edge_b[e]] = 0; }
    for (int i = 0; i < N; i++) if (idx[i] == 0) { idx[i] = (int32_t)act.size(); act.push_back(i); }
    const int NA = (int)act.size(), n = 6 * NA;
    if (n == 0) return SVS_OK;
    std::vector<int32_t> v_off(NA + 1, 0), v_edge, p_off, p_edge, p_u, p_v;
    for (int e = 0; e < E; e++) { if (idx[edge_a[e]] >= 0) v_off[idx[edge_a[e]] + 1]++; if (idx[edge_b[e]] >= 0) v_off[idx[edge_b[e]] + 1]++; }
    for (int v = 0; v < NA; v++) v_off[v + 1] += v_off[v];
    v_edge.resize(v_off[NA]);
    { std::vector<int> fill(v_off.begin(), v_off.end() - 1);
      for (int e = 0; e < E; e++) {
          if (idx[edge_a[e]] >= 0) v_edge[fill[idx[edge_a[e]]]++] = 2 * e;
          if (idx[edge_b[e]] >= 0) v_edge[fill[idx[edge_b[e]]]++] = 2 * e + 1;
      } }
    {   // edges grouped by unordered active vertex pair, creation order inside a pair
        std::vector<std::pair<long long, int>> keyed;
        for (int e = 0; e < E; e++) {
            const int ia = idx[edge_a[e]], ib = idx[edge_b[e]];
            if (ia < 0 || ib < 0) continue;
            const int u = std::min(ia, ib), v = std::max(ia, ib);
            keyed.push_back({(long long)u * NA + v, 2 * e + (ia == u ? 0 : 1)});
        }
        std::stable_sort(keyed.begin(), keyed.end(), [](const std::pair<long long, int> &x, const std::pair<long long, int> &y) { return x.first < y.first; });
        for (size_t i = 0; i < keyed.size(); i++) {
            if (i == 0 || keyed[i].first != keyed[i - 1].first) { p_off.push_back((int32_t)i); p_u.push_back((int32_t)(keyed[i].first / NA)); p_v.push_back((int32_t)(keyed[i].first % NA)); }
            p_edge.push_back(keyed[i].second);
        }
        p_off.push_back((int32_t)keyed.size());
    }
    const int n_pairs = (int)p_u.size();
    // ---- device layout
    struct Seg { const void *src; size_t bytes, off; };
    std::vector<Seg> segs;
    size_t tot = 0;
    auto add = [&](const void *p, size_t bytes) { size_t o = tot; segs.push_back({p, bytes, o}); tot = align_up(tot + bytes, 256); return o; };
    const size_t o_pose = add(poses, (size_t)N * 56), o_poseT = add(poses, (size_t)N * 56), o_idx = add(idx.data(), (size_t)N * 4), o_act = add(act.data(), (size_t)NA * 4);
    const size_t o_ea = add(edge_a, (size_t)E * 4), o_eb = add(edge_b, (size_t)E * 4), o_m = add(meas, (size_t)E * 56);
    const size_t o_vo = add(v_off.data(), v_off.size() * 4), o_ve = add(v_edge.data(), v_edge.size() * 4 + 4);
    const size_t o_po = add(p_off.data(), p_off.size() * 4), o_pe = add(p_edge.data(), p_edge.size() * 4 + 4), o_pu = add(p_u.data(), p_u.size() * 4 + 4), o_pv = add(p_v.data(), p_v.size() * 4 + 4);
    auto res = [&](size_t bytes) { size_t o = tot; tot = align_up(tot + bytes, 256); return o; };
    const size_t o_J = res((size_t)E * 576), o_err = res((size_t)E * 48), o_H = res((size_t)n * n * 8), o_b = res((size_t)n * 8);
    const size_t o_A = res((size_t)(n + 1) * (n + 1) * 8), o_dv = res((size_t)n * 8), o_xs = res((size_t)n * 8), o_x = res((size_t)n * 8);
    const size_t o_cta = res(2048 * 32), o_ctl = res(PC_COUNT * 8), o_perm = res((size_t)n * 4 + 16);
    SVS_CUDA(c, c->d_tmp.reserve(tot));
    uint8_t *db = c->d_tmp.as<uint8_t>();
    SVS_CUDA(c, c->h_in.reserve(tot > 0 ? o_J : 0));
    uint8_t *hb = c->h_in.as<uint8_t>();
    for (const Seg &s : segs) if (s.bytes && s.src) memcpy(hb + s.off, s.src, s.bytes);
    SVS_CUDA(c, cudaMemcpyAsync(db, hb, o_J, cudaMemcpyHostToDevice, c->stream));
    PgDev d;
    d.N = N; d.NA = NA; d.E = E; d.n = n; d.n_pairs = n_pairs;
    d.poses = (double *)(db + o_pose); d.poseT = (double *)(db + o_poseT); d.idx = (int32_t *)(db + o_idx); d.act = (int32_t *)(db + o_act);
    d.edge_a = (int32_t *)(db + o_ea); d.edge_b = (int32_t *)(db + o_eb); d.meas = (double *)(db + o_m);
    d.v_off = (int32_t *)(db + o_vo); d.v_edge = (int32_t *)(db + o_ve); d.p_off = (int32_t *)(db + o_po); d.p_edge = (int32_t *)(db + o_pe);
    d.p_u = (int32_t *)(db + o_pu); d.p_v = (int32_t *)(db + o_pv);
    d.J = (double *)(db + o_J); d.err = (double *)(db + o_err); d.H = (double *)(db + o_H); d.b = (double *)(db + o_b); d.A = (double *)(db + o_A);
    d.dvec = (double *)(db + o_dv); d.xs = (double *)(db + o_xs); d.x = (double *)(db + o_x); d.cta = (double *)(db + o_cta); d.ctl = (double *)(db + o_ctl);
    d.perm = (int *)(db + o_perm); d.jac_mode = jacobian_mode;
    const size_t smem = ((size_t)PG_T + CL_SMEM_DOUBLES) * 8;
    SVS_CUDA(c, svs_i_opt_in_smem(c, reinterpret_cast<const void *>(k_pg_lm)));
    int per_sm = 0;
    SVS_CUDA(c, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_pg_lm, PG_T, smem));
    if (per_sm < 1) SVS_FAIL(c, SVS_ERR_CAPACITY, "pose graph: the cooperative solver does not fit an SM");
    // small graphs: fewer CTAs (a grid barrier costs more the more CTAs take part)
    int grid = std::max(1, std::min(c->sm_count, std::max(4, (n * n / 2) / (64 * 64) + E / PG_T)));
    void *args[] = {&d, &max_iter};
    svs_i_prof_begin(c, KID_BA_WINDOW);
    cudaError_t e = cudaLaunchCooperativeKernel(reinterpret_cast<const void *>(k_pg_lm), dim3(grid), dim3(PG_T), args, smem, c->stream);
    svs_i_prof_end(c);
    SVS_CUDA(c, e);
    c->launches++;
    double ctl[PC_COUNT];
    SVS_CUDA(c, cudaMemcpyAsync(ctl, d.ctl, sizeof(ctl), cudaMemcpyDeviceToHost, c->stream));
    SVS_CUDA(c, cudaMemcpyAsync(poses, d.poses, (size_t)N * 56, cudaMemcpyDeviceToHost, c->stream));
    SVS_CUDA(c, cudaStreamSynchronize(c->stream));
    if (stats) {
        stats->iterations = (int)ctl[PC_ITERS]; stats->trials = (int)ctl[PC_TRIALS]; stats->linearizations = (int)ctl[PC_LINS];
        stats->solves = (int)ctl[PC_TRIALS]; stats->lambda = ctl[PC_LAMBDA]; stats->chi2 = ctl[PC_CUR]; stats->chi2_init = ctl[PC_CHI_INIT];
    }
    return SVS_OK;
}

int svs_pose_graph_move_landmarks(svs_ctx *c, int n_lm, double *lms, const int32_t *lm_kf, int n_kf, const double *old_poses, const double *new_poses)
{
    if (!c || n_lm < 0 || n_kf <= 0 || !old_poses || !new_poses || (n_lm > 0 && (!lms || !lm_kf))) return SVS_ERR_ARG;
    if (n_lm == 0) return SVS_OK;
    for (int i = 0; i < n_lm; i++) if (lm_kf[i] >= n_kf) SVS_FAIL(c, SVS_ERR_ARG, "pose graph: landmark keyframe index out of range");
    SVS_CUDA(c, cudaSetDevice(c->device));
    const size_t lb = align_up((size_t)n_lm * 24, 256), kb = align_up((size_t)n_lm * 4, 256), pb = align_up((size_t)n_kf * 56, 256);
    SVS_CUDA(c, c->d_tmp2.reserve(lb + kb + 2 * pb));
    uint8_t *db = c->d_tmp2.as<uint8_t>();
    SVS_CUDA(c, cudaMemcpyAsync(db, lms, (size_t)n_lm * 24, cudaMemcpyHostToDevice, c->stream));
    SVS_CUDA(c, cudaMemcpyAsync(db + lb, lm_kf, (size_t)n_lm * 4, cudaMemcpyHostToDevice, c->stream));
    SVS_CUDA(c, cudaMemcpyAsync(db + lb + kb, old_poses, (size_t)n_kf * 56, cudaMemcpyHostToDevice, c->stream));
    SVS_CUDA(c, cudaMemcpyAsync(db + lb + kb + pb, new_poses, (size_t)n_kf * 56, cudaMemcpyHostToDevice, c->stream));
    SVS_KERNEL(c, KID_MISC, k_pg_move_landmarks<<<(n_lm + 255) / 256, 256, 0, c->stream>>>(n_lm, (double *)db, (const int32_t *)(db + lb), (const double *)(db + lb + kb),
                                                                                           (const double *)(db + lb + kb + pb)));
    SVS_CUDA(c, cudaMemcpyAsync(lms, db, (size_t)n_lm * 24, cudaMemcpyDeviceToHost, c->stream));
    SVS_CUDA(c, cudaStreamSynchronize(c->stream));
    return SVS_OK;
}

}  // extern "C"
