// Sliding-window bundle adjustment (row a7): the g2o block of Backend::Optimize, reference
// src/backend.cpp:22-164 — VertexPose per keyframe, marginalised VertexXYZ per landmark, EdgeProjection
// (include/StereoVisionSLAM/g2o_types.h:176-229) per observation, Huber(delta), Levenberg-Marquardt with Schur
// complement and a dense pivoted LDLT of the reduced 6N x 6N camera system, no vertex fixed.
// Solver semantics: SURVEY.md Appendix B (upstream g2o; un-vendored, parity unpinned).
//
// k_ba_window: ONE persistent CTA per window problem runs the whole LM loop on-device (no host round trips):
//   linearise   landmark-parallel: residuals, 2x6 / 2x3 Jacobian blocks, Huber weights -> Hll (3x3), bl, and the 6x3
//               pose-landmark block W summed per GROUP = (landmark, keyframe): the left and right observation of a
//               landmark in one keyframe share the pose, so everything downstream (Schur, g, back-substitution) works
//               on ~0.6 E groups and ~0.4x the (edge, edge) pairs
//               pose-parallel (warp per keyframe, butterfly reduction) -> Hpp (6x6), bp
//   schur       landmark-parallel V^-1 = (Hll + lambda I)^-1, W V^-1 per group; the (group, group) pairs of every 6x6 block
//               (i,j) of the reduced system are listed by the host, sorted by block and cut into chunks of <= BA_CH
//               pairs; 4 threads own a chunk (one 3x3 quadrant each) and accumulate (W V^-1)_e1 W_e2^T in registers,
//               then 36 threads per block add the chunk partials in a fixed order -> no atomics, bitwise
//               deterministic, and the only traffic is L1/L2 reads of the 144-byte edge blocks
//   solve       pivoted LDLT (Eigen::LDLT order of operations) in shared memory, 4 lanes per row
//   back-subst  landmark-parallel; trial chi2; g2o's rho test / lambda schedule on thread 0
// The reduced system lives in shared memory (<= ~200 KB, i.e. N <= 26 keyframes); larger windows fall back
// to an L2-resident global buffer through the same code path.
#include "svs_internal.h"
#include "geom_dev.cuh"
#include <omp.h>
#include <algorithm>
#include <chrono>
#include <cstdlib>
#include <cstring>
#include <vector>

struct BaProb {
    int NA, L, E, nblk;
    int pose0;      // first pose (global index into poses[7*])
    int act0;       // act_pose[act0 + a] = local keyframe index of active pose a
    int lm0, e0;    // first landmark / edge (global)
    int loff0;      // l_off[loff0 + l .. ] (L+1 entries), values relative to e0
    int poff0;      // p_off[poff0 + a ..] (NA+1 entries), values relative to e0
    int blk0;       // blk_i/blk_j[blk0 + b]
    int G, grp0;    // groups = (landmark, active pose) with >= 1 edge, numbered landmark-major / pose-ascending; global offset
    int lgoff0;     // lg_off[lgoff0 + l ..] (L+1 entries): groups of landmark l
    int pgoff0;     // pg_off[pgoff0 + a ..] (NA+1 entries) into pg_groups[grp0 + ..]: groups of active pose a
    int pair0;      // pr_e1/pr_e2[pair0 + p]: the (group, group) pairs of this problem, sorted by block
    int nch, ch0;   // chunks: ch_blk[ch0 + c] = block of chunk c, pairs [ch_off[choff0 + c], ch_off[choff0 + c + 1])
    int choff0;
    int bch0;       // blk_ch[bch0 + b .. ] (nblk+1 entries): chunk range of block b
    long long part0;  // first partial-sum record (36 doubles per chunk) in the global scratch
    long long S_off;  // offset (doubles) into the global reduced-system scratch, or -1 when it fits shared memory
};

struct BaArgs {
    const BaProb *probs;
    double *poses, *lms;            // in/out
    const int32_t *act_pose;
    const int32_t *edge_p, *edge_l; // compact active-pose index, local landmark index
    const uint8_t *edge_cam;
    const double *edge_uv;
    const int32_t *l_off, *l_edges, *p_off, *p_edges;
    const int32_t *blk_i, *blk_j, *blk_ch, *pr_e1, *pr_e2, *ch_blk, *ch_off;
    const int32_t *lg_off, *g_lm, *g_pose, *pg_off, *pg_groups;
    double *part;                   // 36 doubles per chunk
    double *Hpl, *WD;               // 18 / group: W and W V^-1
    double *Hll, *Dinv;             // 9 / landmark
    double *bl, *xl, *lmT;          // 3 / landmark
    double *S_glob;
    double *edge_chi2;
    svs_ba_stats *stats;
    double K[2][4], ext[2][7];
    double huber_delta;
    int max_iter, jac_mode;
    int smem_bytes;                 // dynamic shared memory of this launch
};

#define BA_T 512    // one CTA per SM: a window problem is latency-bound, so it gets as many threads as 128 registers each allow
#define BA_CH 16    // (group, group) pairs per Schur chunk

// dynamic shared memory of a window with NA active poses: reductions + poses + Hpp/bp/g/xp/tmp + tr, plus n_S copies of
// the reduced system (0: it lives in global memory, 1: in-place pivoting LDLT, 2: pre-permuted LDLT)
__host__ __device__ inline size_t ba_smem_need(int NA, int n_S)
{
    size_t np = 6 * (size_t)NA, pitch = np | 1;
    size_t d = BA_T + 14 * (size_t)NA + 36 * (size_t)NA + 4 * np;
    return d * 8 + ((np + 1) & ~(size_t)1) * 4 + (size_t)n_S * np * pitch * 8 + 16;
}
#include "ba_ldlt.cuh"

template <int BT>
__device__ __forceinline__ double block_sum(double v, double *red)
{   // deterministic: xor-butterfly inside each warp, then every thread adds the BT/32 warp totals in the same order;
    // result broadcast to all threads; two barriers (the second lets `red` be reused at once)
    const int tid = threadIdx.x;
    v = gd::warp_sum(v);
    if ((tid & 31) == 0) red[tid >> 5] = v;
    __syncthreads();
    double r = 0;
#pragma unroll
    for (int w = 0; w < BT / 32; w++) r += red[w];
    __syncthreads();
    return r;
}
template <int BT>
__device__ __forceinline__ double block_max(double v, double *red)
{
    const int tid = threadIdx.x;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
    if ((tid & 31) == 0) red[tid >> 5] = v;
    __syncthreads();
    double r = red[0];
#pragma unroll
    for (int w = 1; w < BT / 32; w++) r = fmax(r, red[w]);
    __syncthreads();
    return r;
}

// BT = 512: one CTA per SM (fastest single window).  BT = 256: two CTAs per SM — a window takes longer but a launch with more
// windows than SMs runs in ONE wave instead of two (the kernel is latency-bound: 20 % issue-active at 512 threads).
template <int BT>
__global__ void __launch_bounds__(BT, BT == 512 ? 1 : 2)
k_ba_window(BaArgs A)
{
    extern __shared__ double smd[];
    const BaProb P = A.probs[blockIdx.x];
    int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int NA = P.NA, L = P.L, E = P.E, np = 6 * NA, pitch = np | 1;
    // ---- shared memory carve-up
    double *red = smd;                       // BT
    double *poseA = red + BT;              // 7*NA accepted
    double *poseT = poseA + 7 * NA;          // 7*NA trial
    double *Hpp = poseT + 7 * NA;            // 36*NA
    double *bp = Hpp + 36 * NA;              // np
    double *g = bp + np;                     // np
    double *xp = g + np;                     // np
    double *tmp = xp + np;                   // np
    int *tr = reinterpret_cast<int *>(tmp + np);   // np ints
    double *Ssm = reinterpret_cast<double *>(tr + ((np + 1) & ~1));
    double *S = (P.S_off >= 0) ? A.S_glob + P.S_off : Ssm;
    // second buffer for the pre-permuted LDLT when this launch's shared memory has room for it
    double *S2 = (P.S_off < 0 && ba_smem_need(NA, 2) <= (size_t)A.smem_bytes) ? Ssm + np * pitch : nullptr;
    __shared__ int s_piv, s_flag, s_q;
    __shared__ double s_lambda, s_ni, s_cur, s_rho;

    const int32_t *l_off = A.l_off + P.loff0, *l_edges = A.l_edges + P.e0;
    const int32_t *p_off = A.p_off + P.poff0, *p_edges = A.p_edges + P.e0;
    const int32_t *edge_p = A.edge_p + P.e0, *edge_l = A.edge_l + P.e0;
    const uint8_t *edge_cam = A.edge_cam + P.e0;
    const double *edge_uv = A.edge_uv + 2 * (size_t)P.e0;
    double *lms = A.lms + 3 * (size_t)P.lm0, *lmT = A.lmT + 3 * (size_t)P.lm0;
    double *Hll = A.Hll + 9 * (size_t)P.lm0, *Dinv = A.Dinv + 9 * (size_t)P.lm0;
    double *bl = A.bl + 3 * (size_t)P.lm0, *xl = A.xl + 3 * (size_t)P.lm0;
    double *Hpl = A.Hpl + 18 * (size_t)P.grp0, *WD = A.WD + 18 * (size_t)P.grp0;
    const int32_t *lg_off = A.lg_off + P.lgoff0, *g_lm = A.g_lm + P.grp0, *g_pose = A.g_pose + P.grp0;
    const int32_t *pg_off = A.pg_off + P.pgoff0, *pg_groups = A.pg_groups + P.grp0;
    const double hd = A.huber_delta;

    for (int i = tid; i < 7 * NA; i += BT) {
        int a = i / 7;
        double v = A.poses[7 * (size_t)(P.pose0 + A.act_pose[P.act0 + a]) + (i - 7 * a)];
        poseA[i] = v; poseT[i] = v;
    }
    for (int i = tid; i < 3 * L; i += BT) lmT[i] = lms[i];
    __syncthreads();

    int st_it = 0, st_tr = 0, st_lin = 0, st_sol = 0;
    double chi_init = 0;
    if (tid == 0) { s_lambda = 0; s_ni = 2; }
    bool stop = (E == 0 || NA == 0);
    for (int it = 0; it < A.max_iter && !stop; it++) {
        // ================= linearise at the accepted state =================
        double acc = 0;
        for (int l = tid; l < L; l += BT) {
            int s0 = l_off[l], s1 = l_off[l + 1];
            if (s0 == s1) continue;
            double H[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0}, b3[3] = {0, 0, 0};
            const double *pl = lms + 3 * l;
            int gi = lg_off[l] - 1, cur_p = -1;      // the edges of a landmark are listed pose-ascending: one run per group
            for (int s = s0; s < s1; s++) {
                int e = l_edges[s], cam = edge_cam[e];
                const int ep = edge_p[e];
                const bool first = ep != cur_p;
                if (first) { gi++; cur_p = ep; }
                const double *T = poseA + 7 * ep;
                double er[2], a[3], c[3], Jp[12], Jl[6];
                gd::ba_error(T, A.ext[cam], A.K[cam], pl, edge_uv + 2 * e, er, a, c);
                if (A.jac_mode == 1) gd::ba_jac_numeric(T, A.ext[cam], A.K[cam], pl, edge_uv + 2 * e, Jp, Jl);
                else gd::ba_jac_analytic(T, A.ext[cam], A.K[cam], a, c, Jp, Jl);
                double e2 = er[0] * er[0] + er[1] * er[1], r0, r1;
                gd::huber(e2, hd, r0, r1);
                acc += r0;
#pragma unroll
                for (int x = 0; x < 3; x++) {
                    b3[x] -= r1 * (Jl[x] * er[0] + Jl[3 + x] * er[1]);
#pragma unroll
                    for (int y = 0; y < 3; y++) H[x * 3 + y] += r1 * (Jl[x] * Jl[y] + Jl[3 + x] * Jl[3 + y]);
                }
                double Wv[18];
#pragma unroll
                for (int x = 0; x < 6; x++)
#pragma unroll
                    for (int y = 0; y < 3; y++) Wv[x * 3 + y] = r1 * (Jp[x] * Jl[y] + Jp[6 + x] * Jl[3 + y]);
                double2 *W = reinterpret_cast<double2 *>(Hpl + 18 * (size_t)gi);
                if (first) {
#pragma unroll
                    for (int x = 0; x < 9; x++) W[x] = make_double2(Wv[2 * x], Wv[2 * x + 1]);
                } else {      // second observation of the landmark in the same keyframe (this thread wrote the first one)
#pragma unroll
                    for (int x = 0; x < 9; x++) { double2 t = W[x]; W[x] = make_double2(t.x + Wv[2 * x], t.y + Wv[2 * x + 1]); }
                }
            }
#pragma unroll
            for (int x = 0; x < 9; x++) Hll[9 * (size_t)l + x] = H[x];
#pragma unroll
            for (int x = 0; x < 3; x++) bl[3 * (size_t)l + x] = b3[x];
        }
        double cur = block_sum<BT>(acc, red);
        if (it == 0) chi_init = cur;
        for (int a = warp; a < NA; a += BT / 32) {
            double H[21], b6[6];
#pragma unroll
            for (int x = 0; x < 21; x++) H[x] = 0;
#pragma unroll
            for (int x = 0; x < 6; x++) b6[x] = 0;
            const double *T = poseA + 7 * a;
            for (int s = p_off[a] + lane; s < p_off[a + 1]; s += 32) {
                int e = p_edges[s], cam = edge_cam[e];
                const double *pl = lms + 3 * (size_t)edge_l[e];
                double er[2], aa[3], c[3], Jp[12], Jl[6];
                gd::ba_error(T, A.ext[cam], A.K[cam], pl, edge_uv + 2 * e, er, aa, c);
                if (A.jac_mode == 1) gd::ba_jac_numeric(T, A.ext[cam], A.K[cam], pl, edge_uv + 2 * e, Jp, Jl);
                else gd::ba_jac_analytic(T, A.ext[cam], A.K[cam], aa, c, Jp, Jl);
                double e2 = er[0] * er[0] + er[1] * er[1], r0, r1;
                gd::huber(e2, hd, r0, r1);
                int k = 0;
#pragma unroll
                for (int x = 0; x < 6; x++) {
                    b6[x] -= r1 * (Jp[x] * er[0] + Jp[6 + x] * er[1]);
#pragma unroll
                    for (int y = x; y < 6; y++) H[k++] += r1 * (Jp[x] * Jp[y] + Jp[6 + x] * Jp[6 + y]);
                }
            }
#pragma unroll
            for (int x = 0; x < 21; x++) H[x] = gd::warp_sum(H[x]);
#pragma unroll
            for (int x = 0; x < 6; x++) b6[x] = gd::warp_sum(b6[x]);
            if (lane == 0) {
                int k = 0;
#pragma unroll
                for (int x = 0; x < 6; x++) {
                    bp[6 * a + x] = b6[x];
#pragma unroll
                    for (int y = x; y < 6; y++) { Hpp[36 * a + x * 6 + y] = H[k]; Hpp[36 * a + y * 6 + x] = H[k]; k++; }
                }
            }
        }
        __syncthreads();
        st_lin++;
        if (it == 0) {   // lambda_init = tau * max diagonal over all active vertices
            double md = 0;
            for (int i = tid; i < np; i += BT) md = fmax(md, fabs(Hpp[36 * (i / 6) + 7 * (i % 6)]));
            for (int l = tid; l < L; l += BT)
                if (l_off[l] != l_off[l + 1]) md = fmax(md, fmax(fabs(Hll[9 * (size_t)l]), fmax(fabs(Hll[9 * (size_t)l + 4]), fabs(Hll[9 * (size_t)l + 8]))));
            md = block_max<BT>(md, red);
            if (tid == 0) { s_lambda = 1e-5 * md; s_ni = 2; }
            __syncthreads();
        }
        // ================= trial loop =================
        int q = 0;
        double rho = 0;
        do {
            double lambda = s_lambda;
            // ---- reduced system: S = Hpp + lambda I (block diagonal), g = bp
            for (int i = tid; i < np * pitch; i += BT) S[i] = 0.0;
            if (tid == 0) s_flag = 1;
            __syncthreads();
            for (int i = tid; i < 36 * NA; i += BT) {
                int a = i / 36, r = (i % 36) / 6, c2 = i % 6;
                S[(6 * a + r) * pitch + 6 * a + c2] = Hpp[i] + (r == c2 ? lambda : 0.0);
            }
            // ---- V^-1 per landmark and W V^-1 for the landmark's groups (one pass: the thread that inverts a landmark's block
            //      multiplies it into the landmark's own groups, so the inverse never makes a round trip through memory)
            for (int l = tid; l < L; l += BT) {
                int s0 = l_off[l], s1 = l_off[l + 1];
                if (s0 == s1) continue;
                const int g0 = lg_off[l], g1 = lg_off[l + 1];
                double D[9], Di[9];
#pragma unroll
                for (int x = 0; x < 9; x++) D[x] = Hll[9 * (size_t)l + x];
                D[0] += lambda; D[4] += lambda; D[8] += lambda;
                if (!gd::inv3(D, Di)) s_flag = 0;
#pragma unroll
                for (int x = 0; x < 9; x++) Dinv[9 * (size_t)l + x] = Di[x];
                for (int e = g0; e < g1; e++) {
                    const double2 *W2 = reinterpret_cast<const double2 *>(Hpl + 18 * (size_t)e);
                    double W[18];
#pragma unroll
                    for (int x = 0; x < 9; x++) { double2 t = W2[x]; W[2 * x] = t.x; W[2 * x + 1] = t.y; }
                    double X[18];
#pragma unroll
                    for (int x = 0; x < 6; x++)
#pragma unroll
                        for (int y = 0; y < 3; y++) X[x * 3 + y] = W[x * 3] * Di[y] + W[x * 3 + 1] * Di[3 + y] + W[x * 3 + 2] * Di[6 + y];
                    double2 *O = reinterpret_cast<double2 *>(WD + 18 * (size_t)e);
#pragma unroll
                    for (int x = 0; x < 9; x++) O[x] = make_double2(X[2 * x], X[2 * x + 1]);
                }
            }
            __syncthreads();
            // ---- chunk partials: 4 threads per chunk, thread (qr, qc) accumulates the 3x3 quadrant
            //      sum_pairs (W V^-1)_e1[3qr.., :] (W_e2[3qc.., :])^T in registers
            {
                const int32_t *pr_e1 = A.pr_e1 + P.pair0, *pr_e2 = A.pr_e2 + P.pair0;
                const int32_t *ch_off = A.ch_off + P.choff0;
                double *part = A.part + 36 * (size_t)P.part0;
                for (int t = tid; t < 4 * P.nch; t += BT) {
                    int ch = t >> 2, qr = (t >> 1) & 1, qc = t & 1;
                    double a00 = 0, a01 = 0, a02 = 0, a10 = 0, a11 = 0, a12 = 0, a20 = 0, a21 = 0, a22 = 0;
                    // software pipeline: the indices of pair p+2 and the 2 x 9 operands of pair p+1 are in flight while pair p
                    // is accumulated (every load is an L2 round trip; the chain index -> operand -> FMA is what costs)
                    const int p0 = ch_off[ch], p1 = ch_off[ch + 1];
                    int e1n = pr_e1[p0], e2n = pr_e2[p0];
                    const double *X = WD + 18 * (size_t)e1n + 9 * qr, *Y = Hpl + 18 * (size_t)e2n + 9 * qc;
                    double xn[9], yn[9];
#pragma unroll
                    for (int u = 0; u < 9; u++) { xn[u] = X[u]; yn[u] = Y[u]; }
                    if (p0 + 1 < p1) { e1n = pr_e1[p0 + 1]; e2n = pr_e2[p0 + 1]; }
                    for (int p = p0; p < p1; p++) {
                        double x0 = xn[0], x1 = xn[1], x2 = xn[2], x3 = xn[3], x4 = xn[4], x5 = xn[5], x6 = xn[6], x7 = xn[7], x8 = xn[8];
                        double y0 = yn[0], y1 = yn[1], y2 = yn[2], y3 = yn[3], y4 = yn[4], y5 = yn[5], y6 = yn[6], y7 = yn[7], y8 = yn[8];
                        if (p + 1 < p1) {
                            X = WD + 18 * (size_t)e1n + 9 * qr; Y = Hpl + 18 * (size_t)e2n + 9 * qc;
#pragma unroll
                            for (int u = 0; u < 9; u++) { xn[u] = X[u]; yn[u] = Y[u]; }
                            if (p + 2 < p1) { e1n = pr_e1[p + 2]; e2n = pr_e2[p + 2]; }
                        }
                        a00 += x0 * y0 + x1 * y1 + x2 * y2; a01 += x0 * y3 + x1 * y4 + x2 * y5; a02 += x0 * y6 + x1 * y7 + x2 * y8;
                        a10 += x3 * y0 + x4 * y1 + x5 * y2; a11 += x3 * y3 + x4 * y4 + x5 * y5; a12 += x3 * y6 + x4 * y7 + x5 * y8;
                        a20 += x6 * y0 + x7 * y1 + x8 * y2; a21 += x6 * y3 + x7 * y4 + x8 * y5; a22 += x6 * y6 + x7 * y7 + x8 * y8;
                    }
                    double *O = part + 36 * (size_t)ch + 18 * qr + 3 * qc;     // row-major 6x6 record
                    O[0] = a00; O[1] = a01; O[2] = a02; O[6] = a10; O[7] = a11; O[8] = a12; O[12] = a20; O[13] = a21; O[14] = a22;
                }
            }
            __syncthreads();
            // ---- S_ij -= sum of the block's chunk partials (36 threads own a block; coalesced, fixed order => deterministic)
            {
                const int32_t *blk_i = A.blk_i + P.blk0, *blk_j = A.blk_j + P.blk0, *blk_ch = A.blk_ch + P.bch0;
                const double *part = A.part + 36 * (size_t)P.part0;
                for (int t = tid; t < P.nblk * 36; t += BT) {
                    int bk = t / 36, ent = t - 36 * bk, r = ent / 6, c2 = ent - 6 * r;
                    double sum = 0;
                    const double *C = part + ent;
                    for (int s = blk_ch[bk]; s < blk_ch[bk + 1]; s++) sum += C[36 * (size_t)s];
                    int i = blk_i[bk], j = blk_j[bk];
                    double v = S[(6 * i + r) * pitch + 6 * j + c2] - sum;
                    S[(6 * i + r) * pitch + 6 * j + c2] = v;
                    if (i != j) S[(6 * j + c2) * pitch + 6 * i + r] = v;
                }
            }
            // ---- g_i = bp_i - sum_e (W V^-1)_e bl
            for (int a = warp; a < NA; a += BT / 32) {
                double s6[6] = {0, 0, 0, 0, 0, 0};
                for (int s = pg_off[a] + lane; s < pg_off[a + 1]; s += 32) {
                    int e = pg_groups[s];
                    const double *O = WD + 18 * (size_t)e, *b3 = bl + 3 * (size_t)g_lm[e];
#pragma unroll
                    for (int x = 0; x < 6; x++) s6[x] += O[x * 3] * b3[0] + O[x * 3 + 1] * b3[1] + O[x * 3 + 2] * b3[2];
                }
#pragma unroll
                for (int x = 0; x < 6; x++) s6[x] = gd::warp_sum(s6[x]);
                if (lane == 0)
#pragma unroll
                    for (int x = 0; x < 6; x++) g[6 * a + x] = bp[6 * a + x] - s6[x];
            }
            __syncthreads();
            bool ok = s_flag != 0;
            if (ok) ok = S2 ? block_ldlt_solve_pp<BT>(S, S2, pitch, np, g, xp, tr, tmp) : block_ldlt_solve<BT>(S, pitch, np, g, xp, tr, tmp, &s_piv);
            st_sol++;
            if (!ok) { for (int i = tid; i < np; i += BT) xp[i] = 0.0; __syncthreads(); }
            // ---- trial poses first, then ONE landmark pass: back-substitution, trial landmark, scale term and the robust chi2
            //      of the trial state (the thread that moves a landmark evaluates the landmark's edges: same thread -> landmark
            //      map and edge order as the linearisation's chi2, so the sum is formed the same way; one pass and one barrier pair fewer per trial)
            for (int a = tid; a < NA; a += BT) gd::se3_oplus(poseA + 7 * a, xp + 6 * a, poseT + 7 * a);
            __syncthreads();
            double sc = 0, acc_t = 0;
            for (int l = tid; l < L; l += BT) {
                int s0 = l_off[l], s1 = l_off[l + 1];
                if (s0 == s1) continue;
                double c3[3] = {bl[3 * (size_t)l], bl[3 * (size_t)l + 1], bl[3 * (size_t)l + 2]};
                double x3[3] = {0, 0, 0};
                if (ok) {
                    for (int e = lg_off[l]; e < lg_off[l + 1]; e++) {
                        const double *W = Hpl + 18 * (size_t)e, *xx = xp + 6 * g_pose[e];
#pragma unroll
                        for (int y = 0; y < 3; y++) {
                            double sm2 = 0;
#pragma unroll
                            for (int x = 0; x < 6; x++) sm2 += W[x * 3 + y] * xx[x];
                            c3[y] -= sm2;
                        }
                    }
                    const double *Di = Dinv + 9 * (size_t)l;
#pragma unroll
                    for (int x = 0; x < 3; x++) x3[x] = Di[x * 3] * c3[0] + Di[x * 3 + 1] * c3[1] + Di[x * 3 + 2] * c3[2];
                }
                double lt[3];
#pragma unroll
                for (int x = 0; x < 3; x++) {
                    xl[3 * (size_t)l + x] = x3[x];
                    lt[x] = lms[3 * (size_t)l + x] + x3[x];
                    lmT[3 * (size_t)l + x] = lt[x];
                    sc += x3[x] * (lambda * x3[x] + bl[3 * (size_t)l + x]);
                }
                for (int s = s0; s < s1; s++) {
                    int e = l_edges[s], cam = edge_cam[e];
                    double er[2], a[3], c[3];
                    gd::ba_error(poseT + 7 * edge_p[e], A.ext[cam], A.K[cam], lt, edge_uv + 2 * e, er, a, c);
                    double e2 = er[0] * er[0] + er[1] * er[1], r0, r1;
                    gd::huber(e2, hd, r0, r1);
                    acc_t += r0;
                }
            }
            double scl = block_sum<BT>(sc, red);
            double tmpchi = block_sum<BT>(acc_t, red);
            if (warp == 0) {
                double scp = 0;
                for (int i = lane; i < np; i += 32) scp += xp[i] * (lambda * xp[i] + bp[i]);
                scp = gd::warp_sum(scp);
                if (lane == 0) {
                double scale = scp + scl + 1e-3;
                double tc = ok ? tmpchi : DBL_MAX;
                double r = (cur - tc) / scale;
                gd::LmCtl lm = {s_lambda, s_ni};
                int acc2 = gd::lm_accept(lm, r, tc) ? 1 : 0;
                s_lambda = lm.lambda; s_ni = lm.ni; s_rho = r; s_q = acc2;
                if (acc2) s_cur = tc;
                }
            }
            __syncthreads();
            rho = s_rho;
            if (s_q) {
                cur = s_cur;
                for (int i = tid; i < 7 * NA; i += BT) poseA[i] = poseT[i];
                for (int i = tid; i < 3 * L; i += BT) lms[i] = lmT[i];
            }
            __syncthreads();
            q++; st_tr++;
        } while (rho < 0 && q < 10);
        st_it++;
        if (q == 10 || rho == 0) stop = true;
        if (tid == 0) s_cur = cur;
        __syncthreads();
    }
    // ---- outputs: per-edge chi2 as g2o leaves it (errors of the LAST evaluated state, possibly a rejected trial)
    for (int e = tid; e < E; e += BT) {
        int cam = edge_cam[e];
        double er[2], a[3], c[3];
        gd::ba_error(poseT + 7 * edge_p[e], A.ext[cam], A.K[cam], lmT + 3 * (size_t)edge_l[e], edge_uv + 2 * e, er, a, c);
        A.edge_chi2[P.e0 + e] = er[0] * er[0] + er[1] * er[1];
    }
    for (int i = tid; i < 7 * NA; i += BT) {
        int a = i / 7;
        A.poses[7 * (size_t)(P.pose0 + A.act_pose[P.act0 + a]) + (i - 7 * a)] = poseA[i];
    }
    if (tid == 0 && A.stats) {
        svs_ba_stats &s = A.stats[blockIdx.x];
        s.iterations = st_it; s.trials = st_tr; s.linearizations = st_lin; s.solves = st_sol;
        s.lambda = s_lambda; s.chi2 = (st_it > 0) ? s_cur : 0.0; s.chi2_init = chi_init;
    }
}


// ---------------------------------------------------------------------------------------------
// Window structure on the device.  The lists k_ba_window walks (active poses, edges by pose, (landmark, pose) groups by
// landmark and by pose, (group, group) pairs per 6x6 block cut into chunks) used to be built by the host for every window
// of every step (0.3 ms of one core per window: the largest host cost of a keyframe).  One CTA per window builds the SAME
// lists from the raw edge arrays, for the edge order the pipeline emits: landmark-major, keyframe-ascending inside a
// landmark (svs_ba_optimize checks this and keeps the host construction for any other order).  Every step is a stable
// counting sort or a prefix sum whose result does not depend on the thread schedule, so the lists — and with them the
// summation order of k_ba_window — are identical to the host construction's.
//   * a window has <= 32 keyframes here, so a landmark's keyframes are a 32-bit mask; the pairs of block (i, j) are the
//     landmarks whose mask has bits i and j, in landmark order: ballots + popcounts, no sort
//   * the lists are laid out per window at offsets the host derives from UPPER BOUNDS (groups <= E, pairs <= E (N + 1) / 2,
//     chunks <= pairs / BA_CH + blocks), so no size has to travel back to the host before k_ba_window is launched
#define BB_T 512
struct BaBuildArgs {
    BaProb *probs;
    const int32_t *edge_kf, *edge_lm;
    int32_t *act_pose, *edge_p, *l_off, *l_edges, *p_off, *p_edges;
    int32_t *blk_i, *blk_j, *blk_ch, *pr_e1, *pr_e2, *ch_blk, *ch_off;
    int32_t *lg_off, *g_lm, *g_pose, *pg_off, *pg_groups;
    uint32_t *lm_mask;
};

// stable counting sort of n elements by key(e) in [0, NA), NA <= 32: off_out[0 .. NA] = bucket offsets, list_out = element
// indices bucket by bucket, ascending inside a bucket.  Warp w owns a contiguous range; lane a carries bucket a's counters.
template <class KeyF>
__device__ void bb_bucket_sort(KeyF key, int n, int NA, int32_t *off_out, int32_t *list_out, int (*wcnt)[32], int *s_off)
{
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, NW = BB_T / 32;
    const int per = ((n + NW - 1) / NW + 31) & ~31;
    const int lo = min(n, warp * per), hi = min(n, lo + per);
    int cnt = 0;
    for (int e0 = lo; e0 < hi; e0 += 32) {
        const int e = e0 + lane;
        const int k = e < hi ? key(e) : -1;
        for (int a = 0; a < NA; a++) { const unsigned bal = __ballot_sync(0xffffffffu, k == a); if (lane == a) cnt += __popc(bal); }
    }
    wcnt[warp][lane] = cnt;
    __syncthreads();
    if (warp == 0) {
        int tot = 0;
        for (int w = 0; w < NW; w++) tot += wcnt[w][lane];
        if (lane >= NA) tot = 0;
        int incl = tot;
        for (int o = 1; o < 32; o <<= 1) { const int v = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += v; }
        s_off[lane] = incl - tot;
        if (lane == 31) s_off[32] = incl;
    }
    __syncthreads();
    if (tid <= NA) off_out[tid] = s_off[tid];
    int base = s_off[lane];
    for (int w = 0; w < warp; w++) base += wcnt[w][lane];
    for (int e0 = lo; e0 < hi; e0 += 32) {
        const int e = e0 + lane;
        const int k = e < hi ? key(e) : -1;
        for (int a = 0; a < NA; a++) {
            const unsigned bal = __ballot_sync(0xffffffffu, k == a);
            const int ba = __shfl_sync(0xffffffffu, base, a);
            if (k == a) list_out[ba + __popc(bal & ((1u << lane) - 1u))] = e;
            if (lane == a) base += __popc(bal);
        }
    }
    __syncthreads();
}

__global__ void __launch_bounds__(BB_T) k_ba_build(BaBuildArgs B)
{
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, NW = BB_T / 32;
    BaProb P = B.probs[blockIdx.x];
    const int N = P.NA /* keyframes of the window on entry */, L = P.L, E = P.E;
    const int32_t *ekf = B.edge_kf + P.e0, *elm = B.edge_lm + P.e0;
    int32_t *edge_p = B.edge_p + P.e0, *l_edges = B.l_edges + P.e0, *p_edges = B.p_edges + P.e0;
    int32_t *l_off = B.l_off + P.loff0, *p_off = B.p_off + P.poff0, *lg_off = B.lg_off + P.lgoff0, *pg_off = B.pg_off + P.pgoff0;
    int32_t *g_lm = B.g_lm + P.grp0, *g_pose = B.g_pose + P.grp0, *pg_groups = B.pg_groups + P.grp0;
    uint32_t *mask = B.lm_mask + P.lm0;
    __shared__ unsigned s_act;
    __shared__ int s_wsum[BB_T / 32], s_base, s_off[33], s_wcnt[BB_T / 32][32];
    __shared__ int s_bcnt[528], s_bstart[528], s_bch[528], s_bidx[528], s_tot[3];
    // ---- active poses
    if (tid == 0) { s_act = 0; s_base = 0; }
    for (int l = tid; l < L; l += BB_T) mask[l] = 0;
    __syncthreads();
    {
        unsigned act = 0;
        for (int e = tid; e < E; e += BB_T) act |= 1u << ekf[e];
        act = __reduce_or_sync(0xffffffffu, act);
        if (lane == 0 && act) atomicOr(&s_act, act);
    }
    __syncthreads();
    const unsigned act = s_act;
    const int NA = __popc(act);
    if (tid < N && ((act >> tid) & 1u)) B.act_pose[P.act0 + __popc(act & ((1u << tid) - 1u))] = tid;
    auto pidx = [act](int k) { return __popc(act & ((1u << k) - 1u)); };
    for (int e = tid; e < E; e += BB_T) { edge_p[e] = pidx(ekf[e]); l_edges[e] = e; }
    // ---- groups = runs of equal (landmark, pose) in the sorted edge list; CSR of edges and groups by landmark; pose masks
    for (int t0 = 0; t0 < E; t0 += BB_T) {
        const int e = t0 + tid;
        int flag = 0, l = -1, pp = -1, lp = -1;
        if (e < E) {
            l = elm[e]; pp = pidx(ekf[e]);
            if (e == 0) flag = 1;
            else { lp = elm[e - 1]; flag = (lp != l || pidx(ekf[e - 1]) != pp) ? 1 : 0; }
        }
        int incl = flag;
        for (int o = 1; o < 32; o <<= 1) { const int v = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += v; }
        if (lane == 31) s_wsum[warp] = incl;
        __syncthreads();
        int wbase = s_base;
        for (int w = 0; w < warp; w++) wbase += s_wsum[w];
        const int gid = wbase + incl - 1;
        if (flag) {
            g_lm[gid] = l; g_pose[gid] = pp;
            atomicOr(&mask[l], 1u << pp);
            if (lp != l) for (int ll = lp + 1; ll <= l; ll++) { lg_off[ll] = gid; l_off[ll] = e; }
        }
        __syncthreads();
        if (tid == 0) { int tot = s_base; for (int w = 0; w < NW; w++) tot += s_wsum[w]; s_base = tot; }
        __syncthreads();
    }
    const int G = s_base;
    {
        const int last = E > 0 ? elm[E - 1] : -1;
        for (int ll = last + 1 + tid; ll <= L; ll += BB_T) { lg_off[ll] = G; l_off[ll] = E; }
    }
    __syncthreads();
    // ---- edges by pose, groups by pose (stable)
    bb_bucket_sort([edge_p](int e) { return edge_p[e]; }, E, NA, p_off, p_edges, s_wcnt, s_off);
    bb_bucket_sort([g_pose](int g) { return g_pose[g]; }, G, NA, pg_off, pg_groups, s_wcnt, s_off);
    // ---- pairs per upper block (i <= j): the landmarks seen from both poses, in landmark order
    const int NB = NA * (NA + 1) / 2;
    for (int q = warp; q < NB; q += NW) {
        int i = 0, r = q;
        while (r >= NA - i) { r -= NA - i; i++; }
        const int j = i + r;
        const unsigned need = (1u << i) | (1u << j);
        int cnt = 0;
        for (int l0 = 0; l0 < L; l0 += 32) {
            const int l = l0 + lane;
            const unsigned m = l < L ? mask[l] : 0u;
            cnt += __popc(__ballot_sync(0xffffffffu, (m & need) == need));
        }
        if (lane == 0) s_bcnt[q] = cnt;
    }
    __syncthreads();
    if (tid == 0) {
        int run = 0, nch = 0, nblk = 0, q = 0;
        for (int i = 0; i < NA; i++)
            for (int j = i; j < NA; j++, q++) {
                const int cn = s_bcnt[q];
                if (cn > 0) {
                    s_bidx[q] = nblk; s_bstart[q] = run; s_bch[q] = nch;
                    B.blk_i[P.blk0 + nblk] = i; B.blk_j[P.blk0 + nblk] = j; B.blk_ch[P.bch0 + nblk] = nch;
                    nch += (cn + BA_CH - 1) / BA_CH; run += cn; nblk++;
                } else s_bidx[q] = -1;
            }
        B.blk_ch[P.bch0 + nblk] = nch;
        B.ch_off[P.choff0 + nch] = run;
        s_tot[0] = nblk; s_tot[1] = nch; s_tot[2] = run;
    }
    __syncthreads();
    for (int q = tid; q < NB; q += BB_T) {
        if (s_bidx[q] < 0) continue;
        const int cn = s_bcnt[q];
        for (int c0 = 0, cc = 0; c0 < cn; c0 += BA_CH, cc++) { B.ch_blk[P.ch0 + s_bch[q] + cc] = s_bidx[q]; B.ch_off[P.choff0 + s_bch[q] + cc] = s_bstart[q] + c0; }
    }
    for (int q = warp; q < NB; q += NW) {
        if (s_bcnt[q] == 0) continue;
        int i = 0, r = q;
        while (r >= NA - i) { r -= NA - i; i++; }
        const int j = i + r;
        const unsigned need = (1u << i) | (1u << j), lo_i = (1u << i) - 1u, lo_j = (1u << j) - 1u;
        int run = s_bstart[q];
        for (int l0 = 0; l0 < L; l0 += 32) {
            const int l = l0 + lane;
            const unsigned m = l < L ? mask[l] : 0u;
            const bool has = (m & need) == need;
            const unsigned bal = __ballot_sync(0xffffffffu, has);
            if (has) {
                const int pos = P.pair0 + run + __popc(bal & ((1u << lane) - 1u)), g0 = lg_off[l];
                B.pr_e1[pos] = g0 + __popc(m & lo_i); B.pr_e2[pos] = g0 + __popc(m & lo_j);
            }
            run += __popc(bal);
        }
    }
    if (tid == 0) {
        P.NA = NA; P.G = G; P.nblk = s_tot[0]; P.nch = s_tot[1];
        B.probs[blockIdx.x] = P;
    }
}

// ---------------------------------------------------------------------------------------------
// Host-side workspace of svs_ba_optimize, owned by the context (svs_ctx::ba_ws) and reused by every call.
struct alignas(128) BaBuilt {      // one problem's structure pieces; aligned: neighbouring problems are filled by different threads
    std::vector<int32_t> act_pose, l_off, p_off, blk_i, blk_j, blk_ch, pr_e1, pr_e2, ch_blk, ch_off, lg_off, g_lm, g_pose, pg_off, pg_groups;
    int bad = 0;
};
struct BaSeg { const void *src; size_t bytes; size_t off; };
struct BaHostWs {
    std::vector<BaProb> probs;
    std::vector<BaBuilt> built;
    std::vector<int32_t> edge_p, l_edges, p_edges;
    std::vector<BaSeg> segs;
};
void svs_i_ba_ws_free(void *p) { delete static_cast<BaHostWs *>(p); }

static int ba_optimize_on_stream(svs_ctx *c, int n_prob, const int32_t *kf_off, double *poses, const int32_t *lm_off, double *lms,
                                 const int32_t *e_off, const int32_t *edge_kf, const int32_t *edge_lm, const uint8_t *edge_cam,
                                 const double *edge_uv, const double K_left[4], const double K_right[4], const double ext_left[7],
                                 const double ext_right[7], double huber_delta, int max_iter, int jacobian_mode,
                                 double *edge_chi2_out, svs_ba_stats *stats)
{
    if (!c || n_prob < 0 || !kf_off || !lm_off || !e_off || !K_left || !K_right || !ext_left || !ext_right) return SVS_ERR_ARG;
    if (n_prob == 0) return SVS_OK;
    const int sumN = kf_off[n_prob], sumL = lm_off[n_prob], sumE = e_off[n_prob];
    if (sumE > 0 && (!edge_kf || !edge_lm || !edge_cam || !edge_uv || !edge_chi2_out || !poses || !lms)) return SVS_ERR_ARG;
    SVS_CUDA(c, cudaSetDevice(c->device));
    auto now_s = [] { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); };
    const double t_begin = now_s();

    BaArgs A;
    size_t sumG = 0, sumCh = 0, max_smem = 0;
    long long S_tot = 0;
    double t_built = t_begin;
    const size_t smem_cap = 200 * 1024;
    BaHostWs *ws = static_cast<BaHostWs *>(c->ba_ws);
    if (!ws) { ws = new (std::nothrow) BaHostWs(); if (!ws) SVS_FAIL(c, SVS_ERR_CUDA, "ba: out of host memory"); c->ba_ws = ws; }
    const int omp_team = c->host_threads > 0 ? c->host_threads : omp_get_max_threads();

    // ---- where is the window structure built?  On the device (k_ba_build) when every window has <= 32 keyframes, fits the
    // shared-memory solver and lists its edges landmark-major / keyframe-ascending (what Backend::Optimize emits); on the
    // host otherwise.  Both give the same lists.
    bool dev_build = !getenv("SVS_BA_HOST_BUILD");
    if (dev_build) {
        int ok = 1;
#pragma omp parallel for schedule(dynamic, 1) num_threads(omp_team) reduction(& : ok)
        for (int b = 0; b < n_prob; b++) {
            const int N = kf_off[b + 1] - kf_off[b], L = lm_off[b + 1] - lm_off[b], E = e_off[b + 1] - e_off[b], e0 = e_off[b];
            int good = (N >= 0 && N <= 32 && L >= 0 && E >= 0 && ba_smem_need(N, 1) <= smem_cap) ? 1 : 0;
            int lp = -1, kp = -1;
            for (int e = 0; good && e < E; e++) {
                const int k = edge_kf[e0 + e], l = edge_lm[e0 + e];
                if (k < 0 || k >= N || l < 0 || l >= L || l < lp || (l == lp && k < kp)) good = 0;
                lp = l; kp = k;
            }
            ok &= good;
        }
        dev_build = ok != 0;
    }
    if (dev_build) {
        std::vector<BaProb> &probs = ws->probs;
        probs.resize(n_prob);
        long long a_pair = 0, a_ch = 0, a_blk = 0;
        for (int b = 0; b < n_prob; b++) {
            const int N = kf_off[b + 1] - kf_off[b], L = lm_off[b + 1] - lm_off[b], E = e_off[b + 1] - e_off[b];
            BaProb &P = probs[b];
            P.NA = N; P.L = L; P.E = E; P.nblk = 0; P.G = 0; P.nch = 0;
            P.pose0 = kf_off[b]; P.lm0 = lm_off[b]; P.e0 = e_off[b];
            P.act0 = kf_off[b]; P.loff0 = lm_off[b] + b; P.poff0 = kf_off[b] + b; P.lgoff0 = lm_off[b] + b; P.pgoff0 = kf_off[b] + b;
            P.grp0 = e_off[b];
            const long long NB = (long long)N * (N + 1) / 2, PB = ((long long)E * (N + 1) + 1) / 2;
            P.blk0 = (int)a_blk; P.bch0 = (int)(a_blk + b); P.pair0 = (int)a_pair; P.ch0 = (int)a_ch; P.choff0 = (int)(a_ch + b);
            P.part0 = a_ch; P.S_off = -1;
            a_blk += NB; a_pair += PB; a_ch += PB / BA_CH + NB + 1;
            max_smem = std::max(max_smem, ba_smem_need(N, 2) <= smem_cap ? ba_smem_need(N, 2) : ba_smem_need(N, 1));
            if (a_pair > 0x7fffffffLL / 2 || a_ch > 0x7fffffffLL / 2) SVS_FAIL(c, SVS_ERR_CAPACITY, "ba: too many windows in one call");
        }
        sumG = (size_t)sumE; sumCh = (size_t)a_ch; S_tot = 0;
        t_built = now_s();
        // inputs host -> device
        std::vector<BaSeg> &segs = ws->segs;
        segs.clear();
        size_t tot = 0;
        auto add = [&](const void *p, size_t bytes) { size_t o = tot; segs.push_back({p, bytes, o}); tot = align_up(tot + bytes, 16); return o; };
        const size_t o_probs = add(probs.data(), probs.size() * sizeof(BaProb));
        const size_t o_ek = add(edge_kf, (size_t)sumE * 4), o_el = add(edge_lm, (size_t)sumE * 4), o_ec = add(edge_cam, (size_t)sumE);
        const size_t o_uv = add(edge_uv, (size_t)sumE * 16), o_pose = add(poses, (size_t)sumN * 56), o_lm = add(lms, (size_t)sumL * 24);
        SVS_CUDA(c, c->h_in.reserve(tot + 16));
        SVS_CUDA(c, c->d_in2.reserve(tot + 16));
        uint8_t *hb = c->h_in.as<uint8_t>(), *db = c->d_in2.as<uint8_t>();
        {
            struct Piece { const uint8_t *src; uint8_t *dst; size_t n; };
            static thread_local std::vector<Piece> pieces;
            pieces.clear();
            for (const BaSeg &sg : segs)
                for (size_t o = 0; o < sg.bytes; o += 262144) pieces.push_back({(const uint8_t *)sg.src + o, hb + sg.off + o, std::min<size_t>(262144, sg.bytes - o)});
            const int n_pieces = (int)pieces.size();
            const Piece *pc_ = pieces.data();
#pragma omp parallel for schedule(dynamic, 1) num_threads(omp_team)
            for (int i = 0; i < n_pieces; i++) memcpy(pc_[i].dst, pc_[i].src, pc_[i].n);
        }
        SVS_CUDA(c, cudaMemcpyAsync(db, hb, tot, cudaMemcpyHostToDevice, c->stream));
        // structure lists: device scratch at upper-bound offsets
        size_t st = 0;
        auto res = [&](size_t n_i32) { size_t o = st; st = align_up(st + n_i32 * 4, 16); return o; };
        const size_t s_act = res(sumN), s_ep = res(sumE), s_lo = res((size_t)sumL + n_prob), s_le = res(sumE), s_po = res((size_t)sumN + n_prob),
                     s_pe = res(sumE), s_bi = res(a_blk), s_bj = res(a_blk), s_bc = res((size_t)a_blk + n_prob), s_p1 = res(a_pair), s_p2 = res(a_pair),
                     s_cb = res(a_ch), s_co = res((size_t)a_ch + n_prob), s_lg = res((size_t)sumL + n_prob), s_gl = res(sumE), s_gp = res(sumE),
                     s_pgo = res((size_t)sumN + n_prob), s_pgg = res(sumE), s_mask = res(sumL);
        SVS_CUDA(c, c->d_tmp8.reserve(st + 16));
        uint8_t *sb = c->d_tmp8.as<uint8_t>();
        auto I = [sb](size_t o) { return reinterpret_cast<int32_t *>(sb + o); };
        BaBuildArgs Bd;
        Bd.probs = reinterpret_cast<BaProb *>(db + o_probs);
        Bd.edge_kf = reinterpret_cast<const int32_t *>(db + o_ek); Bd.edge_lm = reinterpret_cast<const int32_t *>(db + o_el);
        Bd.act_pose = I(s_act); Bd.edge_p = I(s_ep); Bd.l_off = I(s_lo); Bd.l_edges = I(s_le); Bd.p_off = I(s_po); Bd.p_edges = I(s_pe);
        Bd.blk_i = I(s_bi); Bd.blk_j = I(s_bj); Bd.blk_ch = I(s_bc); Bd.pr_e1 = I(s_p1); Bd.pr_e2 = I(s_p2); Bd.ch_blk = I(s_cb); Bd.ch_off = I(s_co);
        Bd.lg_off = I(s_lg); Bd.g_lm = I(s_gl); Bd.g_pose = I(s_gp); Bd.pg_off = I(s_pgo); Bd.pg_groups = I(s_pgg);
        Bd.lm_mask = reinterpret_cast<uint32_t *>(sb + s_mask);
        SVS_KERNEL(c, KID_BA_BUILD, k_ba_build<<<n_prob, BB_T, 0, c->stream>>>(Bd));
        A.probs = Bd.probs;
        A.poses = reinterpret_cast<double *>(db + o_pose); A.lms = reinterpret_cast<double *>(db + o_lm);
        A.act_pose = Bd.act_pose; A.edge_p = Bd.edge_p; A.edge_l = Bd.edge_lm;
        A.edge_cam = db + o_ec; A.edge_uv = reinterpret_cast<double *>(db + o_uv);
        A.l_off = Bd.l_off; A.l_edges = Bd.l_edges; A.p_off = Bd.p_off; A.p_edges = Bd.p_edges;
        A.blk_i = Bd.blk_i; A.blk_j = Bd.blk_j; A.blk_ch = Bd.blk_ch; A.pr_e1 = Bd.pr_e1; A.pr_e2 = Bd.pr_e2; A.ch_blk = Bd.ch_blk; A.ch_off = Bd.ch_off;
        A.lg_off = Bd.lg_off; A.g_lm = Bd.g_lm; A.g_pose = Bd.g_pose; A.pg_off = Bd.pg_off; A.pg_groups = Bd.pg_groups;
    } else {
        // ---- host-side structure: active poses, CSR by landmark / pose, per-block pair lists.  The problems are independent:
        // each is built into its own vectors by an OpenMP team (the windows of one step come from different streams); the pieces
        // are then copied ONCE, straight into the pinned staging buffer at prefix offsets.  All vectors live in a per-context
        // workspace and the per-thread temporaries are thread_local, so a steady-state call allocates nothing (fresh vectors cost
        // page faults under the process-wide mmap lock, which serialises the team).
        std::vector<BaProb> &probs = ws->probs;
        std::vector<BaBuilt> &built = ws->built;
        std::vector<int32_t> &edge_p = ws->edge_p, &l_edges = ws->l_edges, &p_edges = ws->p_edges;
        probs.resize(n_prob);
        if ((int)built.size() < n_prob) built.resize(n_prob);
        edge_p.resize(sumE); l_edges.resize(sumE); p_edges.resize(sumE);
    #pragma omp parallel for schedule(dynamic, 1) num_threads(omp_team)
        for (int b = 0; b < n_prob; b++) {
            int N = kf_off[b + 1] - kf_off[b], L = lm_off[b + 1] - lm_off[b], E = e_off[b + 1] - e_off[b], e0 = e_off[b];
            BaProb &P = probs[b];
            BaBuilt &B = built[b];
            B.act_pose.clear(); B.blk_i.clear(); B.blk_j.clear(); B.blk_ch.clear(); B.ch_blk.clear(); B.ch_off.clear();
            B.lg_off.clear(); B.g_lm.clear(); B.g_pose.clear(); B.bad = 0;
            static thread_local std::vector<int> pidx, cnt, fill, pc, pgc, bcount, bstart;
            pidx.assign(N, -1);
            for (int e = 0; e < E; e++) {
                int k = edge_kf[e0 + e], l = edge_lm[e0 + e];
                if (k < 0 || k >= N || l < 0 || l >= L) { B.bad = 1; break; }
                pidx[k] = 0;
            }
            if (B.bad) continue;
            int NA = 0;
            for (int k = 0; k < N; k++) if (pidx[k] == 0) { pidx[k] = NA++; B.act_pose.push_back(k); }
            P.NA = NA; P.L = L; P.E = E; P.pose0 = kf_off[b]; P.lm0 = lm_off[b]; P.e0 = e0;
            for (int e = 0; e < E; e++) edge_p[e0 + e] = pidx[edge_kf[e0 + e]];
            // CSR by landmark; inside a landmark the edges are listed pose-ascending (stable: creation order inside a pose),
            // so that the edges of one GROUP = (landmark, pose) form a run
            cnt.assign(L + 1, 0);
            for (int e = 0; e < E; e++) cnt[edge_lm[e0 + e] + 1]++;
            for (int l = 0; l < L; l++) cnt[l + 1] += cnt[l];
            B.l_off.assign(cnt.begin(), cnt.end());
            fill.assign(cnt.begin(), cnt.end() - 1);
            for (int e = 0; e < E; e++) l_edges[e0 + fill[edge_lm[e0 + e]]++] = e;
            for (int l = 0; l < L; l++) {      // stable insertion sort: a landmark has a handful of edges
                int32_t *a = l_edges.data() + e0 + cnt[l];
                const int m = cnt[l + 1] - cnt[l];
                for (int i = 1; i < m; i++) {
                    const int32_t v = a[i], pv = edge_p[e0 + v];
                    int j = i - 1;
                    while (j >= 0 && edge_p[e0 + a[j]] > pv) { a[j + 1] = a[j]; j--; }
                    a[j + 1] = v;
                }
            }
            // CSR by active pose
            pc.assign(NA + 1, 0);
            for (int e = 0; e < E; e++) pc[edge_p[e0 + e] + 1]++;
            for (int a = 0; a < NA; a++) pc[a + 1] += pc[a];
            B.p_off.assign(pc.begin(), pc.end());
            fill.assign(pc.begin(), pc.end() - 1);
            for (int e = 0; e < E; e++) p_edges[e0 + fill[edge_p[e0 + e]]++] = e;
            // groups, landmark-major / pose-ascending; CSR by landmark and by pose
            int G = 0;
            pgc.assign(NA + 1, 0);
            B.lg_off.reserve(L + 1); B.g_lm.reserve(E); B.g_pose.reserve(E);
            for (int l = 0; l < L; l++) {
                B.lg_off.push_back(G);
                int cur = -1;
                for (int s1 = cnt[l]; s1 < cnt[l + 1]; s1++) {
                    int pp = edge_p[e0 + l_edges[e0 + s1]];
                    if (pp != cur) { cur = pp; B.g_lm.push_back(l); B.g_pose.push_back(pp); pgc[pp + 1]++; G++; }
                }
            }
            B.lg_off.push_back(G);
            P.G = G;
            for (int a = 0; a < NA; a++) pgc[a + 1] += pgc[a];
            B.pg_off.assign(pgc.begin(), pgc.end());
            fill.assign(pgc.begin(), pgc.end() - 1);
            B.pg_groups.resize(G);
            for (int g = 0; g < G; g++) B.pg_groups[fill[B.g_pose[g]]++] = g;
            // (group, group) pairs per upper block (i <= j): the groups of a landmark have distinct, ascending poses, so the
            // pairs are (a, b) with a <= b in list order.  Sorted by block id i*NA + j (counting sort; inside a block in
            // landmark order), then cut into chunks of <= BA_CH pairs that never span two blocks
            bcount.assign((size_t)NA * NA + 1, 0);
            const int32_t *lgo = B.lg_off.data(), *gp = B.g_pose.data();
            for (int l = 0; l < L; l++)
                for (int g1 = lgo[l]; g1 < lgo[l + 1]; g1++)
                    for (int g2 = g1; g2 < lgo[l + 1]; g2++) bcount[(size_t)gp[g1] * NA + gp[g2] + 1]++;
            int nblk = 0, run = 0, nch = 0;
            bstart.assign((size_t)NA * NA, 0);
            for (size_t k = 0; k < (size_t)NA * NA; k++) {
                int cn = bcount[k + 1];
                if (cn > 0) {
                    nblk++; bstart[k] = run;
                    B.blk_i.push_back((int)(k / NA)); B.blk_j.push_back((int)(k % NA));
                    B.blk_ch.push_back(nch);
                    for (int c0 = 0; c0 < cn; c0 += BA_CH) { B.ch_blk.push_back(nblk - 1); B.ch_off.push_back(run + c0); nch++; }
                    run += cn;
                }
            }
            B.blk_ch.push_back(nch);
            B.ch_off.push_back(run);
            P.nblk = nblk; P.nch = nch;
            B.pr_e1.resize(run); B.pr_e2.resize(run);
            for (int l = 0; l < L; l++)
                for (int g1 = lgo[l]; g1 < lgo[l + 1]; g1++)
                    for (int g2 = g1; g2 < lgo[l + 1]; g2++) {
                        int q = bstart[(size_t)gp[g1] * NA + gp[g2]]++;
                        B.pr_e1[q] = g1; B.pr_e2[q] = g2;
                    }
        }
        // ---- prefix offsets of the per-problem pieces
        size_t n_act = 0, n_lo = 0, n_po = 0, n_blk = 0, n_bch = 0, n_pr = 0, n_ch = 0, n_co = 0, n_lg = 0, n_g = 0, n_pgo = 0;
        for (int b = 0; b < n_prob; b++) {
            if (built[b].bad) SVS_FAIL(c, SVS_ERR_ARG, "ba: edge index out of range");
            BaProb &P = probs[b];
            const BaBuilt &B = built[b];
            P.act0 = (int)n_act; P.loff0 = (int)n_lo; P.poff0 = (int)n_po; P.blk0 = (int)n_blk; P.bch0 = (int)n_bch; P.pair0 = (int)n_pr;
            P.ch0 = (int)n_ch; P.choff0 = (int)n_co; P.part0 = (long long)n_ch; P.grp0 = (int)n_g; P.lgoff0 = (int)n_lg; P.pgoff0 = (int)n_pgo;
            n_act += B.act_pose.size(); n_lo += B.l_off.size(); n_po += B.p_off.size(); n_blk += B.blk_i.size(); n_bch += B.blk_ch.size();
            n_pr += B.pr_e1.size(); n_ch += B.ch_blk.size(); n_co += B.ch_off.size(); n_lg += B.lg_off.size(); n_g += B.g_lm.size();
            n_pgo += B.pg_off.size();
            const int NA = P.NA;
            if (ba_smem_need(NA, 2) <= smem_cap) { P.S_off = -1; max_smem = std::max(max_smem, ba_smem_need(NA, 2)); }
            else if (ba_smem_need(NA, 1) <= smem_cap) { P.S_off = -1; max_smem = std::max(max_smem, ba_smem_need(NA, 1)); }
            else { P.S_off = S_tot; size_t np = 6 * (size_t)NA; S_tot += (long long)(np * (np | 1)); max_smem = std::max(max_smem, ba_smem_need(NA, 0)); }
        }
        if (max_smem > 220 * 1024) SVS_FAIL(c, SVS_ERR_CAPACITY, "ba: window too large for the single-CTA solver");

        t_built = now_s();
        // ---- pack host -> device: flat arrays as 256 KB pieces, per-problem pieces straight from their vectors
        typedef BaSeg Seg;
        std::vector<Seg> &segs = ws->segs;
        segs.clear();
        size_t tot = 0;
        auto add = [&](const void *p, size_t bytes) { size_t o = tot; if (p) segs.push_back({p, bytes, o}); tot = align_up(tot + bytes, 16); return o; };
        size_t o_probs = add(probs.data(), probs.size() * sizeof(BaProb));
        size_t o_act = add(nullptr, n_act * 4);
        size_t o_ep = add(edge_p.data(), (size_t)sumE * 4);
        size_t o_el = add(edge_lm, (size_t)sumE * 4);
        size_t o_ec = add(edge_cam, (size_t)sumE);
        size_t o_uv = add(edge_uv, (size_t)sumE * 16);
        size_t o_lo = add(nullptr, n_lo * 4);
        size_t o_le = add(l_edges.data(), (size_t)sumE * 4);
        size_t o_po = add(nullptr, n_po * 4);
        size_t o_pe = add(p_edges.data(), (size_t)sumE * 4);
        size_t o_bi = add(nullptr, n_blk * 4);
        size_t o_bj = add(nullptr, n_blk * 4);
        size_t o_bc = add(nullptr, n_bch * 4);
        size_t o_p1 = add(nullptr, n_pr * 4);
        size_t o_p2 = add(nullptr, n_pr * 4);
        size_t o_cb = add(nullptr, n_ch * 4);
        size_t o_co = add(nullptr, n_co * 4);
        size_t o_lg = add(nullptr, n_lg * 4);
        size_t o_gl = add(nullptr, n_g * 4);
        size_t o_gp = add(nullptr, n_g * 4);
        size_t o_pgo = add(nullptr, n_pgo * 4);
        size_t o_pgg = add(nullptr, n_g * 4);
        sumG = n_g; sumCh = n_ch;
        size_t o_pose = add(poses, (size_t)sumN * 56);
        size_t o_lm = add(lms, (size_t)sumL * 24);
        SVS_CUDA(c, c->h_in.reserve(tot + 16));
        SVS_CUDA(c, c->d_in2.reserve(tot + 16));
        uint8_t *hb = c->h_in.as<uint8_t>(), *db = c->d_in2.as<uint8_t>();
        {
            struct Piece { const uint8_t *src; uint8_t *dst; size_t n; };
            static thread_local std::vector<Piece> pieces;
            pieces.clear();
            for (const Seg &sg : segs)
                for (size_t o = 0; o < sg.bytes; o += 262144) pieces.push_back({(const uint8_t *)sg.src + o, hb + sg.off + o, std::min<size_t>(262144, sg.bytes - o)});
            const int n_pieces = (int)pieces.size();
            const Piece *pc_ = pieces.data();
    #pragma omp parallel for schedule(dynamic, 1) num_threads(omp_team)
            for (int i = 0; i < n_pieces + n_prob; i++) {
                if (i < n_pieces) { memcpy(pc_[i].dst, pc_[i].src, pc_[i].n); continue; }
                const BaProb &P = probs[i - n_pieces];
                const BaBuilt &B = built[i - n_pieces];
                auto put = [hb](size_t seg, size_t at, const std::vector<int32_t> &src) { if (!src.empty()) memcpy(hb + seg + at * 4, src.data(), src.size() * 4); };
                put(o_act, P.act0, B.act_pose); put(o_lo, P.loff0, B.l_off); put(o_po, P.poff0, B.p_off); put(o_bi, P.blk0, B.blk_i);
                put(o_bj, P.blk0, B.blk_j); put(o_bc, P.bch0, B.blk_ch); put(o_p1, P.pair0, B.pr_e1); put(o_p2, P.pair0, B.pr_e2);
                put(o_cb, P.ch0, B.ch_blk); put(o_co, P.choff0, B.ch_off); put(o_lg, P.lgoff0, B.lg_off); put(o_gl, P.grp0, B.g_lm);
                put(o_gp, P.grp0, B.g_pose); put(o_pgo, P.pgoff0, B.pg_off); put(o_pgg, P.grp0, B.pg_groups);
            }
        }
        SVS_CUDA(c, cudaMemcpyAsync(db, hb, tot, cudaMemcpyHostToDevice, c->stream));
        A.probs = reinterpret_cast<BaProb *>(db + o_probs);
        A.poses = reinterpret_cast<double *>(db + o_pose); A.lms = reinterpret_cast<double *>(db + o_lm);
        A.act_pose = reinterpret_cast<int32_t *>(db + o_act);
        A.edge_p = reinterpret_cast<int32_t *>(db + o_ep); A.edge_l = reinterpret_cast<int32_t *>(db + o_el);
        A.edge_cam = db + o_ec; A.edge_uv = reinterpret_cast<double *>(db + o_uv);
        A.l_off = reinterpret_cast<int32_t *>(db + o_lo); A.l_edges = reinterpret_cast<int32_t *>(db + o_le);
        A.p_off = reinterpret_cast<int32_t *>(db + o_po); A.p_edges = reinterpret_cast<int32_t *>(db + o_pe);
        A.blk_i = reinterpret_cast<int32_t *>(db + o_bi); A.blk_j = reinterpret_cast<int32_t *>(db + o_bj);
        A.blk_ch = reinterpret_cast<int32_t *>(db + o_bc);
        A.pr_e1 = reinterpret_cast<int32_t *>(db + o_p1); A.pr_e2 = reinterpret_cast<int32_t *>(db + o_p2);
        A.ch_blk = reinterpret_cast<int32_t *>(db + o_cb); A.ch_off = reinterpret_cast<int32_t *>(db + o_co);
        A.lg_off = reinterpret_cast<int32_t *>(db + o_lg); A.g_lm = reinterpret_cast<int32_t *>(db + o_gl);
        A.g_pose = reinterpret_cast<int32_t *>(db + o_gp); A.pg_off = reinterpret_cast<int32_t *>(db + o_pgo);
        A.pg_groups = reinterpret_cast<int32_t *>(db + o_pgg);
    }
    // scratch
    size_t sc_b = (sumG * 36 + (size_t)sumL * 27 + (size_t)S_tot + sumCh * 36) * 8 + 64;
    SVS_CUDA(c, c->d_tmp.reserve(sc_b));
    size_t out_b = (size_t)sumE * 8 + (size_t)n_prob * sizeof(svs_ba_stats);
    SVS_CUDA(c, c->d_out.reserve(out_b + 16));
    SVS_CUDA(c, c->h_out.reserve(out_b + (size_t)sumN * 56 + (size_t)sumL * 24 + 64));
    double *scr = c->d_tmp.as<double>();
    A.Hpl = scr; A.WD = A.Hpl + sumG * 18;
    A.Hll = A.WD + sumG * 18; A.Dinv = A.Hll + (size_t)sumL * 9;
    A.bl = A.Dinv + (size_t)sumL * 9; A.xl = A.bl + (size_t)sumL * 3; A.lmT = A.xl + (size_t)sumL * 3;
    A.S_glob = A.lmT + (size_t)sumL * 3;
    A.part = A.S_glob + (size_t)S_tot;
    A.edge_chi2 = c->d_out.as<double>();
    A.stats = reinterpret_cast<svs_ba_stats *>(c->d_out.as<uint8_t>() + (size_t)sumE * 8);
    for (int i = 0; i < 4; i++) { A.K[0][i] = K_left[i]; A.K[1][i] = K_right[i]; }
    for (int i = 0; i < 7; i++) { A.ext[0][i] = ext_left[i]; A.ext[1][i] = ext_right[i]; }
    A.huber_delta = huber_delta; A.max_iter = max_iter; A.jac_mode = jacobian_mode; A.smem_bytes = (int)max_smem;
    // more windows than SMs: 256-thread CTAs, two per SM, so that the launch is one wave (unless the windows are so large that
    // two reduced systems do not fit an SM's shared memory)
    bool two_per_sm = n_prob > c->sm_count && 2 * (max_smem + 1024) <= 220 * 1024 && !getenv("SVS_BA_ONE_PER_SM");
    if (c->ba_threads == 256 && 2 * (max_smem + 1024) <= 220 * 1024) two_per_sm = true;
    if (c->ba_threads == 512) two_per_sm = false;
    if (two_per_sm) {
        SVS_CUDA(c, svs_i_opt_in_smem(c, reinterpret_cast<const void *>(k_ba_window<256>)));
        SVS_KERNEL(c, KID_BA_WINDOW, k_ba_window<256><<<n_prob, 256, max_smem, c->stream>>>(A));
    } else {
        SVS_CUDA(c, svs_i_opt_in_smem(c, reinterpret_cast<const void *>(k_ba_window<512>)));
        SVS_KERNEL(c, KID_BA_WINDOW, k_ba_window<512><<<n_prob, BA_T, max_smem, c->stream>>>(A));
    }
    uint8_t *ho = c->h_out.as<uint8_t>();
    SVS_CUDA(c, cudaMemcpyAsync(ho, c->d_out.p, out_b, cudaMemcpyDeviceToHost, c->stream));
    size_t ho_pose = align_up(out_b, 16), ho_lm = ho_pose + (size_t)sumN * 56;
    SVS_CUDA(c, cudaMemcpyAsync(ho + ho_pose, A.poses, (size_t)sumN * 56, cudaMemcpyDeviceToHost, c->stream));
    if (sumL) SVS_CUDA(c, cudaMemcpyAsync(ho + ho_lm, A.lms, (size_t)sumL * 24, cudaMemcpyDeviceToHost, c->stream));
    const double t_queued = now_s();
    SVS_CUDA(c, svs_i_wait(c));
    if (sumE) memcpy(edge_chi2_out, ho, (size_t)sumE * 8);
    if (stats) memcpy(stats, ho + (size_t)sumE * 8, (size_t)n_prob * sizeof(svs_ba_stats));
    memcpy(poses, ho + ho_pose, (size_t)sumN * 56);
    if (sumL) memcpy(lms, ho + ho_lm, (size_t)sumL * 24);
    const double t_end = now_s();
    c->ba_host_s[0] += t_built - t_begin; c->ba_host_s[1] += t_queued - t_built; c->ba_host_s[2] += t_end - t_queued;
    return SVS_OK;
}

// The window solver is latency-bound and long (milliseconds on a fraction of the SMs) while the per-frame kernels of the other
// context groups are short and fill the machine: with svs_ctx::stream_ba (a high-priority stream) its CTAs are placed as
// soon as an SM has room instead of queueing behind every CTA of an LK launch that was enqueued earlier.  The call is
// synchronous (it ends with svs_i_wait on the stream it used), so swapping the context stream for its duration is safe.
extern "C" int svs_ba_optimize(svs_ctx *c, int n_prob, const int32_t *kf_off, double *poses, const int32_t *lm_off, double *lms,
                               const int32_t *e_off, const int32_t *edge_kf, const int32_t *edge_lm, const uint8_t *edge_cam,
                               const double *edge_uv, const double K_left[4], const double K_right[4], const double ext_left[7],
                               const double ext_right[7], double huber_delta, int max_iter, int jacobian_mode,
                               double *edge_chi2_out, svs_ba_stats *stats)
{
    if (!c) return SVS_ERR_ARG;
    cudaStream_t s_main = c->stream;
    if (c->stream_ba && n_prob > 0) {
        SVS_CUDA(c, cudaSetDevice(c->device));
        SVS_CUDA(c, cudaEventRecord(c->ev_ba, s_main));
        SVS_CUDA(c, cudaStreamWaitEvent(c->stream_ba, c->ev_ba, 0));
        c->stream = c->stream_ba;
    }
    const int r = ba_optimize_on_stream(c, n_prob, kf_off, poses, lm_off, lms, e_off, edge_kf, edge_lm, edge_cam, edge_uv, K_left, K_right,
                                        ext_left, ext_right, huber_delta, max_iter, jacobian_mode, edge_chi2_out, stats);
    if (c->stream != s_main) {
        if (r != SVS_OK) cudaStreamSynchronize(c->stream);      // an error return may leave work queued: drain before switching back
        c->stream = s_main;
    }
    return r;
}

extern "C" int svs_ba_host_seconds(svs_ctx *c, double out[3])
{
    if (!c || !out) return SVS_ERR_ARG;
    for (int i = 0; i < 3; i++) out[i] = c->ba_host_s[i];
    return SVS_OK;
}
