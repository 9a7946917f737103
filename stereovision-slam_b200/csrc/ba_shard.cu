// Landmark-sharded bundle adjustment for windows too large for one CTA (BASELINE config 4: N = 50 keyframes,
// L = 1e5 landmarks, E ~ 5e5 edges) and for multi-GPU runs: every rank holds ALL N poses and a
// disjoint subset of the landmarks with all their edges.  Same problem and solver semantics as k_ba_window
// (reference src/backend.cpp:22-164, g2o LM + Schur + dense pivoted LDLT).
//
// k_bs_lm: ONE persistent cooperative kernel per GPU runs the whole optimizer.optimize(max_iter) (src/backend.cpp:163-164):
// the LM accept / reject control, lambda schedule and the stopping tests live on the device, there is no host round trip
// and no library collective.  Phases are separated by grid barriers; every sum has a fixed order (per-CTA partials, per-item
// partials, chunk partials added in index order), so the result is bitwise reproducible and identical on every rank.
//
// Exchange between ranks (fused with the solver, no NCCL): each rank owns an exchange WINDOW in its device memory
//     [64 arrival flags (u64)] [slot 0: xn doubles] [slot 1: xn doubles]
// that every peer maps (cudaIpc* across processes, plain pointers inside one process).  Per LM trial a rank writes its
// partial payload  [S upper-triangle 6x6 blocks | g | bp | chi2 | singular-V count]  — S_partial = Hpp_local - sum W V^-1 W^T,
// 364 KB at N = 50 — into its own slot, publishes a sequence number in every peer's flag row (st.release.sys), waits for the
// peers' numbers (ld.acquire.sys) and then ALL CTAs pull the peers' slots over NVLink and add them in rank order
// (ld.relaxed.sys, never cached) — an all-gather + local reduce: one NVLink crossing, identical bits everywhere.  A second,
// 2-double exchange carries [trial chi2 | landmark part of the scale term].  Slots alternate, the flags only grow.
// Every rank then solves the reduced system redundantly: pre-permuted (Eigen::LDLT pivot order = diagonal sorted by
// magnitude, see ba_ldlt.cuh) right-looking blocked LDL^T over the whole grid (coop_ldlt.cuh): 32 x 32 diagonal blocks by one
// warp, panel rows and 64 x 64 trailing tiles by every CTA, the right-hand side carried as an extra row so the forward
// substitution is free.
#include "svs_internal.h"
#include "geom_dev.cuh"
#include "coop_ldlt.cuh"
#include <cooperative_groups.h>
#include <algorithm>
#include <cstring>
#include <new>
#include <vector>

namespace cg = cooperative_groups;

#define CB_T 512        // threads per CTA, one CTA per SM
#define CB_CH 16        // (group, group) pairs per Schur chunk
#define CB_PW 32        // LDLT panel width
#define CB_ITEM 512     // edges / groups per pose-side work item
#define CB_MAXR 8       // ranks
#define CB_FLAGS 64     // u64 arrival flags at the head of a window

// in-kernel phase clock (gtid 0, %globaltimer): nanoseconds accumulated per phase into ctl[CT_TIME + k]
#define CB_TSTAMP(k) do { if (gtid == 0) { const unsigned long long now_ = cb_globaltimer(); ctl[CT_TIME + (k)] += (double)(now_ - t_prev); t_prev = now_; } } while (0)
enum { CT_LAMBDA = 0, CT_NI, CT_CUR, CT_RHO, CT_ACCEPT, CT_OK, CT_SCALE_POSE, CT_ABORT, CT_CHI_INIT, CT_ITERS, CT_TRIALS, CT_STOP,
       CT_LINS, CT_SIGN, CT_CUR_LOCAL, CT_HB, CT_TIME = 16, CT_COUNT = 32 };

struct CbDev {
    int N, L, E, G, nch, n_pitems, n_gitems, np, nbu, xn;
    double *poses, *poseT, *lms, *lmT;
    const int32_t *edge_p, *edge_l, *l_off, *l_edges, *lg_off, *g_lm, *g_pose;
    const uint8_t *edge_cam;
    const double *edge_uv;
    const int32_t *p_edges, *pitem_pose, *pitem_lo, *pitem_hi, *ppart_off;
    const int32_t *pg_groups, *gitem_pose, *gitem_lo, *gitem_hi, *gpart_off;
    const int32_t *pr_e1, *pr_e2, *ch_off, *ubk_ch, *ubk_i, *ubk_j;
    double *Hpl, *WD, *Hll, *Dinv, *bl, *xl, *part, *ppart, *gpart, *cta;
    double *Hpp, *bp, *Xsum, *X2, *A, *dvec, *xs, *xp, *ctl, *edge_chi2;
    unsigned long long *seq;          // [0] exchange sequence number (persists across launches)
    int n_ranks, rank;
    double *win[CB_MAXR];
    double K[2][4], ext[2][7];
    double huber_delta;
    int jac_mode;
    int bmax_local;                   // largest keyframe distance |i - j| over this shard's co-observing (pose, pose) blocks
    int band_cap;                     // doubles of dynamic shared memory behind `red` (set per launch)
};

struct svs_ba_shard {
    CbDev d;
    DevBuf buf, window;
    size_t window_bytes = 0;
    int grid_limit = 0;
};

// ------------------------------------------------------------------------------------------------ device helpers
__device__ __forceinline__ unsigned long long cb_globaltimer()
{
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    return t;
}
__device__ __forceinline__ int ubk_id(int i, int j, int N) { return i * N - (i * (i - 1)) / 2 + (j - i); }      // i <= j

__device__ __forceinline__ double cb_block_sum(double v, double *red)
{   // deterministic tree; result to all threads
    const int tid = threadIdx.x;
    red[tid] = v;
    __syncthreads();
    for (int s = CB_T / 2; s > 0; s >>= 1) { if (tid < s) red[tid] += red[tid + s]; __syncthreads(); }
    const double r = red[0];
    __syncthreads();
    return r;
}
__device__ __forceinline__ double cb_block_max(double v, double *red)
{
    const int tid = threadIdx.x;
    red[tid] = v;
    __syncthreads();
    for (int s = CB_T / 2; s > 0; s >>= 1) { if (tid < s) red[tid] = fmax(red[tid], red[tid + s]); __syncthreads(); }
    const double r = red[0];
    __syncthreads();
    return r;
}
// CTA 0: ordered sum / max of one column of the per-CTA partial table (8 doubles per CTA)
__device__ __forceinline__ double cb_cta_column(const double *cta, int col, int n_cta, bool is_max, double *red)
{
    double a = 0;
    for (int i = threadIdx.x; i < n_cta; i += CB_T) { const double v = cta[8 * i + col]; a = is_max ? fmax(a, v) : a + v; }
    return is_max ? cb_block_max(a, red) : cb_block_sum(a, red);
}

__device__ __forceinline__ void st_release_sys(unsigned long long *p, unsigned long long v)
{
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long *p)
{
    unsigned long long v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ double ld_relaxed_sys(const double *p)
{
    double v;
    asm volatile("ld.relaxed.sys.global.f64 %0, [%1];" : "=d"(v) : "l"(p) : "memory");
    return v;
}

// All ranks: own slot (seq & 1) holds `count` doubles (written by any CTA BEFORE the grid barrier the caller has just
// passed).  Result: dst[i] = sum over ranks (max for i >= max_idx, when max_idx >= 0) in rank order.  Returns false after a timeout.
__device__ bool cb_exchange(const CbDev &D, cg::grid_group &grid, int count, int max_idx, double *dst)
{
    const unsigned long long seq = D.seq[0] + 1;
    const int slot = (int)(seq & 1);
    if (D.n_ranks > 1) {
        if (blockIdx.x == 0 && threadIdx.x < D.n_ranks) {
            const int p = threadIdx.x;
            __threadfence_system();
            unsigned long long *theirs = reinterpret_cast<unsigned long long *>(D.win[p]) + D.rank;
            st_release_sys(theirs, seq);                                   // "rank's data number seq is in its window"
            const unsigned long long *mine = reinterpret_cast<const unsigned long long *>(D.win[D.rank]) + p;
            const long long t0 = clock64();
            while (ld_acquire_sys(mine) < seq) {
                if (clock64() - t0 > 20000000000LL) { D.ctl[CT_ABORT] = 1.0; break; }   // ~10 s: a peer never arrived
                __nanosleep(200);
            }
        }
        grid.sync();
        const size_t base = CB_FLAGS + (size_t)slot * D.xn;
        for (int i = blockIdx.x * CB_T + threadIdx.x; i < count; i += gridDim.x * CB_T) {
            double acc = ld_relaxed_sys(D.win[0] + base + i);
            for (int p = 1; p < D.n_ranks; p++) {
                const double v = ld_relaxed_sys(D.win[p] + base + i);
                acc = (max_idx >= 0 && i >= max_idx) ? fmax(acc, v) : acc + v;
            }
            dst[i] = acc;
        }
    } else {
        const double *src = D.win[0] + CB_FLAGS + (size_t)slot * D.xn;
        for (int i = blockIdx.x * CB_T + threadIdx.x; i < count; i += gridDim.x * CB_T) dst[i] = src[i];
    }
    grid.sync();
    if (blockIdx.x == 0 && threadIdx.x == 0) D.seq[0] = seq;
    grid.sync();
    return D.ctl[CT_ABORT] == 0.0;
}

// ------------------------------------------------------------------------------------------------ the optimizer
__global__ void __launch_bounds__(CB_T, 1) k_bs_lm(CbDev D, int max_iter)
{
    cg::grid_group grid = cg::this_grid();
    extern __shared__ double smd[];
    double *red = smd;                         // CB_T
    double *sm_ldlt = smd + CB_T;              // CL_SMEM_DOUBLES (coop_ldlt.cuh)
    double *sdiag = sm_ldlt + (D.np + 1 <= CB_T ? CL_SMALL_SMEM_DOUBLES(D.np) : (size_t)CL_SMEM_DOUBLES); // np + 2
    int *sperm = reinterpret_cast<int *>(sdiag + D.np + 2);     // np
    __shared__ double s_w[CB_T / 32][28];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int gtid = blockIdx.x * CB_T + tid, gsz = gridDim.x * CB_T;
    const int N = D.N, L = D.L, np = D.np, n_cta = gridDim.x;
    const double hd = D.huber_delta;
    double *ctl = D.ctl;
    unsigned ldlt_gen = 0;
    unsigned long long t_prev = cb_globaltimer();

    for (int i = gtid; i < 7 * N; i += gsz) D.poseT[i] = D.poses[i];
    for (int i = gtid; i < 3 * L; i += gsz) D.lmT[i] = D.lms[i];
    if (gtid == 0) {
        ctl[CT_LAMBDA] = 0; ctl[CT_NI] = 2; ctl[CT_CUR] = 0; ctl[CT_ABORT] = 0; ctl[CT_ITERS] = 0; ctl[CT_TRIALS] = 0; ctl[CT_STOP] = 0;
        ctl[CT_LINS] = 0; ctl[CT_CHI_INIT] = 0;
        for (int i = CT_TIME; i < CT_COUNT; i++) ctl[i] = 0;
        D.seq[2] = 0;          // sub-grid barrier counter of the LDLT (coop_ldlt.cuh)
    }
    grid.sync();

    for (int it = 0; it < max_iter; it++) {
        // ================= linearise at the accepted state =================
        {   // landmark side: Hll, bl, W per GROUP = (landmark, keyframe); robust chi2; max |diag Hll|
            double acc = 0, md = 0;
            for (int l = gtid; l < L; l += gsz) {
                const int s0 = D.l_off[l], s1 = D.l_off[l + 1];
                if (s0 == s1) continue;
                double H[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0}, b3[3] = {0, 0, 0};
                const double *pl = D.lms + 3 * (size_t)l;
                int gi = D.lg_off[l] - 1, cur_p = -1;
                for (int s = s0; s < s1; s++) {
                    const int e = D.l_edges[s], cam = D.edge_cam[e], ep = D.edge_p[e];
                    const bool first = ep != cur_p;
                    if (first) { gi++; cur_p = ep; }
                    const double *T = D.poses + 7 * ep;
                    double er[2], a[3], c[3], Jp[12], Jl[6];
                    gd::ba_error(T, D.ext[cam], D.K[cam], pl, D.edge_uv + 2 * (size_t)e, er, a, c);
                    if (D.jac_mode == 1) gd::ba_jac_numeric(T, D.ext[cam], D.K[cam], pl, D.edge_uv + 2 * (size_t)e, Jp, Jl);
                    else gd::ba_jac_analytic(T, D.ext[cam], D.K[cam], a, c, Jp, Jl);
                    double e2 = er[0] * er[0] + er[1] * er[1], r0, r1;
                    gd::huber(e2, hd, r0, r1);
                    acc += r0;
#pragma unroll
                    for (int x = 0; x < 3; x++) {
                        b3[x] -= r1 * (Jl[x] * er[0] + Jl[3 + x] * er[1]);
#pragma unroll
                        for (int y = 0; y < 3; y++) H[x * 3 + y] += r1 * (Jl[x] * Jl[y] + Jl[3 + x] * Jl[3 + y]);
                    }
                    double2 *W = reinterpret_cast<double2 *>(D.Hpl + 18 * (size_t)gi);
#pragma unroll
                    for (int x = 0; x < 9; x++) {
                        const int u = 2 * x, v = 2 * x + 1;
                        const double w0 = r1 * (Jp[u / 3] * Jl[u % 3] + Jp[6 + u / 3] * Jl[3 + u % 3]);
                        const double w1 = r1 * (Jp[v / 3] * Jl[v % 3] + Jp[6 + v / 3] * Jl[3 + v % 3]);
                        if (first) W[x] = make_double2(w0, w1);
                        else { const double2 t = W[x]; W[x] = make_double2(t.x + w0, t.y + w1); }
                    }
                }
#pragma unroll
                for (int x = 0; x < 9; x++) D.Hll[9 * (size_t)l + x] = H[x];
#pragma unroll
                for (int x = 0; x < 3; x++) D.bl[3 * (size_t)l + x] = b3[x];
                md = fmax(md, fmax(fabs(H[0]), fmax(fabs(H[4]), fabs(H[8]))));
            }
            const double tot = cb_block_sum(acc, red), mx = cb_block_max(md, red);
            if (tid == 0) { D.cta[8 * blockIdx.x + 0] = tot; D.cta[8 * blockIdx.x + 1] = mx; }
        }
        // pose side: one work item = up to CB_ITEM edges of one keyframe -> 21 + 6 partial sums
        for (int item = blockIdx.x; item < D.n_pitems; item += gridDim.x) {
            const int a = D.pitem_pose[item];
            const double *T = D.poses + 7 * a;
            double v27[27];
#pragma unroll
            for (int x = 0; x < 27; x++) v27[x] = 0;
            for (int s = D.pitem_lo[item] + tid; s < D.pitem_hi[item]; s += CB_T) {
                const int e = D.p_edges[s], cam = D.edge_cam[e];
                const double *p3 = D.lms + 3 * (size_t)D.edge_l[e];
                double er[2], aa[3], c[3], Jp[12], Jl[6];
                gd::ba_error(T, D.ext[cam], D.K[cam], p3, D.edge_uv + 2 * (size_t)e, er, aa, c);
                if (D.jac_mode == 1) gd::ba_jac_numeric(T, D.ext[cam], D.K[cam], p3, D.edge_uv + 2 * (size_t)e, Jp, Jl);
                else gd::ba_jac_analytic(T, D.ext[cam], D.K[cam], aa, c, Jp, Jl);
                double e2 = er[0] * er[0] + er[1] * er[1], r0, r1;
                gd::huber(e2, hd, r0, r1);
                int k = 0;
#pragma unroll
                for (int x = 0; x < 6; x++) {
                    v27[21 + x] -= r1 * (Jp[x] * er[0] + Jp[6 + x] * er[1]);
#pragma unroll
                    for (int y = x; y < 6; y++) v27[k++] += r1 * (Jp[x] * Jp[y] + Jp[6 + x] * Jp[6 + y]);
                }
            }
#pragma unroll
            for (int x = 0; x < 27; x++) { const double w = gd::warp_sum(v27[x]); if (lane == 0) s_w[warp][x] = w; }
            __syncthreads();
            if (tid < 27) {
                double s = 0;
                for (int w = 0; w < CB_T / 32; w++) s += s_w[w][tid];
                D.ppart[27 * (size_t)item + tid] = s;
            }
            __syncthreads();
        }
        grid.sync();
        CB_TSTAMP(0);
        if (blockIdx.x == 0) {      // ordered sums of the item partials -> Hpp, bp; chi2; iteration-0 payload
            for (int t = tid; t < 27 * N; t += CB_T) {
                const int a = t / 27, v = t - 27 * a;
                double s = 0;
                for (int i2 = D.ppart_off[a]; i2 < D.ppart_off[a + 1]; i2++) s += D.ppart[27 * (size_t)i2 + v];
                if (v >= 21) D.bp[6 * a + v - 21] = s;
                else {
                    int x = 0, k = v;
                    while (k >= 6 - x) { k -= 6 - x; x++; }
                    const int y = x + k;
                    D.Hpp[36 * a + x * 6 + y] = s; D.Hpp[36 * a + y * 6 + x] = s;
                }
            }
            const double cur = cb_cta_column(D.cta, 0, n_cta, false, red);
            const double mh = cb_cta_column(D.cta, 1, n_cta, true, red);
            __syncthreads();
            double *slot = D.win[D.rank] + CB_FLAGS + (size_t)((D.seq[0] + 1) & 1) * D.xn;
            if (it == 0) {
                for (int i = tid; i < np; i += CB_T) slot[i] = D.Hpp[36 * (i / 6) + 7 * (i % 6)];
                if (tid == 0) { slot[np] = mh; slot[np + 1] = (double)D.bmax_local; slot[np + 2] = -(double)D.band_cap; }
            }
            if (tid == 0) { ctl[CT_CUR_LOCAL] = cur; ctl[CT_LINS] += 1; }     // local chi2; the global one comes with the payload
        }
        grid.sync();
        CB_TSTAMP(1);
        if (it == 0) {      // lambda_init = tau * max diagonal over ALL vertices (global Hpp diagonal, every rank's Hll)
            if (!cb_exchange(D, grid, np + 3, np, D.Xsum)) return;
            if (gtid == 0) {
                double md = D.Xsum[np];
                for (int i = 0; i < np; i++) md = fmax(md, fabs(D.Xsum[i]));
                ctl[CT_LAMBDA] = 1e-5 * md; ctl[CT_NI] = 2;
                // scalar half bandwidth of S over ALL ranks' blocks; -1 (dense path) unless the band fits EVERY rank's
                // shared memory, so that all ranks factor the same way and stay bit-identical replicas of each other
                const int hb_all = min(np - 1, 6 * ((int)D.Xsum[np + 1] + 1) - 1);
                const bool fits = hb_all >= 1 && hb_all <= 255 && CL_BAND_DOUBLES(np, hb_all) <= (size_t)(-D.Xsum[np + 2]);
                ctl[CT_HB] = fits ? (double)hb_all : -1.0;
            }
            grid.sync();
        }
        // ================= trial loop =================
        int q = 0;
        double rho = 0;
        do {
            const double lambda = ctl[CT_LAMBDA];
            {   // V^-1 per landmark and W V^-1 per group
                double bad = 0;
                for (int l = gtid; l < L; l += gsz) {
                    if (D.l_off[l] == D.l_off[l + 1]) continue;
                    double M[9], Di[9];
#pragma unroll
                    for (int x = 0; x < 9; x++) M[x] = D.Hll[9 * (size_t)l + x];
                    M[0] += lambda; M[4] += lambda; M[8] += lambda;
                    if (!gd::inv3(M, Di)) bad += 1;
#pragma unroll
                    for (int x = 0; x < 9; x++) D.Dinv[9 * (size_t)l + x] = Di[x];
                    for (int g = D.lg_off[l]; g < D.lg_off[l + 1]; g++) {
                        const double *W = D.Hpl + 18 * (size_t)g;
                        double2 *O = reinterpret_cast<double2 *>(D.WD + 18 * (size_t)g);
                        double X[18];
#pragma unroll
                        for (int x = 0; x < 6; x++)
#pragma unroll
                            for (int y = 0; y < 3; y++) X[x * 3 + y] = W[x * 3] * Di[y] + W[x * 3 + 1] * Di[3 + y] + W[x * 3 + 2] * Di[6 + y];
#pragma unroll
                        for (int x = 0; x < 9; x++) O[x] = make_double2(X[2 * x], X[2 * x + 1]);
                    }
                }
                const double tb = cb_block_sum(bad, red);
                if (tid == 0) D.cta[8 * blockIdx.x + 2] = tb;
            }
            grid.sync();
            CB_TSTAMP(2);
            // chunk partials: 4 threads per chunk of <= CB_CH (group, group) pairs of one 6x6 block, one 3x3 quadrant each
            for (int t = gtid; t < 4 * D.nch; t += gsz) {
                const int ch = t >> 2, qr = (t >> 1) & 1, qc = t & 1;
                double a00 = 0, a01 = 0, a02 = 0, a10 = 0, a11 = 0, a12 = 0, a20 = 0, a21 = 0, a22 = 0;
                for (int p = D.ch_off[ch]; p < D.ch_off[ch + 1]; p++) {
                    const double *X = D.WD + 18 * (size_t)D.pr_e1[p] + 9 * qr, *Y = D.Hpl + 18 * (size_t)D.pr_e2[p] + 9 * qc;
                    const double x0 = X[0], x1 = X[1], x2 = X[2], x3 = X[3], x4 = X[4], x5 = X[5], x6 = X[6], x7 = X[7], x8 = X[8];
                    const double y0 = Y[0], y1 = Y[1], y2 = Y[2], y3 = Y[3], y4 = Y[4], y5 = Y[5], y6 = Y[6], y7 = Y[7], y8 = Y[8];
                    a00 += x0 * y0 + x1 * y1 + x2 * y2; a01 += x0 * y3 + x1 * y4 + x2 * y5; a02 += x0 * y6 + x1 * y7 + x2 * y8;
                    a10 += x3 * y0 + x4 * y1 + x5 * y2; a11 += x3 * y3 + x4 * y4 + x5 * y5; a12 += x3 * y6 + x4 * y7 + x5 * y8;
                    a20 += x6 * y0 + x7 * y1 + x8 * y2; a21 += x6 * y3 + x7 * y4 + x8 * y5; a22 += x6 * y6 + x7 * y7 + x8 * y8;
                }
                double *O = D.part + 36 * (size_t)ch + 18 * qr + 3 * qc;
                O[0] = a00; O[1] = a01; O[2] = a02; O[6] = a10; O[7] = a11; O[8] = a12; O[12] = a20; O[13] = a21; O[14] = a22;
            }
            // g items: sum over the groups of one keyframe of (W V^-1) bl
            for (int item = blockIdx.x; item < D.n_gitems; item += gridDim.x) {
                double s6[6] = {0, 0, 0, 0, 0, 0};
                for (int s = D.gitem_lo[item] + tid; s < D.gitem_hi[item]; s += CB_T) {
                    const int g = D.pg_groups[s];
                    const double *O = D.WD + 18 * (size_t)g, *b3 = D.bl + 3 * (size_t)D.g_lm[g];
#pragma unroll
                    for (int x = 0; x < 6; x++) s6[x] += O[x * 3] * b3[0] + O[x * 3 + 1] * b3[1] + O[x * 3 + 2] * b3[2];
                }
#pragma unroll
                for (int x = 0; x < 6; x++) { const double w = gd::warp_sum(s6[x]); if (lane == 0) s_w[warp][x] = w; }
                __syncthreads();
                if (tid < 6) {
                    double s = 0;
                    for (int w = 0; w < CB_T / 32; w++) s += s_w[w][tid];
                    D.gpart[6 * (size_t)item + tid] = s;
                }
                __syncthreads();
            }
            grid.sync();
            CB_TSTAMP(3);
            {   // payload into the own slot: [S_partial upper blocks | g_partial | bp_local | chi2_local | bad]
                double *slot = D.win[D.rank] + CB_FLAGS + (size_t)((D.seq[0] + 1) & 1) * D.xn;
                for (int t = gtid; t < D.nbu * 36; t += gsz) {
                    const int u = t / 36, ent = t - 36 * u;
                    // a near-diagonal block has hundreds of chunks: four independent partial sums keep four loads in flight (the
                    // order of the additions stays fixed, so the result is still reproducible bit for bit)
                    double s0 = 0, s1 = 0, s2 = 0, s3 = 0;
                    const int c_lo = D.ubk_ch[u], c_hi = D.ubk_ch[u + 1];
                    int c2 = c_lo;
                    for (; c2 + 4 <= c_hi; c2 += 4) {
                        s0 += D.part[36 * (size_t)c2 + ent]; s1 += D.part[36 * (size_t)(c2 + 1) + ent];
                        s2 += D.part[36 * (size_t)(c2 + 2) + ent]; s3 += D.part[36 * (size_t)(c2 + 3) + ent];
                    }
                    for (; c2 < c_hi; c2++) s0 += D.part[36 * (size_t)c2 + ent];
                    const double sum = (s0 + s1) + (s2 + s3);
                    const int i = D.ubk_i[u];
                    slot[t] = ((i == D.ubk_j[u]) ? D.Hpp[36 * i + ent] : 0.0) - sum;
                }
                const size_t og = (size_t)D.nbu * 36;
                for (int t = gtid; t < np; t += gsz) {
                    const int a = t / 6, x = t - 6 * a;
                    double s = 0;
                    for (int i2 = D.gpart_off[a]; i2 < D.gpart_off[a + 1]; i2++) s += D.gpart[6 * (size_t)i2 + x];
                    slot[og + t] = D.bp[t] - s;
                    slot[og + np + t] = D.bp[t];
                }
                if (blockIdx.x == 0) {
                    const double tb = cb_cta_column(D.cta, 2, n_cta, false, red);
                    if (tid == 0) { slot[og + 2 * np] = ctl[CT_CUR_LOCAL]; slot[og + 2 * np + 1] = tb; }
                }
            }
            grid.sync();
            CB_TSTAMP(4);
            if (!cb_exchange(D, grid, D.xn, -1, D.Xsum)) return;
            CB_TSTAMP(5);
            // ================= reduced system: S = Xsum_S + lambda I, solve S x = g =================
            const double *XS = D.Xsum, *Xg = D.Xsum + (size_t)D.nbu * 36, *Xbp = Xg + np;
            const double chi_cur = Xg[2 * np];
            const bool v_ok = Xg[2 * np + 1] == 0.0;
            const int pitch = np + 1;
            // A window whose landmarks are seen by keyframes at most b apart has a block-banded S: when the band fits
            // shared memory one CTA factors it in natural order (no fill outside the band, no grid barrier); otherwise the
            // dense system goes through the grid-wide blocked LDL^T in Eigen's pivot order.  Same inertia test, same
            // solution up to rounding (deviation from Eigen's elimination order for banded windows).
            const int hb = (int)ctl[CT_HB];
            const bool band = hb > 0;
            if (band) {
                if (blockIdx.x == 0) {
                    double *Bb = sm_ldlt, *zb = Bb + (size_t)np * (hb + 1);
                    const int w = hb + 1;
                    for (int t = tid; t < np * w; t += CB_T) {
                        const int r = t / w, d = t - r * w, c2 = r - d;
                        double v = 0.0;
                        if (c2 >= 0) {
                            const int i = c2 / 6, j = r / 6;
                            v = XS[(size_t)ubk_id(i, j, N) * 36 + (c2 - 6 * i) * 6 + (r - 6 * j)];
                            if (d == 0) v += lambda;
                        }
                        Bb[t] = v;
                    }
                    for (int i = tid; i < np; i += CB_T) zb[i] = Xg[i];
                    CB_TSTAMP(6);
                    band_ldlt_solve_cta<CB_T>(Bb, zb, np, hb, D.xs, ctl + CT_SIGN);
                }
            } else {
                // pivot order: diagonal sorted by decreasing magnitude (ties: lower index first), computed by every CTA
                for (int i = tid; i < np; i += CB_T) {
                    const int a = i / 6, r = i - 6 * a;
                    sdiag[i] = fabs(XS[(size_t)ubk_id(a, a, N) * 36 + r * 7] + lambda);
                }
                __syncthreads();
                for (int i = tid; i < np; i += CB_T) {
                    const double di = sdiag[i];
                    int r = 0;
                    for (int j = 0; j < np; j++) { const double dj = sdiag[j]; r += (dj > di || (dj == di && j < i)) ? 1 : 0; }
                    sperm[r] = i;
                }
                __syncthreads();
                // permuted lower triangle + the right-hand side as row np
                for (int t = gtid; t < (np + 1) * np; t += gsz) {
                    const int a = t / np, b = t - a * np;
                    if (a == np) { D.A[(size_t)a * pitch + b] = Xg[sperm[b]]; continue; }
                    if (b > a) continue;
                    const int r = sperm[a], c2 = sperm[b], i = r / 6, j = c2 / 6;
                    double v = (i <= j) ? XS[(size_t)ubk_id(i, j, N) * 36 + (r - 6 * i) * 6 + (c2 - 6 * j)]
                                        : XS[(size_t)ubk_id(j, i, N) * 36 + (c2 - 6 * j) * 6 + (r - 6 * i)];
                    if (a == b) v += lambda;
                    D.A[(size_t)a * pitch + b] = v;
                }
                if (gtid == 0) ctl[CT_SIGN] = 0;
                grid.sync();
                CB_TSTAMP(6);
                // right-looking blocked LDL^T over the whole grid + back-substitution (coop_ldlt.cuh); solution in D.xs (permuted)
                if (np + 1 <= CB_T) coop_ldlt_solve_small<CB_T>(grid, D.A, np, pitch, D.dvec, D.xs, ctl + CT_SIGN, sm_ldlt, reinterpret_cast<unsigned *>(D.seq + 2), ldlt_gen);
                else coop_ldlt_solve<CB_T>(grid, D.A, np, pitch, D.dvec, D.xs, ctl + CT_SIGN, sm_ldlt, reinterpret_cast<unsigned *>(D.seq + 2), ldlt_gen);
            }
            CB_TSTAMP(7);
            if (blockIdx.x == 0) {
                const int sign = (int)ctl[CT_SIGN];
                const bool ok = v_ok && (sign == 1 || sign == 0);
                const double *xs = D.xs;
                for (int i = tid; i < np; i += CB_T) D.xp[band ? i : sperm[i]] = ok ? xs[i] : 0.0;
                __syncthreads();
                for (int a = tid; a < N; a += CB_T) gd::se3_oplus(D.poses + 7 * a, D.xp + 6 * a, D.poseT + 7 * a);
                if (tid == 0) {
                    double scp = 0;
                    for (int i = 0; i < np; i++) scp += D.xp[i] * (lambda * D.xp[i] + Xbp[i]);
                    ctl[CT_SCALE_POSE] = scp; ctl[CT_OK] = ok ? 1.0 : 0.0;
                }
            }
            grid.sync();
            CB_TSTAMP(8);
            {   // back-substitution, trial landmarks, landmark part of the scale term, trial chi2
                const bool ok = ctl[CT_OK] != 0.0;
                double sc = 0, acc = 0;
                for (int l = gtid; l < L; l += gsz) {
                    const int s0 = D.l_off[l], s1 = D.l_off[l + 1];
                    if (s0 == s1) continue;
                    double c3[3] = {D.bl[3 * (size_t)l], D.bl[3 * (size_t)l + 1], D.bl[3 * (size_t)l + 2]}, x3[3] = {0, 0, 0};
                    if (ok) {
                        for (int g = D.lg_off[l]; g < D.lg_off[l + 1]; g++) {
                            const double *W = D.Hpl + 18 * (size_t)g, *xx = D.xp + 6 * D.g_pose[g];
#pragma unroll
                            for (int y = 0; y < 3; y++) {
                                double sm2 = 0;
#pragma unroll
                                for (int x = 0; x < 6; x++) sm2 += W[x * 3 + y] * xx[x];
                                c3[y] -= sm2;
                            }
                        }
                        const double *Di = D.Dinv + 9 * (size_t)l;
#pragma unroll
                        for (int x = 0; x < 3; x++) x3[x] = Di[x * 3] * c3[0] + Di[x * 3 + 1] * c3[1] + Di[x * 3 + 2] * c3[2];
                    }
                    double pt[3];
#pragma unroll
                    for (int x = 0; x < 3; x++) {
                        D.xl[3 * (size_t)l + x] = x3[x];
                        pt[x] = D.lms[3 * (size_t)l + x] + x3[x];
                        D.lmT[3 * (size_t)l + x] = pt[x];
                        sc += x3[x] * (lambda * x3[x] + D.bl[3 * (size_t)l + x]);
                    }
                    for (int s = s0; s < s1; s++) {
                        const int e = D.l_edges[s], cam = D.edge_cam[e];
                        double er[2], a[3], c[3];
                        gd::ba_error(D.poseT + 7 * D.edge_p[e], D.ext[cam], D.K[cam], pt, D.edge_uv + 2 * (size_t)e, er, a, c);
                        double e2 = er[0] * er[0] + er[1] * er[1], r0, r1;
                        gd::huber(e2, hd, r0, r1);
                        acc += r0;
                    }
                }
                const double t1 = cb_block_sum(sc, red), t2 = cb_block_sum(acc, red);
                if (tid == 0) { D.cta[8 * blockIdx.x + 3] = t1; D.cta[8 * blockIdx.x + 4] = t2; }
            }
            grid.sync();
            CB_TSTAMP(9);
            if (blockIdx.x == 0) {
                const double scl = cb_cta_column(D.cta, 3, n_cta, false, red), chn = cb_cta_column(D.cta, 4, n_cta, false, red);
                double *slot = D.win[D.rank] + CB_FLAGS + (size_t)((D.seq[0] + 1) & 1) * D.xn;
                if (tid == 0) { slot[0] = chn; slot[1] = scl; }
            }
            grid.sync();
            CB_TSTAMP(10);
            if (!cb_exchange(D, grid, 2, -1, D.X2)) return;
            CB_TSTAMP(11);
            if (gtid == 0) {
                const bool ok = ctl[CT_OK] != 0.0;
                const double tc = ok ? D.X2[0] : DBL_MAX;
                const double scale = ctl[CT_SCALE_POSE] + D.X2[1] + 1e-3;
                const double r = (chi_cur - tc) / scale;
                gd::LmCtl lm = {ctl[CT_LAMBDA], ctl[CT_NI]};
                const int acc2 = gd::lm_accept(lm, r, tc) ? 1 : 0;
                ctl[CT_LAMBDA] = lm.lambda; ctl[CT_NI] = lm.ni; ctl[CT_RHO] = r; ctl[CT_ACCEPT] = acc2;
                if (it == 0 && q == 0) ctl[CT_CHI_INIT] = chi_cur;
                ctl[CT_CUR] = acc2 ? tc : chi_cur;
                ctl[CT_TRIALS] += 1;
            }
            grid.sync();
            CB_TSTAMP(12);
            rho = ctl[CT_RHO];
            if (ctl[CT_ACCEPT] != 0.0) {
                for (int i = gtid; i < 7 * N; i += gsz) D.poses[i] = D.poseT[i];
                for (int i = gtid; i < 3 * L; i += gsz) D.lms[i] = D.lmT[i];
            }
            grid.sync();
            CB_TSTAMP(13);
            q++;
        } while (rho < 0 && q < 10);
        if (gtid == 0) ctl[CT_ITERS] += 1;
        if (q == 10 || rho == 0) break;
    }
    // per-edge chi2 as g2o leaves it: errors of the LAST evaluated state (possibly a rejected trial)
    for (int e = gtid; e < D.E; e += gsz) {
        const int cam = D.edge_cam[e];
        double er[2], a[3], c[3];
        gd::ba_error(D.poseT + 7 * D.edge_p[e], D.ext[cam], D.K[cam], D.lmT + 3 * (size_t)D.edge_l[e], D.edge_uv + 2 * (size_t)e, er, a, c);
        D.edge_chi2[e] = er[0] * er[0] + er[1] * er[1];
    }
}

// ------------------------------------------------------------------------------------------------------------
extern "C" {

svs_ba_shard *svs_ba_shard_create(svs_ctx *c, int n_kf, const double *poses, int n_lm, const double *lms, int n_edge,
                                  const int32_t *edge_kf, const int32_t *edge_lm, const uint8_t *edge_cam, const double *edge_uv,
                                  const double K_left[4], const double K_right[4], const double ext_left[7], const double ext_right[7],
                                  double huber_delta, int jacobian_mode)
{
    if (!c || n_kf <= 0 || n_lm < 0 || n_edge < 0 || !poses) return nullptr;
    if (cudaSetDevice(c->device) != cudaSuccess) return nullptr;
    const int N = n_kf, L = n_lm, E = n_edge, np = 6 * N;
    for (int e = 0; e < E; e++) if (edge_kf[e] < 0 || edge_kf[e] >= N || edge_lm[e] < 0 || edge_lm[e] >= L) { c->err = "ba_shard: edge index out of range"; return nullptr; }
    // ---- structure (every keyframe is in the system: with landmark sharding a pose may have no LOCAL edge)
    std::vector<int32_t> l_off(L + 1, 0), l_edges(E), p_off(N + 1, 0), p_edges(E);
    for (int e = 0; e < E; e++) { l_off[edge_lm[e] + 1]++; p_off[edge_kf[e] + 1]++; }
    for (int l = 0; l < L; l++) l_off[l + 1] += l_off[l];
    for (int a = 0; a < N; a++) p_off[a + 1] += p_off[a];
    { std::vector<int> fl(l_off.begin(), l_off.end() - 1), fp(p_off.begin(), p_off.end() - 1);
      for (int e = 0; e < E; e++) { l_edges[fl[edge_lm[e]]++] = e; p_edges[fp[edge_kf[e]]++] = e; } }
    for (int l = 0; l < L; l++) {      // pose-ascending inside a landmark (stable), so that a GROUP = (landmark, keyframe) is a run
        int32_t *a = l_edges.data() + l_off[l];
        const int m = l_off[l + 1] - l_off[l];
        for (int i = 1; i < m; i++) {
            const int32_t v = a[i], pv = edge_kf[v];
            int j = i - 1;
            while (j >= 0 && edge_kf[a[j]] > pv) { a[j + 1] = a[j]; j--; }
            a[j + 1] = v;
        }
    }
    std::vector<int32_t> lg_off, g_lm, g_pose, pg_off(N + 1, 0), pg_groups;
    lg_off.reserve(L + 1); g_lm.reserve(E); g_pose.reserve(E);
    int G = 0;
    for (int l = 0; l < L; l++) {
        lg_off.push_back(G);
        int cur = -1;
        for (int s1 = l_off[l]; s1 < l_off[l + 1]; s1++) {
            const int pp = edge_kf[l_edges[s1]];
            if (pp != cur) { cur = pp; g_lm.push_back(l); g_pose.push_back(pp); pg_off[pp + 1]++; G++; }
        }
    }
    lg_off.push_back(G);
    for (int a = 0; a < N; a++) pg_off[a + 1] += pg_off[a];
    pg_groups.resize(G);
    { std::vector<int> fill(pg_off.begin(), pg_off.end() - 1);
      for (int g = 0; g < G; g++) pg_groups[fill[g_pose[g]]++] = g; }
    // work items of the pose-side sums (<= CB_ITEM edges / groups of one keyframe each)
    std::vector<int32_t> pitem_pose, pitem_lo, pitem_hi, ppart_off(N + 1, 0), gitem_pose, gitem_lo, gitem_hi, gpart_off(N + 1, 0);
    for (int a = 0; a < N; a++) {
        ppart_off[a] = (int32_t)pitem_pose.size();
        for (int s = p_off[a]; s < p_off[a + 1]; s += CB_ITEM) { pitem_pose.push_back(a); pitem_lo.push_back(s); pitem_hi.push_back(std::min(p_off[a + 1], s + CB_ITEM)); }
        gpart_off[a] = (int32_t)gitem_pose.size();
        for (int s = pg_off[a]; s < pg_off[a + 1]; s += CB_ITEM) { gitem_pose.push_back(a); gitem_lo.push_back(s); gitem_hi.push_back(std::min(pg_off[a + 1], s + CB_ITEM)); }
    }
    ppart_off[N] = (int32_t)pitem_pose.size(); gpart_off[N] = (int32_t)gitem_pose.size();
    // (group, group) pairs per DENSE upper block id(i, j), i <= j, in landmark order; chunks of <= CB_CH pairs
    const int nbu = N * (N + 1) / 2;
    auto bid = [N](int i, int j) { return i * N - (i * (i - 1)) / 2 + (j - i); };
    std::vector<long long> bcount(nbu + 1, 0);
    for (int l = 0; l < L; l++)
        for (int g1 = lg_off[l]; g1 < lg_off[l + 1]; g1++)
            for (int g2 = g1; g2 < lg_off[l + 1]; g2++) bcount[bid(g_pose[g1], g_pose[g2]) + 1]++;
    std::vector<long long> bstart(nbu, 0);
    std::vector<int32_t> ubk_ch(nbu + 1, 0), ubk_i(nbu), ubk_j(nbu), ch_off;
    long long run = 0;
    for (int i = 0; i < N; i++)
        for (int j = i; j < N; j++) {
            const int u = bid(i, j);
            ubk_i[u] = i; ubk_j[u] = j;
        }
    for (int u = 0; u < nbu; u++) {
        ubk_ch[u] = (int32_t)ch_off.size();
        bstart[u] = run;
        for (long long c0 = 0; c0 < bcount[u + 1]; c0 += CB_CH) ch_off.push_back((int32_t)(run + c0));
        run += bcount[u + 1];
    }
    ubk_ch[nbu] = (int32_t)ch_off.size();
    int bmax_local = 0;
    for (int u = 0; u < nbu; u++) if (bcount[u + 1] > 0) bmax_local = std::max(bmax_local, ubk_j[u] - ubk_i[u]);
    const int nch = (int)ch_off.size();
    ch_off.push_back((int32_t)run);
    if (run > 0x7fffffffLL / 2) { c->err = "ba_shard: too many edge pairs for one shard"; return nullptr; }
    std::vector<int32_t> pr_e1((size_t)run), pr_e2((size_t)run);
    for (int l = 0; l < L; l++)
        for (int g1 = lg_off[l]; g1 < lg_off[l + 1]; g1++)
            for (int g2 = g1; g2 < lg_off[l + 1]; g2++) {
                const long long q = bstart[bid(g_pose[g1], g_pose[g2])]++;
                pr_e1[(size_t)q] = g1; pr_e2[(size_t)q] = g2;
            }
    // chunk boundaries must not span two blocks: they do not (chunks restart at every block start)
    svs_ba_shard *sh = new (std::nothrow) svs_ba_shard();
    if (!sh) return nullptr;
    const int xn = nbu * 36 + 2 * np + 8;
    struct Seg { const void *src; size_t bytes, off; };
    std::vector<Seg> segs;
    size_t tot = 0;
    auto add = [&](const void *p, size_t bytes) { size_t o = tot; segs.push_back({p, bytes, o}); tot = align_up(tot + bytes, 256); return o; };
    const size_t o_pose = add(poses, (size_t)N * 56), o_poseT = add(poses, (size_t)N * 56), o_lm = add(lms, (size_t)L * 24), o_lmT = add(lms, (size_t)L * 24);
    const size_t o_ep = add(edge_kf, (size_t)E * 4), o_el = add(edge_lm, (size_t)E * 4), o_ec = add(edge_cam, (size_t)E), o_uv = add(edge_uv, (size_t)E * 16);
    const size_t o_lo = add(l_off.data(), l_off.size() * 4), o_le = add(l_edges.data(), (size_t)E * 4), o_pe = add(p_edges.data(), (size_t)E * 4);
    const size_t o_lg = add(lg_off.data(), lg_off.size() * 4), o_gl = add(g_lm.data(), (size_t)G * 4), o_gp = add(g_pose.data(), (size_t)G * 4);
    const size_t o_pgg = add(pg_groups.data(), (size_t)G * 4);
    const size_t o_pip = add(pitem_pose.data(), pitem_pose.size() * 4), o_pil = add(pitem_lo.data(), pitem_lo.size() * 4), o_pih = add(pitem_hi.data(), pitem_hi.size() * 4);
    const size_t o_ppo = add(ppart_off.data(), ppart_off.size() * 4);
    const size_t o_gip = add(gitem_pose.data(), gitem_pose.size() * 4), o_gil = add(gitem_lo.data(), gitem_lo.size() * 4), o_gih = add(gitem_hi.data(), gitem_hi.size() * 4);
    const size_t o_gpo = add(gpart_off.data(), gpart_off.size() * 4);
    const size_t o_p1 = add(pr_e1.data(), pr_e1.size() * 4), o_p2 = add(pr_e2.data(), pr_e2.size() * 4), o_co = add(ch_off.data(), ch_off.size() * 4);
    const size_t o_uc = add(ubk_ch.data(), ubk_ch.size() * 4), o_ui = add(ubk_i.data(), ubk_i.size() * 4), o_uj = add(ubk_j.data(), ubk_j.size() * 4);
    const size_t scratch0 = tot;
    auto res = [&](size_t bytes) { size_t o = tot; tot = align_up(tot + bytes, 256); return o; };
    const int max_cta = 2048;
    const size_t o_Hpl = res((size_t)G * 144), o_WD = res((size_t)G * 144), o_Hll = res((size_t)L * 72), o_Di = res((size_t)L * 72);
    const size_t o_bl = res((size_t)L * 24), o_xl = res((size_t)L * 24), o_part = res((size_t)nch * 288), o_pp = res(pitem_pose.size() * 216 + 8);
    const size_t o_gpt = res(gitem_pose.size() * 48 + 8), o_cta = res((size_t)max_cta * 64), o_Hpp = res((size_t)N * 288), o_bp = res((size_t)np * 8);
    const size_t o_X = res((size_t)xn * 8), o_X2 = res(64), o_A = res((size_t)(np + 1) * (np + 1) * 8), o_dv = res((size_t)np * 8), o_xs = res((size_t)np * 8), o_xp = res((size_t)np * 8);
    const size_t o_ctl = res(CT_COUNT * 8), o_chi = res((size_t)E * 8 + 8), o_seq = res(64);
    sh->window_bytes = CB_FLAGS * 8 + 2 * (size_t)xn * 8;
    if (sh->buf.reserve(tot) != cudaSuccess || sh->window.reserve(sh->window_bytes) != cudaSuccess) { c->err = "ba_shard: cudaMalloc failed"; sh->buf.release(); sh->window.release(); delete sh; return nullptr; }
    uint8_t *db = sh->buf.as<uint8_t>();
    for (const Seg &s : segs) if (s.bytes) cudaMemcpyAsync(db + s.off, s.src, s.bytes, cudaMemcpyHostToDevice, c->stream);
    cudaMemsetAsync(db + scratch0, 0, tot - scratch0, c->stream);
    cudaMemsetAsync(sh->window.p, 0, sh->window_bytes, c->stream);
    CbDev &d = sh->d;
    d.N = N; d.L = L; d.E = E; d.G = G; d.nch = nch; d.n_pitems = (int)pitem_pose.size(); d.n_gitems = (int)gitem_pose.size(); d.np = np; d.nbu = nbu; d.xn = xn; d.bmax_local = bmax_local; d.band_cap = 0;
    d.poses = (double *)(db + o_pose); d.poseT = (double *)(db + o_poseT); d.lms = (double *)(db + o_lm); d.lmT = (double *)(db + o_lmT);
    d.edge_p = (int32_t *)(db + o_ep); d.edge_l = (int32_t *)(db + o_el); d.edge_cam = db + o_ec; d.edge_uv = (double *)(db + o_uv);
    d.l_off = (int32_t *)(db + o_lo); d.l_edges = (int32_t *)(db + o_le); d.p_edges = (int32_t *)(db + o_pe);
    d.lg_off = (int32_t *)(db + o_lg); d.g_lm = (int32_t *)(db + o_gl); d.g_pose = (int32_t *)(db + o_gp); d.pg_groups = (int32_t *)(db + o_pgg);
    d.pitem_pose = (int32_t *)(db + o_pip); d.pitem_lo = (int32_t *)(db + o_pil); d.pitem_hi = (int32_t *)(db + o_pih); d.ppart_off = (int32_t *)(db + o_ppo);
    d.gitem_pose = (int32_t *)(db + o_gip); d.gitem_lo = (int32_t *)(db + o_gil); d.gitem_hi = (int32_t *)(db + o_gih); d.gpart_off = (int32_t *)(db + o_gpo);
    d.pr_e1 = (int32_t *)(db + o_p1); d.pr_e2 = (int32_t *)(db + o_p2); d.ch_off = (int32_t *)(db + o_co);
    d.ubk_ch = (int32_t *)(db + o_uc); d.ubk_i = (int32_t *)(db + o_ui); d.ubk_j = (int32_t *)(db + o_uj);
    d.Hpl = (double *)(db + o_Hpl); d.WD = (double *)(db + o_WD); d.Hll = (double *)(db + o_Hll); d.Dinv = (double *)(db + o_Di);
    d.bl = (double *)(db + o_bl); d.xl = (double *)(db + o_xl); d.part = (double *)(db + o_part); d.ppart = (double *)(db + o_pp);
    d.gpart = (double *)(db + o_gpt); d.cta = (double *)(db + o_cta); d.Hpp = (double *)(db + o_Hpp); d.bp = (double *)(db + o_bp);
    d.Xsum = (double *)(db + o_X); d.X2 = (double *)(db + o_X2); d.A = (double *)(db + o_A); d.dvec = (double *)(db + o_dv); d.xs = (double *)(db + o_xs); d.xp = (double *)(db + o_xp);
    d.ctl = (double *)(db + o_ctl); d.edge_chi2 = (double *)(db + o_chi); d.seq = (unsigned long long *)(db + o_seq);
    d.n_ranks = 1; d.rank = 0;
    for (int r = 0; r < CB_MAXR; r++) d.win[r] = nullptr;
    d.win[0] = sh->window.as<double>();
    for (int i = 0; i < 4; i++) { d.K[0][i] = K_left[i]; d.K[1][i] = K_right[i]; }
    for (int i = 0; i < 7; i++) { d.ext[0][i] = ext_left[i]; d.ext[1][i] = ext_right[i]; }
    d.huber_delta = huber_delta; d.jac_mode = jacobian_mode;
    if (cudaStreamSynchronize(c->stream) != cudaSuccess) { c->err = "ba_shard: upload failed"; sh->buf.release(); sh->window.release(); delete sh; return nullptr; }
    return sh;
}

void svs_ba_shard_destroy(svs_ctx *c, svs_ba_shard *sh)
{
    if (!sh) return;
    if (c) { cudaSetDevice(c->device); cudaStreamSynchronize(c->stream); }
    sh->buf.release(); sh->window.release();
    delete sh;
}

int svs_ba_shard_window(const svs_ba_shard *sh, void **window_dev, size_t *bytes)
{
    if (!sh || !window_dev || !bytes) return SVS_ERR_ARG;
    *window_dev = sh->window.p; *bytes = sh->window_bytes;
    return SVS_OK;
}

int svs_ba_shard_set_peers(svs_ctx *c, svs_ba_shard *sh, int n_ranks, int rank, void *const *peer_window_dev)
{
    if (!c || !sh || !peer_window_dev || n_ranks < 1 || n_ranks > CB_MAXR || rank < 0 || rank >= n_ranks) return SVS_ERR_ARG;
    if (peer_window_dev[rank] != sh->window.p) SVS_FAIL(c, SVS_ERR_ARG, "ba_shard_set_peers: peer_window_dev[rank] must be this shard's own window");
    for (int r = 0; r < n_ranks; r++) if (!peer_window_dev[r]) SVS_FAIL(c, SVS_ERR_ARG, "ba_shard_set_peers: null window");
    sh->d.n_ranks = n_ranks; sh->d.rank = rank;
    for (int r = 0; r < CB_MAXR; r++) sh->d.win[r] = r < n_ranks ? reinterpret_cast<double *>(peer_window_dev[r]) : nullptr;
    return SVS_OK;
}

int svs_ba_shard_set_grid_limit(svs_ba_shard *sh, int max_ctas)
{
    if (!sh) return SVS_ERR_ARG;
    sh->grid_limit = max_ctas;
    return SVS_OK;
}

int svs_ba_shard_launch(svs_ctx *c, svs_ba_shard *sh, int max_iter)
{
    if (!c || !sh || max_iter < 0) return SVS_ERR_ARG;
    SVS_CUDA(c, cudaSetDevice(c->device));
    const int np = sh->d.np;
    size_t smem = ((size_t)CB_T + (np + 1 <= CB_T ? CL_SMALL_SMEM_DOUBLES(np) : (size_t)CL_SMEM_DOUBLES) + np + 2) * 8 + (size_t)np * 4 + 16;
    SVS_CUDA(c, svs_i_opt_in_smem(c, reinterpret_cast<const void *>(k_bs_lm)));
    {   // room for the banded factorisation at THIS shard's bandwidth (landmark-sharded windows see the same bandwidth on
        // every rank; the kernel takes the band path only when the global band fits every rank, see CT_HB)
        cudaFuncAttributes fa;
        SVS_CUDA(c, cudaFuncGetAttributes(&fa, reinterpret_cast<const void *>(k_bs_lm)));
        const int hb = std::min(np - 1, 6 * (sh->d.bmax_local + 1) - 1);
        const size_t want = ((size_t)CB_T + CL_BAND_DOUBLES(np, hb)) * 8 + 16;
        if (want <= (size_t)fa.maxDynamicSharedSizeBytes) smem = std::max(smem, want);
    }
    int per_sm = 0;
    SVS_CUDA(c, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_bs_lm, CB_T, smem));
    if (per_sm < 1) SVS_FAIL(c, SVS_ERR_CAPACITY, "ba_shard: the cooperative solver does not fit an SM");
    int grid = c->sm_count;                       // one CTA per SM: every phase is a grid-stride loop
    if (sh->grid_limit > 0) grid = std::min(grid, sh->grid_limit);
    grid = std::max(1, std::min(grid, 2048));
    CbDev d = sh->d;
    d.band_cap = getenv("SVS_BS_DENSE") ? 0 : (int)(smem / 8) - CB_T - 2;
    void *args[] = {&d, &max_iter};
    svs_i_prof_begin(c, KID_BA_WINDOW);
    cudaError_t e = cudaLaunchCooperativeKernel(reinterpret_cast<const void *>(k_bs_lm), dim3(grid), dim3(CB_T), args, smem, c->stream);
    svs_i_prof_end(c);
    SVS_CUDA(c, e);
    c->launches++;
    return SVS_OK;
}

int svs_ba_shard_finish(svs_ctx *c, svs_ba_shard *sh, svs_ba_stats *stats)
{
    if (!c || !sh) return SVS_ERR_ARG;
    SVS_CUDA(c, cudaSetDevice(c->device));
    double ctl[CT_COUNT];
    SVS_CUDA(c, cudaMemcpyAsync(ctl, sh->d.ctl, sizeof(ctl), cudaMemcpyDeviceToHost, c->stream));
    SVS_CUDA(c, cudaStreamSynchronize(c->stream));
    if (ctl[CT_ABORT] != 0.0) SVS_FAIL(c, SVS_ERR_CUDA, "ba_shard: a peer rank never arrived at an exchange (timeout)");
    if (stats) {
        stats->iterations = (int)ctl[CT_ITERS]; stats->trials = (int)ctl[CT_TRIALS]; stats->linearizations = (int)ctl[CT_LINS];
        stats->solves = (int)ctl[CT_TRIALS]; stats->lambda = ctl[CT_LAMBDA]; stats->chi2 = ctl[CT_CUR]; stats->chi2_init = ctl[CT_CHI_INIT];
    }
    return SVS_OK;
}

int svs_ba_shard_optimize(svs_ctx *c, svs_ba_shard *sh, int max_iter, svs_ba_stats *stats)
{
    SVS_TRY(svs_ba_shard_launch(c, sh, max_iter));
    return svs_ba_shard_finish(c, sh, stats);
}

/* Nanoseconds the last svs_ba_shard_optimize spent per phase, measured inside the kernel with %globaltimer (diagnostics):
 * 0 linearise | 1 Hpp/bp reduce | 2 V^-1, W V^-1 | 3 Schur chunks + g | 4 payload | 5 exchange | 6 permute | 7 LDLT + back-subst |
 * 8 trial poses | 9 landmark back-subst + chi2 | 10 reduce | 11 exchange (2 doubles) | 12 accept test | 13 commit */
int svs_ba_shard_phase_ns(svs_ctx *c, svs_ba_shard *sh, double out[16])
{
    if (!c || !sh || !out) return SVS_ERR_ARG;
    SVS_CUDA(c, cudaSetDevice(c->device));
    SVS_CUDA(c, cudaMemcpyAsync(out, sh->d.ctl + CT_TIME, 16 * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    SVS_CUDA(c, cudaStreamSynchronize(c->stream));
    return SVS_OK;
}

int svs_ba_shard_get(svs_ctx *c, svs_ba_shard *sh, double *poses_out, double *lms_out, double *edge_chi2_out)
{
    if (!c || !sh) return SVS_ERR_ARG;
    SVS_CUDA(c, cudaSetDevice(c->device));
    CbDev &d = sh->d;
    if (poses_out) SVS_CUDA(c, cudaMemcpyAsync(poses_out, d.poses, (size_t)d.N * 56, cudaMemcpyDeviceToHost, c->stream));
    if (lms_out && d.L) SVS_CUDA(c, cudaMemcpyAsync(lms_out, d.lms, (size_t)d.L * 24, cudaMemcpyDeviceToHost, c->stream));
    if (edge_chi2_out && d.E) SVS_CUDA(c, cudaMemcpyAsync(edge_chi2_out, d.edge_chi2, (size_t)d.E * 8, cudaMemcpyDeviceToHost, c->stream));
    SVS_CUDA(c, cudaStreamSynchronize(c->stream));
    return SVS_OK;
}

// ---- CUDA IPC helpers so that a caller with any bootstrap (MPI, sockets, torch.distributed) can map the peers' windows
int svs_ipc_export(svs_ctx *c, const void *dev_ptr, unsigned char handle_out[64])
{
    if (!c || !dev_ptr || !handle_out) return SVS_ERR_ARG;
    SVS_CUDA(c, cudaSetDevice(c->device));
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "cudaIpcMemHandle_t is 64 bytes");
    cudaIpcMemHandle_t h;
    SVS_CUDA(c, cudaIpcGetMemHandle(&h, const_cast<void *>(dev_ptr)));
    memcpy(handle_out, &h, 64);
    return SVS_OK;
}
int svs_ipc_import(svs_ctx *c, const unsigned char handle[64], void **dev_ptr_out)
{
    if (!c || !handle || !dev_ptr_out) return SVS_ERR_ARG;
    SVS_CUDA(c, cudaSetDevice(c->device));
    cudaIpcMemHandle_t h;
    memcpy(&h, handle, 64);
    SVS_CUDA(c, cudaIpcOpenMemHandle(dev_ptr_out, h, cudaIpcMemLazyEnablePeerAccess));
    return SVS_OK;
}
int svs_ipc_release(svs_ctx *c, void *dev_ptr)
{
    if (!c || !dev_ptr) return SVS_ERR_ARG;
    SVS_CUDA(c, cudaSetDevice(c->device));
    SVS_CUDA(c, cudaIpcCloseMemHandle(dev_ptr));
    return SVS_OK;
}

}  // extern "C"
