// Landmark-sharded bundle adjustment for windows too large for one CTA (BASELINE config 4: N = 50 keyframes,
// L = 1e5 landmarks, E ~ 5e5 edges) and for multi-GPU runs: every rank holds ALL N poses and a
// disjoint subset of the landmarks with all their edges.  Same problem and solver semantics as k_ba_window
// (reference src/backend.cpp:22-164, g2o LM + Schur + dense pivoted LDLT); the LM control loop
// lives in the caller because two quantities must be summed over ranks (NCCL all-reduce, or nothing when there is
// one shard):
//     after  svs_ba_shard_linearize : lin  = [Hpp (36N) | bp (6N) | chi2 | pad]      SUM   (+ max diagonal: MAX)
//     after  svs_ba_shard_schur     : red  = [-sum W V^-1 W^T (6N x 6N) | -sum W V^-1 bl (6N)]   SUM
//     after  svs_ba_shard_try       : tri  = [trial chi2 | landmark part of the scale term]       SUM
// Every rank then holds identical S, g and solves the reduced system redundantly (deterministic), back-substitutes
// its own landmarks, and takes the same accept / reject decision.
// Grid-wide kernels, no atomics: pose-side sums are owned by one CTA per keyframe, Schur blocks by 36 threads each.
#include "svs_internal.h"
#include "geom_dev.cuh"
#include <algorithm>
#include <cstring>
#include <new>
#include <vector>

#define BS_T 256
#include "ba_ldlt.cuh"

struct BsDev {
    int N, L, E, nblk, npairs;
    double *poses, *poseT, *lms, *lmT;
    const int32_t *edge_p, *edge_l, *l_off, *l_edges, *p_off, *p_edges, *blk_i, *blk_j, *blk_off, *ep_off, *ep_pos;
    const uint8_t *edge_cam;
    const double *edge_uv;
    double *Hpl, *WD, *Hll, *Dinv, *bl, *xl, *contrib, *xp, *partial, *Sfull, *gfull, *tmp;
    int *tr;
    double K[2][4], ext[2][7];
    double huber_delta;
    int jac_mode;
};

struct svs_ba_shard {
    BsDev d;
    DevBuf buf;
    int n_partial = 0;
};

__device__ __forceinline__ double bs_block_sum(double v, double *red)
{
    int tid = threadIdx.x;
    red[tid] = v;
    __syncthreads();
    for (int s = BS_T / 2; s > 0; s >>= 1) { if (tid < s) red[tid] += red[tid + s]; __syncthreads(); }
    double r = red[0];
    __syncthreads();
    return r;
}

// ---- linearise: thread per landmark (Hll, bl, Hpl, robust chi2 partial per CTA)
__global__ void __launch_bounds__(BS_T) k_bs_linearize_lm(BsDev D)
{
    __shared__ double red[BS_T];
    int l = blockIdx.x * BS_T + threadIdx.x;
    double acc = 0;
    if (l < D.L && D.l_off[l] != D.l_off[l + 1]) {
        double H[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0}, b3[3] = {0, 0, 0};
        const double *pl = D.lms + 3 * (size_t)l;
        for (int s = D.l_off[l]; s < D.l_off[l + 1]; s++) {
            int e = D.l_edges[s], cam = D.edge_cam[e];
            const double *T = D.poses + 7 * D.edge_p[e];
            double er[2], a[3], c[3], Jp[12], Jl[6];
            gd::ba_error(T, D.ext[cam], D.K[cam], pl, D.edge_uv + 2 * (size_t)e, er, a, c);
            if (D.jac_mode == 1) gd::ba_jac_numeric(T, D.ext[cam], D.K[cam], pl, D.edge_uv + 2 * (size_t)e, Jp, Jl);
            else gd::ba_jac_analytic(T, D.ext[cam], D.K[cam], a, c, Jp, Jl);
            double e2 = er[0] * er[0] + er[1] * er[1], r0, r1;
            gd::huber(e2, D.huber_delta, r0, r1);
            acc += r0;
#pragma unroll
            for (int x = 0; x < 3; x++) {
                b3[x] -= r1 * (Jl[x] * er[0] + Jl[3 + x] * er[1]);
#pragma unroll
                for (int y = 0; y < 3; y++) H[x * 3 + y] += r1 * (Jl[x] * Jl[y] + Jl[3 + x] * Jl[3 + y]);
            }
            double *W = D.Hpl + 18 * (size_t)e;
#pragma unroll
            for (int x = 0; x < 6; x++)
#pragma unroll
                for (int y = 0; y < 3; y++) W[x * 3 + y] = r1 * (Jp[x] * Jl[y] + Jp[6 + x] * Jl[3 + y]);
        }
#pragma unroll
        for (int x = 0; x < 9; x++) D.Hll[9 * (size_t)l + x] = H[x];
#pragma unroll
        for (int x = 0; x < 3; x++) D.bl[3 * (size_t)l + x] = b3[x];
    }
    double tot = bs_block_sum(acc, red);
    if (threadIdx.x == 0) D.partial[blockIdx.x] = tot;
}

// ---- one CTA per keyframe: Hpp block (36), bp (6) into lin; also the max |diagonal| candidates
__global__ void __launch_bounds__(BS_T) k_bs_linearize_pose(BsDev D, double *lin)
{
    __shared__ double red[BS_T];
    int a = blockIdx.x, tid = threadIdx.x;
    double H[21], b6[6];
#pragma unroll
    for (int x = 0; x < 21; x++) H[x] = 0;
#pragma unroll
    for (int x = 0; x < 6; x++) b6[x] = 0;
    const double *T = D.poses + 7 * a;
    for (int s = D.p_off[a] + tid; s < D.p_off[a + 1]; s += BS_T) {
        int e = D.p_edges[s], cam = D.edge_cam[e];
        const double *pl = D.lms + 3 * (size_t)D.edge_l[e];
        double er[2], aa[3], c[3], Jp[12], Jl[6];
        gd::ba_error(T, D.ext[cam], D.K[cam], pl, D.edge_uv + 2 * (size_t)e, er, aa, c);
        if (D.jac_mode == 1) gd::ba_jac_numeric(T, D.ext[cam], D.K[cam], pl, D.edge_uv + 2 * (size_t)e, Jp, Jl);
        else gd::ba_jac_analytic(T, D.ext[cam], D.K[cam], aa, c, Jp, Jl);
        double e2 = er[0] * er[0] + er[1] * er[1], r0, r1;
        gd::huber(e2, D.huber_delta, r0, r1);
        int k = 0;
#pragma unroll
        for (int x = 0; x < 6; x++) {
            b6[x] -= r1 * (Jp[x] * er[0] + Jp[6 + x] * er[1]);
#pragma unroll
            for (int y = x; y < 6; y++) H[k++] += r1 * (Jp[x] * Jp[y] + Jp[6 + x] * Jp[6 + y]);
        }
    }
    int k = 0;
    for (int x = 0; x < 6; x++) {
        double v = bs_block_sum(b6[x], red);
        if (tid == 0) lin[36 * (size_t)D.N + 6 * a + x] = v;
        for (int y = x; y < 6; y++) {
            double h = bs_block_sum(H[k++], red);
            if (tid == 0) { lin[36 * (size_t)a + x * 6 + y] = h; lin[36 * (size_t)a + y * 6 + x] = h; }
        }
    }
}

// ---- deterministic sum of per-CTA partials into out[0]; optional landmark max diagonal into out[1]
__global__ void __launch_bounds__(BS_T) k_bs_sum_partials(const double *partial, int n, double *out)
{
    __shared__ double red[BS_T];
    double acc = 0;
    for (int i = threadIdx.x; i < n; i += BS_T) acc += partial[i];
    double t = bs_block_sum(acc, red);
    if (threadIdx.x == 0) out[0] = t;
}
__global__ void __launch_bounds__(BS_T) k_bs_maxdiag(BsDev D, const double *lin, double *out)
{
    __shared__ double red[BS_T];
    double md = 0;
    for (int i = threadIdx.x; i < 6 * D.N; i += BS_T) md = fmax(md, fabs(lin[36 * (size_t)(i / 6) + 7 * (i % 6)]));
    for (int l = threadIdx.x; l < D.L; l += BS_T)
        if (D.l_off[l] != D.l_off[l + 1])
            md = fmax(md, fmax(fabs(D.Hll[9 * (size_t)l]), fmax(fabs(D.Hll[9 * (size_t)l + 4]), fabs(D.Hll[9 * (size_t)l + 8]))));
    red[threadIdx.x] = md;
    __syncthreads();
    for (int s = BS_T / 2; s > 0; s >>= 1) { if (threadIdx.x < s) red[threadIdx.x] = fmax(red[threadIdx.x], red[threadIdx.x + s]); __syncthreads(); }
    if (threadIdx.x == 0) out[0] = red[0];
}

// ---- Schur: V^-1 per landmark; per edge W V^-1 and the pair records; block sums; g
__global__ void __launch_bounds__(BS_T) k_bs_dinv(BsDev D, double lambda, int *flag)
{
    int l = blockIdx.x * BS_T + threadIdx.x;
    if (l >= D.L || D.l_off[l] == D.l_off[l + 1]) return;
    double M[9], Di[9];
#pragma unroll
    for (int x = 0; x < 9; x++) M[x] = D.Hll[9 * (size_t)l + x];
    M[0] += lambda; M[4] += lambda; M[8] += lambda;
    if (!gd::inv3(M, Di)) *flag = 0;
#pragma unroll
    for (int x = 0; x < 9; x++) D.Dinv[9 * (size_t)l + x] = Di[x];
}
__global__ void __launch_bounds__(BS_T) k_bs_edge_records(BsDev D)
{
    int e = blockIdx.x * BS_T + threadIdx.x;
    if (e >= D.E) return;
    int l = D.edge_l[e], p1 = D.edge_p[e];
    const double *Di = D.Dinv + 9 * (size_t)l, *W = D.Hpl + 18 * (size_t)e;
    double X[18];
#pragma unroll
    for (int x = 0; x < 6; x++)
#pragma unroll
        for (int y = 0; y < 3; y++) X[x * 3 + y] = W[x * 3] * Di[y] + W[x * 3 + 1] * Di[3 + y] + W[x * 3 + 2] * Di[6 + y];
    double *O = D.WD + 18 * (size_t)e;
#pragma unroll
    for (int x = 0; x < 18; x++) O[x] = X[x];
    int q = D.ep_off[e];
    for (int s = D.l_off[l]; s < D.l_off[l + 1]; s++) {
        int e2 = D.l_edges[s];
        if (p1 > D.edge_p[e2]) continue;
        const double *Y = D.Hpl + 18 * (size_t)e2;
        double *C = D.contrib + 36 * (size_t)D.ep_pos[q++];
#pragma unroll
        for (int r = 0; r < 6; r++)
#pragma unroll
            for (int c2 = 0; c2 < 6; c2++) C[r * 6 + c2] = X[r * 3] * Y[c2 * 3] + X[r * 3 + 1] * Y[c2 * 3 + 1] + X[r * 3 + 2] * Y[c2 * 3 + 2];
    }
}
// red (np x np, row-major, pitch np) = - sum of records per block (both triangles), red_g below
__global__ void __launch_bounds__(288) k_bs_block_sum(BsDev D, double *red)
{
    int t = blockIdx.x * 288 + threadIdx.x;
    if (t >= D.nblk * 36) return;
    int bk = t / 36, ent = t - 36 * bk, r = ent / 6, c2 = ent - 6 * r, np = 6 * D.N;
    double sum = 0;
    const double *C = D.contrib + ent;
    for (int s = D.blk_off[bk]; s < D.blk_off[bk + 1]; s++) sum += C[36 * (size_t)s];
    int i = D.blk_i[bk], j = D.blk_j[bk];
    red[(size_t)(6 * i + r) * np + 6 * j + c2] = -sum;
    if (i != j) red[(size_t)(6 * j + c2) * np + 6 * i + r] = -sum;
}
__global__ void __launch_bounds__(BS_T) k_bs_g(BsDev D, double *red_g)
{
    __shared__ double red[BS_T];
    int a = blockIdx.x, tid = threadIdx.x;
    double s6[6] = {0, 0, 0, 0, 0, 0};
    for (int s = D.p_off[a] + tid; s < D.p_off[a + 1]; s += BS_T) {
        int e = D.p_edges[s];
        const double *O = D.WD + 18 * (size_t)e, *b3 = D.bl + 3 * (size_t)D.edge_l[e];
#pragma unroll
        for (int x = 0; x < 6; x++) s6[x] += O[x * 3] * b3[0] + O[x * 3 + 1] * b3[1] + O[x * 3 + 2] * b3[2];
    }
    for (int x = 0; x < 6; x++) { double v = bs_block_sum(s6[x], red); if (tid == 0) red_g[6 * a + x] = -v; }
}

// ---- solve (one CTA): S = Hpp + lambda I + red, g = bp + red_g ; pivoted LDLT ; xp ; trial poses ; pose scale term
__global__ void __launch_bounds__(BS_T) k_bs_solve(BsDev D, const double *lin, const double *red, double lambda, int flag_ok, double *tri)
{
    __shared__ int s_piv;
    int tid = threadIdx.x, N = D.N, np = 6 * N, pitch = np | 1;
    for (int i = tid; i < np * np; i += BS_T) {
        int r = i / np, c = i - r * np;
        double v = red[i];
        if (r / 6 == c / 6) v += lin[36 * (size_t)(r / 6) + (r % 6) * 6 + (c % 6)] + (r == c ? lambda : 0.0);
        D.Sfull[(size_t)r * pitch + c] = v;
    }
    for (int i = tid; i < np; i += BS_T) D.gfull[i] = lin[36 * (size_t)N + i] + red[(size_t)np * np + i];
    __syncthreads();
    bool ok = flag_ok != 0;
    if (ok) ok = block_ldlt_solve(D.Sfull, pitch, np, D.gfull, D.xp, D.tr, D.tmp, &s_piv);
    if (!ok) { for (int i = tid; i < np; i += BS_T) D.xp[i] = 0.0; }
    __syncthreads();
    for (int a = tid; a < N; a += BS_T) gd::se3_oplus(D.poses + 7 * a, D.xp + 6 * a, D.poseT + 7 * a);
    if (tid == 0) {
        double scp = 0;
        for (int i = 0; i < np; i++) scp += D.xp[i] * (lambda * D.xp[i] + lin[36 * (size_t)N + i]);
        tri[2] = scp;
        tri[3] = ok ? 1.0 : 0.0;
    }
}
// ---- back-substitution + trial landmarks + landmark scale partial (thread per landmark)
__global__ void __launch_bounds__(BS_T) k_bs_backsub(BsDev D, double lambda, const double *tri)
{
    __shared__ double red[BS_T];
    int l = blockIdx.x * BS_T + threadIdx.x;
    bool ok = tri[3] != 0.0;
    double sc = 0;
    if (l < D.L && D.l_off[l] != D.l_off[l + 1]) {
        double c3[3] = {D.bl[3 * (size_t)l], D.bl[3 * (size_t)l + 1], D.bl[3 * (size_t)l + 2]}, x3[3] = {0, 0, 0};
        if (ok) {
            for (int s = D.l_off[l]; s < D.l_off[l + 1]; s++) {
                int e = D.l_edges[s];
                const double *W = D.Hpl + 18 * (size_t)e, *xx = D.xp + 6 * D.edge_p[e];
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
#pragma unroll
        for (int x = 0; x < 3; x++) {
            D.xl[3 * (size_t)l + x] = x3[x];
            D.lmT[3 * (size_t)l + x] = D.lms[3 * (size_t)l + x] + x3[x];
            sc += x3[x] * (lambda * x3[x] + D.bl[3 * (size_t)l + x]);
        }
    }
    double t = bs_block_sum(sc, red);
    if (threadIdx.x == 0) D.partial[blockIdx.x] = t;
}
// ---- robust chi2 of the trial state (thread per landmark, per-CTA partial)
__global__ void __launch_bounds__(BS_T) k_bs_chi2(BsDev D, int trial_state)
{
    __shared__ double red[BS_T];
    int l = blockIdx.x * BS_T + threadIdx.x;
    const double *pz = trial_state ? D.poseT : D.poses, *lz = trial_state ? D.lmT : D.lms;
    double acc = 0;
    if (l < D.L) {
        for (int s = D.l_off[l]; s < D.l_off[l + 1]; s++) {
            int e = D.l_edges[s], cam = D.edge_cam[e];
            double er[2], a[3], c[3];
            gd::ba_error(pz + 7 * D.edge_p[e], D.ext[cam], D.K[cam], lz + 3 * (size_t)l, D.edge_uv + 2 * (size_t)e, er, a, c);
            double e2 = er[0] * er[0] + er[1] * er[1], r0, r1;
            gd::huber(e2, D.huber_delta, r0, r1);
            acc += r0;
        }
    }
    double t = bs_block_sum(acc, red);
    if (threadIdx.x == 0) D.partial[blockIdx.x] = t;
}
__global__ void k_bs_accept(BsDev D)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < 7 * D.N) D.poses[i] = D.poseT[i];
    if (i < 3 * D.L) D.lms[i] = D.lmT[i];
}
__global__ void k_bs_edge_chi2(BsDev D, double *out)
{
    int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= D.E) return;
    int cam = D.edge_cam[e];
    double er[2], a[3], c[3];
    gd::ba_error(D.poseT + 7 * D.edge_p[e], D.ext[cam], D.K[cam], D.lmT + 3 * (size_t)D.edge_l[e], D.edge_uv + 2 * (size_t)e, er, a, c);
    out[e] = er[0] * er[0] + er[1] * er[1];
}

// ------------------------------------------------------------------------------------------------------------
extern "C" {

svs_ba_shard *svs_ba_shard_create(svs_ctx *c, int n_kf, const double *poses, int n_lm, const double *lms, int n_edge,
                                  const int32_t *edge_kf, const int32_t *edge_lm, const uint8_t *edge_cam, const double *edge_uv,
                                  const double K_left[4], const double K_right[4], const double ext_left[7], const double ext_right[7],
                                  double huber_delta, int jacobian_mode)
{
    if (!c || n_kf <= 0 || n_lm < 0 || n_edge < 0 || !poses) return nullptr;
    cudaSetDevice(c->device);
    const int N = n_kf, L = n_lm, E = n_edge;
    for (int e = 0; e < E; e++) if (edge_kf[e] < 0 || edge_kf[e] >= N || edge_lm[e] < 0 || edge_lm[e] >= L) { c->err = "ba_shard: edge index out of range"; return nullptr; }
    // structure (every keyframe is in the system: with landmark sharding a pose may have no LOCAL edge)
    std::vector<int32_t> l_off(L + 1, 0), l_edges(E), p_off(N + 1, 0), p_edges(E), blk_i, blk_j, blk_off, ep_off, ep_pos;
    for (int e = 0; e < E; e++) { l_off[edge_lm[e] + 1]++; p_off[edge_kf[e] + 1]++; }
    for (int l = 0; l < L; l++) l_off[l + 1] += l_off[l];
    for (int a = 0; a < N; a++) p_off[a + 1] += p_off[a];
    { std::vector<int> fl(l_off.begin(), l_off.end() - 1), fp(p_off.begin(), p_off.end() - 1);
      for (int e = 0; e < E; e++) { l_edges[fl[edge_lm[e]]++] = e; p_edges[fp[edge_kf[e]]++] = e; } }
    std::vector<long long> bcount((size_t)N * N + 1, 0);
    for (int l = 0; l < L; l++)
        for (int s1 = l_off[l]; s1 < l_off[l + 1]; s1++)
            for (int s2 = l_off[l]; s2 < l_off[l + 1]; s2++) {
                int i = edge_kf[l_edges[s1]], j = edge_kf[l_edges[s2]];
                if (i <= j) bcount[(size_t)i * N + j + 1]++;
            }
    std::vector<long long> bstart((size_t)N * N, 0);
    long long run = 0;
    blk_off.push_back(0);
    for (size_t k = 0; k < (size_t)N * N; k++) {
        if (bcount[k + 1] > 0) {
            bstart[k] = run; run += bcount[k + 1];
            blk_i.push_back((int)(k / N)); blk_j.push_back((int)(k % N)); blk_off.push_back((int32_t)run);
        }
    }
    if (run > 0x7fffffffLL / 40) { c->err = "ba_shard: too many edge pairs for one shard"; return nullptr; }
    ep_pos.resize((size_t)run);
    { std::vector<long long> bf(bstart);
      long long q = 0;
      for (int e = 0; e < E; e++) {
          ep_off.push_back((int32_t)q);
          int l = edge_lm[e], i = edge_kf[e];
          for (int s2 = l_off[l]; s2 < l_off[l + 1]; s2++) { int j = edge_kf[l_edges[s2]]; if (i <= j) ep_pos[(size_t)q++] = (int32_t)bf[(size_t)i * N + j]++; }
      }
      ep_off.push_back((int32_t)q); }
    svs_ba_shard *sh = new (std::nothrow) svs_ba_shard();
    if (!sh) return nullptr;
    const int np = 6 * N, nblk = (int)blk_i.size(), n_part = std::max((L + BS_T - 1) / BS_T, 1);
    struct Seg { const void *src; size_t bytes, off; };
    std::vector<Seg> segs;
    size_t tot = 0;
    auto add = [&](const void *p, size_t bytes) { size_t o = tot; segs.push_back({p, bytes, o}); tot = align_up(tot + bytes, 256); return o; };
    size_t o_pose = add(poses, (size_t)N * 56), o_poseT = add(poses, (size_t)N * 56), o_lm = add(lms, (size_t)L * 24), o_lmT = add(lms, (size_t)L * 24);
    size_t o_ep = add(edge_kf, (size_t)E * 4), o_el = add(edge_lm, (size_t)E * 4), o_ec = add(edge_cam, (size_t)E), o_uv = add(edge_uv, (size_t)E * 16);
    size_t o_lo = add(l_off.data(), l_off.size() * 4), o_le = add(l_edges.data(), (size_t)E * 4);
    size_t o_po = add(p_off.data(), p_off.size() * 4), o_pe = add(p_edges.data(), (size_t)E * 4);
    size_t o_bi = add(blk_i.data(), blk_i.size() * 4), o_bj = add(blk_j.data(), blk_j.size() * 4), o_bo = add(blk_off.data(), blk_off.size() * 4);
    size_t o_eo = add(ep_off.data(), ep_off.size() * 4), o_epp = add(ep_pos.data(), ep_pos.size() * 4);
    size_t scratch0 = tot;
    auto res = [&](size_t bytes) { size_t o = tot; tot = align_up(tot + bytes, 256); return o; };
    size_t o_Hpl = res((size_t)E * 144), o_WD = res((size_t)E * 144), o_Hll = res((size_t)L * 72), o_Di = res((size_t)L * 72);
    size_t o_bl = res((size_t)L * 24), o_xl = res((size_t)L * 24), o_con = res((size_t)run * 288), o_xp = res((size_t)np * 8);
    size_t o_part = res((size_t)n_part * 8), o_S = res((size_t)np * (np | 1) * 8), o_g = res((size_t)np * 8), o_tmp = res((size_t)np * 8), o_tr = res((size_t)np * 4);
    if (sh->buf.reserve(tot) != cudaSuccess) { c->err = "ba_shard: cudaMalloc failed"; delete sh; return nullptr; }
    uint8_t *db = sh->buf.as<uint8_t>();
    for (const Seg &s : segs) if (s.bytes) cudaMemcpyAsync(db + s.off, s.src, s.bytes, cudaMemcpyHostToDevice, c->stream);
    cudaMemsetAsync(db + scratch0, 0, tot - scratch0, c->stream);
    BsDev &d = sh->d;
    d.N = N; d.L = L; d.E = E; d.nblk = nblk; d.npairs = (int)run;
    d.poses = (double *)(db + o_pose); d.poseT = (double *)(db + o_poseT); d.lms = (double *)(db + o_lm); d.lmT = (double *)(db + o_lmT);
    d.edge_p = (int32_t *)(db + o_ep); d.edge_l = (int32_t *)(db + o_el); d.edge_cam = db + o_ec; d.edge_uv = (double *)(db + o_uv);
    d.l_off = (int32_t *)(db + o_lo); d.l_edges = (int32_t *)(db + o_le); d.p_off = (int32_t *)(db + o_po); d.p_edges = (int32_t *)(db + o_pe);
    d.blk_i = (int32_t *)(db + o_bi); d.blk_j = (int32_t *)(db + o_bj); d.blk_off = (int32_t *)(db + o_bo);
    d.ep_off = (int32_t *)(db + o_eo); d.ep_pos = (int32_t *)(db + o_epp);
    d.Hpl = (double *)(db + o_Hpl); d.WD = (double *)(db + o_WD); d.Hll = (double *)(db + o_Hll); d.Dinv = (double *)(db + o_Di);
    d.bl = (double *)(db + o_bl); d.xl = (double *)(db + o_xl); d.contrib = (double *)(db + o_con); d.xp = (double *)(db + o_xp);
    d.partial = (double *)(db + o_part); d.Sfull = (double *)(db + o_S); d.gfull = (double *)(db + o_g); d.tmp = (double *)(db + o_tmp);
    d.tr = (int *)(db + o_tr);
    for (int i = 0; i < 4; i++) { d.K[0][i] = K_left[i]; d.K[1][i] = K_right[i]; }
    for (int i = 0; i < 7; i++) { d.ext[0][i] = ext_left[i]; d.ext[1][i] = ext_right[i]; }
    d.huber_delta = huber_delta; d.jac_mode = jacobian_mode;
    sh->n_partial = n_part;
    cudaStreamSynchronize(c->stream);
    return sh;
}

void svs_ba_shard_destroy(svs_ctx *c, svs_ba_shard *sh)
{
    if (!sh) return;
    if (c) { cudaSetDevice(c->device); cudaStreamSynchronize(c->stream); }
    sh->buf.release();
    delete sh;
}

int svs_ba_shard_lin_size(const svs_ba_shard *sh) { return sh ? 42 * sh->d.N + 2 : 0; }
int svs_ba_shard_red_size(const svs_ba_shard *sh) { return sh ? 36 * sh->d.N * sh->d.N + 6 * sh->d.N : 0; }

// lin_dev[42N+2] = [Hpp | bp | local chi2 | 0]; maxdiag_dev[1] = local max |diagonal| (poses from lin + local landmarks)
int svs_ba_shard_linearize(svs_ctx *c, svs_ba_shard *sh, double *lin_dev, double *maxdiag_dev)
{
    if (!c || !sh || !lin_dev) return SVS_ERR_ARG;
    SVS_CUDA(c, cudaSetDevice(c->device));
    BsDev &d = sh->d;
    SVS_KERNEL(c, KID_BA_WINDOW, k_bs_linearize_lm<<<sh->n_partial, BS_T, 0, c->stream>>>(d));
    SVS_KERNEL(c, KID_BA_WINDOW, k_bs_linearize_pose<<<d.N, BS_T, 0, c->stream>>>(d, lin_dev));
    SVS_KERNEL(c, KID_BA_WINDOW, k_bs_sum_partials<<<1, BS_T, 0, c->stream>>>(d.partial, sh->n_partial, lin_dev + 42 * (size_t)d.N));
    if (maxdiag_dev) SVS_KERNEL(c, KID_BA_WINDOW, k_bs_maxdiag<<<1, BS_T, 0, c->stream>>>(d, lin_dev, maxdiag_dev));
    return SVS_OK;
}

// red_dev[(6N)^2 + 6N] = local [-sum W V^-1 W^T | -sum W V^-1 bl]; flag_dev[1] int: 0 when a V was singular
int svs_ba_shard_schur(svs_ctx *c, svs_ba_shard *sh, double lambda, double *red_dev, int *flag_dev)
{
    if (!c || !sh || !red_dev || !flag_dev) return SVS_ERR_ARG;
    SVS_CUDA(c, cudaSetDevice(c->device));
    BsDev &d = sh->d;
    int one = 1;
    SVS_CUDA(c, cudaMemcpyAsync(flag_dev, &one, 4, cudaMemcpyHostToDevice, c->stream));
    SVS_CUDA(c, cudaMemsetAsync(red_dev, 0, (size_t)svs_ba_shard_red_size(sh) * 8, c->stream));
    if (d.L > 0) SVS_KERNEL(c, KID_BA_WINDOW, k_bs_dinv<<<(d.L + BS_T - 1) / BS_T, BS_T, 0, c->stream>>>(d, lambda, flag_dev));
    if (d.E > 0) SVS_KERNEL(c, KID_BA_WINDOW, k_bs_edge_records<<<(d.E + BS_T - 1) / BS_T, BS_T, 0, c->stream>>>(d));
    if (d.nblk > 0) SVS_KERNEL(c, KID_BA_WINDOW, k_bs_block_sum<<<(d.nblk * 36 + 287) / 288, 288, 0, c->stream>>>(d, red_dev));
    SVS_KERNEL(c, KID_BA_WINDOW, k_bs_g<<<d.N, BS_T, 0, c->stream>>>(d, red_dev + 36 * (size_t)d.N * d.N));
    return SVS_OK;
}

// With the all-reduced lin and red: solve, apply the trial update, and produce
// tri_dev[4] = [local trial chi2 | local landmark scale term | pose scale term (replicated) | solve ok]
int svs_ba_shard_try(svs_ctx *c, svs_ba_shard *sh, const double *lin_dev, const double *red_dev, double lambda, int flag_ok, double *tri_dev)
{
    if (!c || !sh || !lin_dev || !red_dev || !tri_dev) return SVS_ERR_ARG;
    SVS_CUDA(c, cudaSetDevice(c->device));
    BsDev &d = sh->d;
    SVS_KERNEL(c, KID_BA_WINDOW, k_bs_solve<<<1, BS_T, 0, c->stream>>>(d, lin_dev, red_dev, lambda, flag_ok, tri_dev));
    SVS_KERNEL(c, KID_BA_WINDOW, k_bs_backsub<<<sh->n_partial, BS_T, 0, c->stream>>>(d, lambda, tri_dev));
    SVS_KERNEL(c, KID_BA_WINDOW, k_bs_sum_partials<<<1, BS_T, 0, c->stream>>>(d.partial, sh->n_partial, tri_dev + 1));
    SVS_KERNEL(c, KID_BA_WINDOW, k_bs_chi2<<<sh->n_partial, BS_T, 0, c->stream>>>(d, 1));
    SVS_KERNEL(c, KID_BA_WINDOW, k_bs_sum_partials<<<1, BS_T, 0, c->stream>>>(d.partial, sh->n_partial, tri_dev));
    return SVS_OK;
}

int svs_ba_shard_accept(svs_ctx *c, svs_ba_shard *sh)
{
    if (!c || !sh) return SVS_ERR_ARG;
    SVS_CUDA(c, cudaSetDevice(c->device));
    BsDev &d = sh->d;
    int n = std::max(7 * d.N, 3 * d.L);
    SVS_KERNEL(c, KID_BA_WINDOW, k_bs_accept<<<(n + 255) / 256, 256, 0, c->stream>>>(d));
    return SVS_OK;
}

int svs_ba_shard_get(svs_ctx *c, svs_ba_shard *sh, double *poses_out, double *lms_out, double *edge_chi2_out)
{
    if (!c || !sh) return SVS_ERR_ARG;
    SVS_CUDA(c, cudaSetDevice(c->device));
    BsDev &d = sh->d;
    if (poses_out) SVS_CUDA(c, cudaMemcpyAsync(poses_out, d.poses, (size_t)d.N * 56, cudaMemcpyDeviceToHost, c->stream));
    if (lms_out && d.L) SVS_CUDA(c, cudaMemcpyAsync(lms_out, d.lms, (size_t)d.L * 24, cudaMemcpyDeviceToHost, c->stream));
    if (edge_chi2_out && d.E) {
        SVS_CUDA(c, c->d_out.reserve((size_t)d.E * 8));
        SVS_KERNEL(c, KID_BA_WINDOW, k_bs_edge_chi2<<<(d.E + 255) / 256, 256, 0, c->stream>>>(d, c->d_out.as<double>()));
        SVS_CUDA(c, cudaMemcpyAsync(edge_chi2_out, c->d_out.p, (size_t)d.E * 8, cudaMemcpyDeviceToHost, c->stream));
    }
    SVS_CUDA(c, cudaStreamSynchronize(c->stream));
    return SVS_OK;
}

}  // extern "C"
