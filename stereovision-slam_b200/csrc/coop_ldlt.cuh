// Grid-cooperative dense LDL^T solve with Eigen::LDLT's pivot order, for systems too large for one CTA's shared memory
// (the pose graph's 6K x 6K system, g2o::LinearSolverDense in LoopClosure::PoseGraphOptimization, src/loopclosure.cpp:664-672).
// Eigen's unblocked LDLT picks at step k the largest |diagonal| of the NOT yet updated trailing part (left-looking), i.e. the
// pivot order is the original diagonal sorted by decreasing magnitude: the caller writes the symmetrically permuted lower
// triangle into A once and this routine factorises it WITHOUT pivoting, right-looking, in 32-column panels:
//   (a) warp 0 of CTA 0 factorises the 32 x 32 diagonal block in shared memory (32 dependent steps, warp-synchronous)
//   (b) every CTA forward-substitutes a share of the rows below it (one row per thread)
//   (c) every CTA updates 64 x 64 tiles of the trailing matrix from shared-memory copies of the panel rows
// The right-hand side rides along as row n (so z = D^-1 L^-1 g falls out of the factorisation); CTA 0 finishes with the
// blocked back-substitution x = L^-T z.  Three barriers per panel, among the participating CTAs only (coop_sub_sync).
#pragma once
#include <cooperative_groups.h>
#include <cfloat>

#define CL_PW 32
#define CL_TILE 64
// shared scratch in doubles: panel block + its pivots + two tile operand panels
#define CL_SMEM_DOUBLES (CL_PW * (CL_PW + 1) + CL_PW + 2 * CL_TILE * (CL_PW + 1))

// A: (n + 1) x pitch, lower triangle of the PERMUTED matrix in rows 0..n-1, right-hand side (permuted) in row n.
// dvec: n pivots (out).  xs: n, solution in permuted order (out, valid in CTA 0's view after the final barrier).
// sign_io: one double in global memory, 0 on entry; Eigen's sign tracking (1 / -1 / 2 mixed / 3 zero first pivot) on exit.
// Barrier among the first n_part CTAs of the grid (a grid-wide barrier costs ~5 us with 148 x 512 threads, this one ~1 us with
// 16 CTAs): monotonic counter in global memory, release / acquire at GPU scope.  `gen` is the caller's running target.
__device__ __forceinline__ void coop_sub_sync(unsigned *ctr, unsigned n_part, unsigned &gen)
{
    __syncthreads();
    gen += n_part;
    if (threadIdx.x == 0) {
        __threadfence();
        atomicAdd(ctr, 1u);
        unsigned v;
        do { asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(ctr) : "memory"); } while (v < gen);
    }
    __syncthreads();
}

// sync_ctr: one unsigned in global memory, ZERO on entry of the kernel's first call (the routine keeps it monotonic: pass the
// same `gen` variable, initialised to 0, to every call of one kernel launch).
template <int T>
__device__ void coop_ldlt_solve(cooperative_groups::grid_group &grid, double *A, int n, int pitch, double *dvec, double *xs,
                                double *sign_io, double *sm, unsigned *sync_ctr, unsigned &gen)
{
    // the panel loop runs on as many CTAs as the trailing update has 64 x 64 tiles (at least 8): the others wait at the end
    const int nt0 = (n + 1 + CL_TILE - 1) / CL_TILE;
    const unsigned n_part = (unsigned)min((int)gridDim.x, max(8, nt0 * (nt0 + 1) / 2));
    const bool part = blockIdx.x < n_part;
    const int psz = (int)n_part * T;
    double *Ds = sm;                              // [CL_PW][CL_PW + 1]
    double *ds = Ds + CL_PW * (CL_PW + 1);        // [CL_PW]
    double *Li = ds + CL_PW;                      // [CL_TILE][CL_PW + 1]
    double *Lj = Li + CL_TILE * (CL_PW + 1);      // [CL_TILE][CL_PW + 1]  (rows of L times the pivots)
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int gtid = blockIdx.x * T + tid;
    if (part)
    for (int k0 = 0; k0 < n; k0 += CL_PW) {
        const int w = min(CL_PW, n - k0), k1 = k0 + w;
        if (blockIdx.x == 0) {
            for (int t = tid; t < w * w; t += T) { const int i = t / w, c = t - i * w; Ds[i * (CL_PW + 1) + c] = A[(size_t)(k0 + i) * pitch + k0 + c]; }
            __syncthreads();
            if (warp == 0) {
                int sign = (int)sign_io[0];
                for (int kk = 0; kk < w; kk++) {
                    const double akk = Ds[kk * (CL_PW + 1) + kk];
                    if (k0 + kk == 0 && !(fabs(akk) > 0.0)) sign = 3;
                    if (sign == 1) { if (akk < 0) sign = 2; }
                    else if (sign == -1) { if (akk > 0) sign = 2; }
                    else if (sign == 0) { if (akk > 0) sign = 1; else if (akk < 0) sign = -1; }
                    if (fabs(akk) > 0.0 && lane > kk && lane < w) {
                        const double u = Ds[lane * (CL_PW + 1) + kk], inv = 1.0 / akk;
                        for (int c = kk + 1; c <= lane; c++) Ds[lane * (CL_PW + 1) + c] -= u * (Ds[c * (CL_PW + 1) + kk] * inv);
                    }
                    __syncwarp();
                }
                if (lane == 0) sign_io[0] = (double)sign;
                // pivots, and the unit-lower block (unscaled entries divided by their pivot)
                if (lane < w) {
                    dvec[k0 + lane] = Ds[lane * (CL_PW + 1) + lane];
                    for (int c = 0; c < lane; c++) {
                        const double dk = Ds[c * (CL_PW + 1) + c], u = Ds[lane * (CL_PW + 1) + c];
                        A[(size_t)(k0 + lane) * pitch + k0 + c] = (fabs(dk) > 0.0) ? u / dk : u;
                    }
                }
            }
        }
        coop_sub_sync(sync_ctr, n_part, gen);
        // (b) rows k1..n: u_c = a_ic - sum_{j<c} u_j L11[c][j], l_ic = u_c / d_c
        {
            for (int t = tid; t < w * w; t += T) { const int i = t / w, c = t - i * w; Ds[i * (CL_PW + 1) + c] = (c < i) ? A[(size_t)(k0 + i) * pitch + k0 + c] : 0.0; }
            for (int t = tid; t < w; t += T) ds[t] = dvec[k0 + t];
            __syncthreads();
            for (int i = k1 + gtid; i <= n; i += psz) {
                double *row = A + (size_t)i * pitch + k0;
                double u[CL_PW];
#pragma unroll
                for (int c = 0; c < CL_PW; c++) {
                    if (c < w) {
                        double s = row[c];
#pragma unroll
                        for (int j = 0; j < c; j++) s -= u[j] * Ds[c * (CL_PW + 1) + j];
                        u[c] = s;
                    }
                }
#pragma unroll
                for (int c = 0; c < CL_PW; c++)
                    if (c < w) { const double dk = ds[c]; row[c] = (fabs(dk) > 0.0) ? u[c] / dk : u[c]; }
            }
        }
        coop_sub_sync(sync_ctr, n_part, gen);
        // (c) trailing update, 64 x 64 tiles of the lower triangle (rows up to n = the right-hand side row)
        {
            const int m = n + 1 - k1;                       // rows / columns k1 .. n (column n is never touched)
            if (m > 0 && k1 < n) {
                const int nt = (m + CL_TILE - 1) / CL_TILE;
                const int n_tiles = nt * (nt + 1) / 2;
                for (int tile = blockIdx.x; tile < n_tiles; tile += (int)n_part) {
                    int ti = 0, rem = tile;                 // tile -> (ti >= tj) of the lower triangle
                    while (rem > ti) { rem -= ti + 1; ti++; }
                    const int tj = rem;
                    const int i0 = k1 + ti * CL_TILE, j0 = k1 + tj * CL_TILE;
                    __syncthreads();
                    for (int t = tid; t < CL_TILE * CL_PW; t += T) {
                        const int r = t / CL_PW, c = t - r * CL_PW;
                        const int i = i0 + r, j = j0 + r;
                        Li[r * (CL_PW + 1) + c] = (i <= n && c < w) ? A[(size_t)i * pitch + k0 + c] : 0.0;
                        Lj[r * (CL_PW + 1) + c] = (j < n && c < w) ? A[(size_t)j * pitch + k0 + c] * ds[c] : 0.0;
                    }
                    __syncthreads();
                    const int tr = (tid >> 5) * 4, tc = tid & 31;          // rows tr..tr+3, columns tc and tc + 32
                    double acc[4][2] = {{0, 0}, {0, 0}, {0, 0}, {0, 0}};
#pragma unroll 8
                    for (int c = 0; c < CL_PW; c++) {
                        const double b0 = Lj[tc * (CL_PW + 1) + c], b1 = Lj[(tc + 32) * (CL_PW + 1) + c];
#pragma unroll
                        for (int q = 0; q < 4; q++) { const double a = Li[(tr + q) * (CL_PW + 1) + c]; acc[q][0] += a * b0; acc[q][1] += a * b1; }
                    }
#pragma unroll
                    for (int q = 0; q < 4; q++)
#pragma unroll
                        for (int h = 0; h < 2; h++) {
                            const int i = i0 + tr + q, j = j0 + tc + 32 * h;
                            if (i <= n && j < n && j <= i) A[(size_t)i * pitch + j] -= acc[q][h];
                        }
                }
            }
        }
        coop_sub_sync(sync_ctr, n_part, gen);
    }
    if (blockIdx.x == 0) {      // z = row n where the pivot is usable; x = L^-T z, blocks of 32 from the end
        for (int i = tid; i < n; i += T) { const double dk = dvec[i]; xs[i] = (fabs(dk) > DBL_MIN) ? A[(size_t)n * pitch + i] : 0.0; }
        __syncthreads();
        for (int b1 = n; b1 > 0; b1 -= CL_PW) {
            const int b0 = max(0, b1 - CL_PW), w = b1 - b0;
            for (int t = tid; t < w * w; t += T) { const int i = t / w, c = t - i * w; Ds[i * (CL_PW + 1) + c] = A[(size_t)(b0 + i) * pitch + b0 + c]; }
            for (int t = tid; t < w; t += T) ds[t] = xs[b0 + t];
            __syncthreads();
            if (warp == 0) {
                for (int i = w - 1; i >= 0; i--) {
                    const double xi = ds[i];
                    if (lane < i) ds[lane] -= Ds[i * (CL_PW + 1) + lane] * xi;
                    __syncwarp();
                }
            }
            __syncthreads();
            for (int t = tid; t < w; t += T) xs[b0 + t] = ds[t];
            for (int j = tid; j < b0; j += T) {
                double s = 0;
                for (int i = 0; i < w; i++) s += A[(size_t)(b0 + i) * pitch + j] * ds[i];
                xs[j] -= s;
            }
            __syncthreads();
        }
    }
    grid.sync();
}

// Variant for systems with at most T - 1 unknowns (the sharded BA's 6N x 6N reduced system, N <= 85): the whole 32-column panel
// (all rows below the diagonal block, one row per thread) is factorised by CTA 0 in shared memory with ONE block barrier per
// column (the row owner divides by the pivot lazily, so a step only reads column kk of the other rows), which replaces the
// warp-sequential diagonal block + the separate row pass of the general routine and one of its three barriers per panel.
// Shared scratch: (n + 1) * (CL_PW + 1) + CL_PW + 2 * CL_TILE * (CL_PW + 1) doubles.
#define CL_SMALL_SMEM_DOUBLES(n) ((size_t)((n) + 1) * (CL_PW + 1) + CL_PW + 2 * CL_TILE * (CL_PW + 1))
template <int T>
__device__ void coop_ldlt_solve_small(cooperative_groups::grid_group &grid, double *A, int n, int pitch, double *dvec, double *xs,
                                      double *sign_io, double *sm, unsigned *sync_ctr, unsigned &gen)
{
    double *P = sm;                                        // [n + 1][CL_PW + 1]
    double *ds = P + (size_t)(n + 1) * (CL_PW + 1);        // [CL_PW]
    double *Li = ds + CL_PW, *Lj = Li + CL_TILE * (CL_PW + 1);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int nt0 = (n + 1 + CL_TILE - 1) / CL_TILE;
    const unsigned n_part = (unsigned)min((int)gridDim.x, max(8, nt0 * (nt0 + 1) / 2));
    if (blockIdx.x < n_part)
    for (int k0 = 0; k0 < n; k0 += CL_PW) {
        const int w = min(CL_PW, n - k0), k1 = k0 + w, rows = n + 1 - k0;
        if (blockIdx.x == 0) {
            for (int t = tid; t < rows * w; t += T) { const int i = t / w, c = t - i * w; P[i * (CL_PW + 1) + c] = A[(size_t)(k0 + i) * pitch + k0 + c]; }
            __syncthreads();
            int sign = (int)sign_io[0];
            for (int kk = 0; kk < w; kk++) {
                const double akk = P[kk * (CL_PW + 1) + kk];
                if (k0 + kk == 0 && !(fabs(akk) > 0.0)) sign = 3;
                if (sign == 1) { if (akk < 0) sign = 2; }
                else if (sign == -1) { if (akk > 0) sign = 2; }
                else if (sign == 0) { if (akk > 0) sign = 1; else if (akk < 0) sign = -1; }
                const int i = kk + 1 + tid;
                if (fabs(akk) > 0.0 && i < rows) {
                    const double u = P[i * (CL_PW + 1) + kk] * (1.0 / akk);
                    const int cmax = min(w - 1, i);
                    for (int c = kk + 1; c <= cmax; c++) P[i * (CL_PW + 1) + c] -= u * P[c * (CL_PW + 1) + kk];
                }
                __syncthreads();
            }
            for (int t = tid; t < rows * w; t += T) {
                const int i = t / w, c = t - i * w;
                if (i < c) continue;
                const double dk = P[c * (CL_PW + 1) + c];
                if (i == c) dvec[k0 + c] = dk;
                else A[(size_t)(k0 + i) * pitch + k0 + c] = (fabs(dk) > 0.0) ? P[i * (CL_PW + 1) + c] / dk : P[i * (CL_PW + 1) + c];
            }
            if (tid == 0) sign_io[0] = (double)sign;
        }
        coop_sub_sync(sync_ctr, n_part, gen);
        {
            const int m = n + 1 - k1;
            if (m > 0 && k1 < n) {
                for (int t = tid; t < w; t += T) ds[t] = dvec[k0 + t];
                const int nt = (m + CL_TILE - 1) / CL_TILE;
                const int n_tiles = nt * (nt + 1) / 2;
                for (int tile = blockIdx.x; tile < n_tiles; tile += (int)n_part) {
                    int ti = 0, rem = tile;
                    while (rem > ti) { rem -= ti + 1; ti++; }
                    const int tj = rem;
                    const int i0 = k1 + ti * CL_TILE, j0 = k1 + tj * CL_TILE;
                    __syncthreads();
                    for (int t = tid; t < CL_TILE * CL_PW; t += T) {
                        const int r = t / CL_PW, c = t - r * CL_PW;
                        const int i = i0 + r, j = j0 + r;
                        Li[r * (CL_PW + 1) + c] = (i <= n && c < w) ? A[(size_t)i * pitch + k0 + c] : 0.0;
                        Lj[r * (CL_PW + 1) + c] = (j < n && c < w) ? A[(size_t)j * pitch + k0 + c] * ds[c] : 0.0;
                    }
                    __syncthreads();
                    const int tr = (tid >> 5) * 4, tc = tid & 31;
                    double acc[4][2] = {{0, 0}, {0, 0}, {0, 0}, {0, 0}};
#pragma unroll 8
                    for (int c = 0; c < CL_PW; c++) {
                        const double b0 = Lj[tc * (CL_PW + 1) + c], b1 = Lj[(tc + 32) * (CL_PW + 1) + c];
#pragma unroll
                        for (int q = 0; q < 4; q++) { const double a = Li[(tr + q) * (CL_PW + 1) + c]; acc[q][0] += a * b0; acc[q][1] += a * b1; }
                    }
#pragma unroll
                    for (int q = 0; q < 4; q++)
#pragma unroll
                        for (int h = 0; h < 2; h++) {
                            const int i = i0 + tr + q, j = j0 + tc + 32 * h;
                            if (i <= n && j < n && j <= i) A[(size_t)i * pitch + j] -= acc[q][h];
                        }
                }
            }
        }
        coop_sub_sync(sync_ctr, n_part, gen);
    }
    if (blockIdx.x == 0) {      // x = L^-T z with z = row n, blocks of 32 from the end
        double *Ds = P;
        for (int i = tid; i < n; i += T) { const double dk = dvec[i]; xs[i] = (fabs(dk) > DBL_MIN) ? A[(size_t)n * pitch + i] : 0.0; }
        __syncthreads();
        for (int b1 = n; b1 > 0; b1 -= CL_PW) {
            const int b0 = max(0, b1 - CL_PW), w = b1 - b0;
            for (int t = tid; t < w * w; t += T) { const int i = t / w, c = t - i * w; Ds[i * (CL_PW + 1) + c] = A[(size_t)(b0 + i) * pitch + b0 + c]; }
            for (int t = tid; t < w; t += T) ds[t] = xs[b0 + t];
            __syncthreads();
            if (warp == 0) {
                for (int i = w - 1; i >= 0; i--) {
                    const double xi = ds[i];
                    if (lane < i) ds[lane] -= Ds[i * (CL_PW + 1) + lane] * xi;
                    __syncwarp();
                }
            }
            __syncthreads();
            for (int t = tid; t < w; t += T) xs[b0 + t] = ds[t];
            for (int j = tid; j < b0; j += T) {
                double s = 0;
                for (int i = 0; i < w; i++) s += A[(size_t)(b0 + i) * pitch + j] * ds[i];
                xs[j] -= s;
            }
            __syncthreads();
        }
    }
    grid.sync();
}

// ------------------------------------------------------------------------------------------------ banded system, one CTA
// A symmetric system whose entries vanish beyond `hb` sub-diagonals (the reduced camera system of a window in which a
// landmark is seen by keyframes at most hb/6 apart) is factored IN NATURAL ORDER inside one CTA's shared memory: no fill
// outside the band, n steps of <= hb (hb + 3) / 2 updates, one block barrier per step and no grid barrier at all.
// Band storage: Bb[i * (hb + 1) + d] = A(i, i - d), d = 0 .. min(i, hb); z[i] = rhs; both are overwritten.  Behind z the
// routine keeps the inverse pivots (n doubles) and the (a, b) item table (hb (hb + 3) / 2 words).  hb <= 255.
// One SM issues 4 warp-instructions per clock, so the step is kept to ~10 instructions per update: every thread owns
// the same (a, b) offsets in every step, the pivot's reciprocal is computed once (by the thread that finishes the next
// pivot) and the triangular solves multiply by it.
// On exit x[i] (shared or global) holds the solution and *sign_out Eigen's sign tracking as in coop_ldlt_solve.
#define CL_BAND_DOUBLES(n, hb) ((size_t)(n) * ((hb) + 1) + 2 * (size_t)(n) + ((size_t)(hb) * ((hb) + 3) / 2 + 2) / 2 + 1)
template <int T, bool DBG = false>
__device__ void band_ldlt_solve_cta(double *Bb, double *z, int n, int hb, double *x, double *sign_out, long long *dbg = nullptr)
{
    long long c0 = 0, c1 = 0, c2 = 0, c3 = 0, acc0 = 0, acc1 = 0, acc2 = 0, acc3 = 0;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int w = hb + 1, w1 = w + 1;
    double *invd = z + n;
    unsigned *tab = reinterpret_cast<unsigned *>(invd + n);
    for (int a = 1 + warp; a <= hb; a += T / 32)
        for (int b = lane; b <= a; b += 32) tab[a * (a + 1) / 2 - 1 + b] = (unsigned)a | ((unsigned)b << 16);
    int sign = 0;
    auto pivot = [&](int k, double d) {       // thread 1 only: sign tracking + reciprocal of pivot k
        if (k == 0 && !(fabs(d) > 0.0)) sign = 3;
        if (sign == 1) { if (d < 0) sign = 2; }
        else if (sign == -1) { if (d > 0) sign = 2; }
        else if (sign == 0) { if (d > 0) sign = 1; else if (d < 0) sign = -1; }
        invd[k] = (fabs(d) > DBL_MIN) ? 1.0 / d : 0.0;
    };
    if (tid == 1) pivot(0, Bb[0]);
    for (int k = 0; k < n; k++) {
        if (DBG) c0 = clock64();
        __syncthreads();
        if (DBG) c1 = clock64();
        const double inv = invd[k];
        const int m = min(hb, n - 1 - k);
        double *base = Bb + (size_t)k * w;
        if (DBG) { if (inv == 12345.0) c2 = 0; c2 = clock64(); }
        if (inv != 0.0 && m > 0) {
            const double zs = z[k] * inv;
            const int cnt = m * (m + 3) / 2;
            for (int t = tid; t < cnt; t += T) {
                const unsigned ab = tab[t];
                const int a = (int)(ab & 0xffffu), b = (int)(ab >> 16);
                const double ua = base[a * w1];
                if (b == 0) z[k + a] -= ua * zs;
                else {
                    const double v = base[a * w1 - b] - ua * (base[b * w1] * inv);
                    base[a * w1 - b] = v;
                    if (DBG && t == 1) c3 = clock64();
                    if (t == 1) pivot(k + 1, v);
                }
            }
        } else if (tid == 1 && k + 1 < n) pivot(k + 1, base[w]);
        if (DBG) { const long long c4 = clock64(); acc0 += c1 - c0; acc1 += c2 - c1; acc2 += c3 - c2; acc3 += c4 - c3; }
    }
    if (DBG && tid == 1) { dbg[0] = acc0; dbg[1] = acc1; dbg[2] = acc2; dbg[3] = acc3; }
    __syncthreads();
    // z = D^-1 y, then x = L^-T z column by column from the last row up (one warp, no reductions)
    for (int i = tid; i < n; i += T) z[i] *= invd[i];
    __syncthreads();
    if (warp == 0) {
        for (int k = n - 1; k >= 1; k--) {
            __syncwarp();
            const double xk = z[k];
            const int m = min(hb, k);
            for (int a = 1 + lane; a <= m; a += 32) z[k - a] -= Bb[(size_t)k * w + a] * invd[k - a] * xk;
        }
    }
    __syncthreads();
    for (int i = tid; i < n; i += T) x[i] = z[i];
    if (tid == 1) sign_out[0] = (double)sign;
    __syncthreads();
}

// (A register-resident variant — the active window as 4 x 4 tiles of an (R/4)^2 thread grid, the pivot column published
// through shared memory — was measured SLOWER, 2090 vs 1290 cycles per step at hb = 41: with four warps on the SM every
// dependent instruction is exposed, whereas the 16 warps above hide each other's latencies.  Not kept.)
