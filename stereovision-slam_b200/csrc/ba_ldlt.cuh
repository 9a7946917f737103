// Block-cooperative pivoted LDLT (Eigen::LDLT order of operations) shared by k_ba_window and the sharded solver.
#pragma once
#include <cfloat>

// Pre-permuted variant (needs a second n x pitch buffer S2 and n ints).  Eigen's unblocked LDLT picks, at step k, the
// largest |diagonal| entry of the trailing part, and that part of the diagonal is NOT updated before its turn
// (left-looking): the pivot order is simply the diagonal sorted by decreasing magnitude, known before any arithmetic.
// So: rank the diagonal in parallel, write the symmetrically permuted matrix into S2 once, then run the left-looking
// factorisation without pivoting.  Every row group recomputes the pivot row's dot product redundantly (same lanes, same
// order => identical bits), which leaves ONE __syncthreads per step instead of five (a step has O(n) work, so the
// block-wide version is barrier-bound: measured 76 us per solve at n = 60 with 512 threads).  Arithmetic is identical to
// block_ldlt_solve except for the order of exactly equal pivots.  The pivots d_k go to the unused upper-triangle slot
// S2[k][k+1] (the pitch is n + 1) so that S2[k][k] stays read-only while the groups of step k still read it.
template <int BT>
__device__ __forceinline__ bool block_ldlt_solve_pp(const double *S, double *S2, int pitch, int n, const double *g, double *x,
                                                    int *perm, double *tmp)
{
    const int tid = threadIdx.x, lane = tid & 31;
    // rank of |S_ii| in decreasing order (ties: lower index first)
    for (int i = tid; i < n; i += BT) {
        const double di = fabs(S[i * pitch + i]);
        int r = 0;
        for (int j = 0; j < n; j++) { double dj = fabs(S[j * pitch + j]); r += (dj > di || (dj == di && j < i)) ? 1 : 0; }
        perm[r] = i;
    }
    __syncthreads();
    for (int t = tid; t < n * n; t += BT) {
        int a = t / n, b = t - a * n;
        if (b <= a) S2[a * pitch + b] = S[perm[a] * pitch + perm[b]];
    }
    __syncthreads();
    int sign = 0;
    bool zero_first = false;
    const int sub = tid & 3, grp = tid >> 2;
    for (int k = 0; k < n; k++) {
        const double *rowk = S2 + k * pitch;
        const int trips = (n - k + (BT >> 2) - 1) / (BT >> 2);
        double akk = 0;
        for (int m = 0; m < trips; m++) {
            const int i = k + grp + m * (BT >> 2);
            const double *rowi = S2 + (i < n ? i : k) * pitch;
            double acc_i = 0, acc_k = 0;
            for (int j = sub; j < k; j += 4) {
                const double lk = rowk[j], t = S2[j * pitch + j + 1] * lk;
                acc_k += lk * t;
                acc_i += rowi[j] * t;
            }
            acc_i += __shfl_xor_sync(0xffffffffu, acc_i, 1); acc_k += __shfl_xor_sync(0xffffffffu, acc_k, 1);
            acc_i += __shfl_xor_sync(0xffffffffu, acc_i, 2); acc_k += __shfl_xor_sync(0xffffffffu, acc_k, 2);
            akk = rowk[k] - acc_k;
            const bool valid = fabs(akk) > 0.0;
            if (i < n && sub == 0) {
                if (i == k) S2[k * pitch + k + 1] = akk;
                else { double v = S2[i * pitch + k] - acc_i; S2[i * pitch + k] = valid ? v / akk : v; }
            }
        }
        if (k == 0 && !(fabs(akk) > 0.0)) { zero_first = true; break; }
        if (sign == 1) { if (akk < 0) sign = 2; }
        else if (sign == -1) { if (akk > 0) sign = 2; }
        else if (sign == 0) { if (akk > 0) sign = 1; else if (akk < 0) sign = -1; }
        __syncthreads();
    }
    __syncthreads();
    if (zero_first) sign = 0;
    bool ok = (sign == 1 || sign == 0);
    if (ok && tid < 32) {   // triangular solves on one warp, in the permuted order
        for (int i = lane; i < n; i += 32) tmp[i] = g[perm[i]];
        __syncwarp();
        for (int i = 0; i < n; i++) {
            double xi = tmp[i];
            for (int j = i + 1 + lane; j < n; j += 32) tmp[j] -= S2[j * pitch + i] * xi;
            __syncwarp();
        }
        for (int i = lane; i < n; i += 32) { double d = S2[i * pitch + i + 1]; tmp[i] = (fabs(d) > DBL_MIN) ? tmp[i] / d : 0.0; }
        __syncwarp();
        for (int i = n - 1; i >= 0; i--) {
            double xi = tmp[i];
            for (int j = lane; j < i; j += 32) tmp[j] -= S2[i * pitch + j] * xi;
            __syncwarp();
        }
        for (int i = lane; i < n; i += 32) x[perm[i]] = tmp[i];
    }
    __syncthreads();
    return ok;
}

// Pivoted LDLT of the symmetric n x n matrix S (lower triangle used, row pitch `pitch`), then solve S x = g.
// Returns Eigen::LDLT::isPositive().  All threads of the CTA must call it.
template <int BT>
__device__ __forceinline__ bool block_ldlt_solve(double *S, int pitch, int n, const double *g, double *x, int *tr, double *tmp, int *s_piv)
{
    int tid = threadIdx.x, lane = tid & 31;
    int sign = 0;
    for (int k = 0; k < n; k++) {
        if (tid < 32) {
            double best = -1.0;
            int bi = k;
            for (int i = k + lane; i < n; i += 32) { double v = fabs(S[i * pitch + i]); if (v > best) { best = v; bi = i; } }
            for (int o = 16; o > 0; o >>= 1) {
                double ov = __shfl_xor_sync(0xffffffffu, best, o);
                int oi = __shfl_xor_sync(0xffffffffu, bi, o);
                if (ov > best || (ov == best && oi < bi)) { best = ov; bi = oi; }
            }
            if (lane == 0) { *s_piv = bi; tr[k] = bi; }
        }
        __syncthreads();
        int piv = *s_piv;
        if (piv != k) {
            // disjoint element swaps: [0,k) row part, (piv,n) column part, (k,piv) cross part, diagonal
            for (int t = tid; t < n + 1; t += BT) {
                if (t < k) { double a = S[k * pitch + t]; S[k * pitch + t] = S[piv * pitch + t]; S[piv * pitch + t] = a; }
                else if (t == k) { double a = S[k * pitch + k]; S[k * pitch + k] = S[piv * pitch + piv]; S[piv * pitch + piv] = a; }
                else if (t < piv) { double a = S[t * pitch + k]; S[t * pitch + k] = S[piv * pitch + t]; S[piv * pitch + t] = a; }
                else if (t > piv && t < n) { double a = S[t * pitch + k]; S[t * pitch + k] = S[t * pitch + piv]; S[t * pitch + piv] = a; }
            }
            __syncthreads();
        }
        for (int j = tid; j < k; j += BT) tmp[j] = S[j * pitch + j] * S[k * pitch + j];
        __syncthreads();
        if (k > 0) {
            // 4 lanes per row: interleaved partial dot products combined in a fixed order
            int sub = tid & 3;
            int trips = (n - k + (BT >> 2) - 1) / (BT >> 2);   // uniform trip count (shuffles below)
            for (int m = 0; m < trips; m++) {
                int i = k + (tid >> 2) + m * (BT >> 2);
                double acc = 0;
                if (i < n) for (int j = sub; j < k; j += 4) acc += S[i * pitch + j] * tmp[j];
                double a1 = __shfl_xor_sync(0xffffffffu, acc, 1);
                acc += a1;
                double a2 = __shfl_xor_sync(0xffffffffu, acc, 2);
                acc += a2;
                if (i < n && sub == 0) S[i * pitch + k] -= acc;
            }
            __syncthreads();
        }
        double akk = S[k * pitch + k];
        bool valid = fabs(akk) > 0.0;
        if (k == 0 && !valid) { for (int j = tid; j < n; j += BT) tr[j] = j; sign = 0; __syncthreads(); break; }
        if (valid) for (int i = k + 1 + tid; i < n; i += BT) S[i * pitch + k] /= akk;
        if (sign == 1) { if (akk < 0) sign = 2; }
        else if (sign == -1) { if (akk > 0) sign = 2; }
        else if (sign == 0) { if (akk > 0) sign = 1; else if (akk < 0) sign = -1; }
        __syncthreads();
    }
    bool ok = (sign == 1 || sign == 0);
    if (ok && tid < 32) {   // triangular solves on one warp (column-oriented = same rounding as the row-oriented loop)
        for (int i = lane; i < n; i += 32) x[i] = g[i];
        __syncwarp();
        if (lane == 0) for (int k = 0; k < n; k++) if (tr[k] != k) { double t = x[k]; x[k] = x[tr[k]]; x[tr[k]] = t; }
        __syncwarp();
        for (int i = 0; i < n; i++) {
            double xi = x[i];
            for (int j = i + 1 + lane; j < n; j += 32) x[j] -= S[j * pitch + i] * xi;
            __syncwarp();
        }
        for (int i = lane; i < n; i += 32) { double d = S[i * pitch + i]; x[i] = (fabs(d) > DBL_MIN) ? x[i] / d : 0.0; }
        __syncwarp();
        for (int i = n - 1; i >= 0; i--) {
            double xi = x[i];
            for (int j = lane; j < i; j += 32) x[j] -= S[i * pitch + j] * xi;
            __syncwarp();
        }
        if (lane == 0) for (int k = n - 1; k >= 0; k--) if (tr[k] != k) { double t = x[k]; x[k] = x[tr[k]]; x[tr[k]] = t; }
    }
    __syncthreads();
    return ok;
}

