// Block-cooperative pivoted LDLT (Eigen::LDLT order of operations) shared by k_ba_window and the sharded solver.
#pragma once
#include <cfloat>
#ifndef BA_T
#define BA_T 256
#endif

// Single-warp variant for small systems (n <= 128, i.e. windows of up to 21 keyframes): the factorisation is a chain of n
// dependent steps with O(n) work each, so a 512-thread CTA spends it in __syncthreads (measured: 76 us of a 450 us LM
// iteration at n = 60).  One warp with __syncwarp only does the same steps in ~1/6 of the time; the other warps wait at
// the barrier below.  Same pivoting and the same left-looking update order as block_ldlt_solve, lane = row.
__device__ __forceinline__ bool warp_ldlt_solve(double *S, int pitch, int n, const double *g, double *x, int *tr, double *tmp, int *s_flag)
{
    const int tid = threadIdx.x, lane = tid & 31;
    if (tid < 32) {
        int sign = 0;
        bool zero_first = false;
        for (int k = 0; k < n; k++) {
            double best = -1.0;
            int bi = k;
            for (int i = k + lane; i < n; i += 32) { double v = fabs(S[i * pitch + i]); if (v > best) { best = v; bi = i; } }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                double ov = __shfl_xor_sync(0xffffffffu, best, o);
                int oi = __shfl_xor_sync(0xffffffffu, bi, o);
                if (ov > best || (ov == best && oi < bi)) { best = ov; bi = oi; }
            }
            const int piv = bi;
            if (lane == 0) tr[k] = piv;
            if (piv != k) {
                for (int t = lane; t < n + 1; t += 32) {
                    if (t < k) { double a = S[k * pitch + t]; S[k * pitch + t] = S[piv * pitch + t]; S[piv * pitch + t] = a; }
                    else if (t == k) { double a = S[k * pitch + k]; S[k * pitch + k] = S[piv * pitch + piv]; S[piv * pitch + piv] = a; }
                    else if (t < piv) { double a = S[t * pitch + k]; S[t * pitch + k] = S[piv * pitch + t]; S[piv * pitch + t] = a; }
                    else if (t > piv && t < n) { double a = S[t * pitch + k]; S[t * pitch + k] = S[t * pitch + piv]; S[t * pitch + piv] = a; }
                }
                __syncwarp();
            }
            for (int j = lane; j < k; j += 32) tmp[j] = S[j * pitch + j] * S[k * pitch + j];
            __syncwarp();
            if (k > 0) {
                for (int i = k + lane; i < n; i += 32) {
                    const double *row = S + i * pitch;
                    double a0 = 0, a1 = 0;
                    int j = 0;
                    for (; j + 1 < k; j += 2) { a0 += row[j] * tmp[j]; a1 += row[j + 1] * tmp[j + 1]; }
                    if (j < k) a0 += row[j] * tmp[j];
                    S[i * pitch + k] -= a0 + a1;
                }
                __syncwarp();
            }
            double akk = S[k * pitch + k];
            bool valid = fabs(akk) > 0.0;
            if (k == 0 && !valid) { for (int j = lane; j < n; j += 32) tr[j] = j; sign = 0; zero_first = true; __syncwarp(); break; }
            if (valid) for (int i = k + 1 + lane; i < n; i += 32) S[i * pitch + k] /= akk;
            if (sign == 1) { if (akk < 0) sign = 2; }
            else if (sign == -1) { if (akk > 0) sign = 2; }
            else if (sign == 0) { if (akk > 0) sign = 1; else if (akk < 0) sign = -1; }
            __syncwarp();
        }
        (void)zero_first;
        bool ok = (sign == 1 || sign == 0);
        if (ok) {
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
        if (lane == 0) *s_flag = ok ? 1 : 0;
    }
    __syncthreads();
    bool ok = *s_flag != 0;
    __syncthreads();
    return ok;
}

// Pivoted LDLT of the symmetric n x n matrix S (lower triangle used, row pitch `pitch`), then solve S x = g.
// Returns Eigen::LDLT::isPositive().  All threads of the CTA must call it.
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
            for (int t = tid; t < n + 1; t += BA_T) {
                if (t < k) { double a = S[k * pitch + t]; S[k * pitch + t] = S[piv * pitch + t]; S[piv * pitch + t] = a; }
                else if (t == k) { double a = S[k * pitch + k]; S[k * pitch + k] = S[piv * pitch + piv]; S[piv * pitch + piv] = a; }
                else if (t < piv) { double a = S[t * pitch + k]; S[t * pitch + k] = S[piv * pitch + t]; S[piv * pitch + t] = a; }
                else if (t > piv && t < n) { double a = S[t * pitch + k]; S[t * pitch + k] = S[t * pitch + piv]; S[t * pitch + piv] = a; }
            }
            __syncthreads();
        }
        for (int j = tid; j < k; j += BA_T) tmp[j] = S[j * pitch + j] * S[k * pitch + j];
        __syncthreads();
        if (k > 0) {
            // 4 lanes per row: interleaved partial dot products combined in a fixed order
            int sub = tid & 3;
            int trips = (n - k + (BA_T >> 2) - 1) / (BA_T >> 2);   // uniform trip count (shuffles below)
            for (int m = 0; m < trips; m++) {
                int i = k + (tid >> 2) + m * (BA_T >> 2);
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
        if (k == 0 && !valid) { for (int j = tid; j < n; j += BA_T) tr[j] = j; sign = 0; __syncthreads(); break; }
        if (valid) for (int i = k + 1 + tid; i < n; i += BA_T) S[i * pitch + k] /= akk;
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

