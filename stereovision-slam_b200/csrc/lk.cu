// Pyramidal Lucas-Kanade tracking (rows a2/a3): what cv::calcOpticalFlowPyrLK computes at reference
// src/frontend.cpp:105-109 (left -> right) and :353-357 (last -> current), with
// OPTFLOW_USE_INITIAL_FLOW.  Algorithm per SURVEY.md Appendix A.4/A.5.
//
// One warp per keypoint; the warp walks the pyramid levels coarse -> fine inside the kernel.
// Per level the (win+3)^2 patch of the previous image is staged in shared memory, the Scharr
// derivative patch is derived from it, each lane keeps its <=ceil(win^2/32) window samples
// (I, Ix, Iy as int16 values) in registers; a (win+1+12)^2 region of the next image is staged once
// per level (restaged only if the window leaves it), every iteration samples its window from that
// region and reduces the two mismatch sums with redux.sync.  All window sums are exact
// integers (OpenCV accumulates the same integers in f32 lanes), the 2x2 solve is f32 with the
// exact operation order of OpenCV — compile with --fmad=false.
#include "svs_internal.h"
#include <cfloat>
#include <cstdlib>

__device__ __forceinline__ int lk_refl101(int i, int n)
{
    if (n == 1) return 0;
    while (i < 0 || i >= n) i = (i < 0) ? -i : 2 * (n - 1) - i;
    return i;
}
// f32( (sum over the warp of s) * 2^-20 ), correctly rounded from the EXACT integer sum (which may exceed 32 bits):
// the low 16 bits and the signed high part are reduced separately with the hardware 32-bit reductions (< 2^21 and < 2^27
// in magnitude, so both convert to f32 exactly), and ONE fused multiply-add combines them: hi * 2^-4 + lo * 2^-20 is the
// exact scaled sum rounded once — the same value as __ll2float_rn(sum) * 2^-20 (scaling by a power of two commutes with
// the rounding), in 8 instructions instead of a 64-bit recombination and conversion.
__device__ __forceinline__ float warp_sum_scaled_exact(int s)
{
    int lo = s & 0xFFFF, hi = s >> 16;
    int slo = __reduce_add_sync(0xffffffffu, lo);
    int shi = __reduce_add_sync(0xffffffffu, hi);
    return __fmaf_rn((float)shi, 0.0625f, __fmul_rn((float)slo, 1.f / (1 << 20)));
}
__device__ __forceinline__ int cvfloor(float v) { return __float2int_rd(v); }
__device__ __forceinline__ int descale(int x, int n) { return (x + (1 << (n - 1))) >> n; }

// Which window sample (row y, column x) is slot q of a lane.  All window sums are exact integers, so any assignment of the
// WIN^2 samples to (lane, slot) gives the same result.  Generic: sample k = lane + 32 q in row-major order.  WIN = 11 (the
// reference's window): lane = 11 g + x owns column x, rows g, g + 3, g + 6 (, g + 9) — slot q is a CONSTANT byte offset
// (3 rows) from slot 0 in every staged buffer, so the inner loops address with immediates instead of one offset register per
// slot; the column-10 samples of rows 2, 5, 8 (no lane 32) are slot 3 of lanes 22, 23, 24.
template <int WIN>
__device__ __forceinline__ void lk_sample_pos(int lane, int q, int &y, int &x)
{
    if (WIN == 11) {
        const int g = lane / 11, xx = lane - 11 * g;
        if (q < 3 || g < 2) { y = g + 3 * q; x = xx; }
        else { y = 2 + 3 * min(lane - 22, 2); x = 10; }
    } else {
        const int k = lane + 32 * q;
        y = k / WIN; x = k - y * WIN;
    }
}
template <int WIN>
__device__ __forceinline__ bool lk_sample_valid(int lane, int q)
{
    if (WIN == 11) return q < 3 || lane < 25;
    return lane + 32 * q < WIN * WIN;
}

// Word k of the staged next-image region (row k / RWORDS, bytes 4 (k % RWORDS) ..+3): pixels of the reflect-101-padded image
// at (rx0 + 4q + i, ry0 + r).  Interior regions are aligned 32-bit loads; elsewhere the row is reflected once per word and
// only words that straddle the left / right image border are assembled from reflected bytes.
template <int RWORDS>
__device__ __forceinline__ uint32_t lk_region_word(const uint8_t *__restrict__ J, int Jw, int Jh, int Js, int rx0, int ry0, int k, bool interior)
{
    const int r = k / RWORDS, q = k - r * RWORDS;
    int y = ry0 + r;
    const int x = rx0 + 4 * q;
    if (interior) return __ldg(reinterpret_cast<const uint32_t *>(J + (size_t)y * Js + x));
    y = lk_refl101(y, Jh);
    const uint8_t *row = J + (size_t)y * Js;
    if (x >= 0 && x + 4 <= Jw) return __ldg(reinterpret_cast<const uint32_t *>(row + x));
    uint32_t v = 0;
#pragma unroll
    for (int i = 0; i < 4; i++) v |= (uint32_t)__ldg(row + lk_refl101(x + i, Jw)) << (8 * i);
    return v;
}

#define LK_WARPS 4
#define LK_MARGIN 6      // the next-image region staged per level extends this many pixels around the first window

// (8 CTAs per SM at 64 registers compile without spills but measure the same as 6 at 80: the kernel is issue-bound.)
template <int WIN>
__global__ void __launch_bounds__(LK_WARPS * 32, WIN <= 11 ? 6 : 1)
k_lk_track(const __grid_constant__ PyrDesc prev, const __grid_constant__ PyrDesc next, const int32_t *__restrict__ pt_img, const float *__restrict__ prev_xy,
           float *__restrict__ next_xy, int n_pts, int max_iter, double eps2, float eps_lo, float eps_hi, uint8_t *__restrict__ status)
{
    constexpr int PW = WIN + 3;                 // previous-image patch (window + bilinear + Scharr halo)
    constexpr int DW = WIN + 1;                 // derivative / next-image patch
    constexpr int IWORDS = (PW + 3 + 3) / 4;    // aligned 32-bit words covering a patch row at any byte offset
    constexpr int IP = 4 * IWORDS;              // smem row pitch of the previous-image patch (bytes)
    // Next-image REGION: DW + 2*LK_MARGIN rows x RWORDS aligned words, staged once per level around the first window.
    // The iterations of a level move the window by fractions of a pixel, so almost all of them sample straight from
    // this region and touch no global memory (one L2 round trip per level instead of one per iteration).
    constexpr int RH = DW + 2 * LK_MARGIN;
    constexpr int RWORDS = (DW + 2 * LK_MARGIN + 3 + 3) / 4;
    constexpr int RP = 4 * RWORDS;              // region row pitch (bytes); the border-case patch uses the same pitch
    constexpr int RN = (RH * RWORDS + 31) / 32; // region words per lane
    constexpr int IN = (PW * IWORDS + 31) / 32; // previous-patch words per lane
    constexpr int NPL = (WIN * WIN + 31) / 32;  // window samples per lane
    __shared__ __align__(16) uint8_t sI[LK_WARPS][PW * IP];
    __shared__ short2 sD[LK_WARPS][DW * DW];
    __shared__ short2 sT[LK_WARPS][DW * PW];     // separable Scharr: vertical pass
    __shared__ __align__(16) uint8_t sJ[LK_WARPS][RH * RP];
    const int W_BITS = 14;
    const float half = (WIN - 1) * 0.5f;
    int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    int pt = blockIdx.x * LK_WARPS + warp;
    if (pt >= n_pts) return;
    int img = pt_img[pt];
    float px0 = prev_xy[2 * pt], py0 = prev_xy[2 * pt + 1];
    float nxt_x = next_xy[2 * pt], nxt_y = next_xy[2 * pt + 1];
    bool st = true;
    // cvFloor(NaN) is INT_MIN in OpenCV (cvtss2si's "integer indefinite"), so a non-finite start point fails every bounds
    // test below and ends with status 0 and its input value; __float2int_rd maps NaN to 0 instead.  Inf saturates to an
    // out-of-range integer either way but would drag the other coordinate through the iterations: NaN and inf both leave
    // here (warp-uniform: one keypoint per warp).
    if (!(fabsf(px0) <= FLT_MAX && fabsf(py0) <= FLT_MAX && fabsf(nxt_x) <= FLT_MAX && fabsf(nxt_y) <= FLT_MAX)) {
        if (lane == 0) status[pt] = 0;
        return;
    }
    uint8_t *mI = sI[warp], *mJ = sJ[warp];
    short2 *mD = sD[warp], *mT = sT[warp];
    int nlev = min(prev.nlev, next.nlev);
    // this lane's window samples: smem offsets in the previous-image patch and in the next-image region
    int offI[NPL], offJ[NPL], offD[NPL];
#pragma unroll
    for (int q = 0; q < NPL; q++) {
        int y, x;
        lk_sample_pos<WIN>(lane, q, y, x);
        offI[q] = (y + 1) * IP + (x + 1); offJ[q] = y * RP + x; offD[q] = y * DW + x;
    }

    for (int level = nlev - 1; level >= 0; level--) {
        const uint8_t *I = prev.base + (size_t)img * prev.img_pitch + prev.off[level];
        const uint8_t *J = next.base + (size_t)img * next.img_pitch + next.off[level];
        int Iw = prev.w[level], Ih = prev.h[level], Is = prev.stride[level];
        int Jw = next.w[level], Jh = next.h[level], Js = next.stride[level];
        float lscale = (float)(1. / (1 << level));
        float ppx = __fmul_rn(px0, lscale), ppy = __fmul_rn(py0, lscale);
        if (level == nlev - 1) { nxt_x = __fmul_rn(nxt_x, lscale); nxt_y = __fmul_rn(nxt_y, lscale); }
        else { nxt_x = __fmul_rn(nxt_x, 2.f); nxt_y = __fmul_rn(nxt_y, 2.f); }
        ppx = __fsub_rn(ppx, half); ppy = __fsub_rn(ppy, half);
        int ix = cvfloor(ppx), iy = cvfloor(ppy);
        if (ix < -WIN || ix >= Iw || iy < -WIN || iy >= Ih) { if (level == 0) st = false; continue; }
        float a = __fsub_rn(ppx, (float)ix), b = __fsub_rn(ppy, (float)iy);
        int w00 = __float2int_rn(__fmul_rn(__fmul_rn(__fsub_rn(1.f, a), __fsub_rn(1.f, b)), (float)(1 << W_BITS)));
        int w01 = __float2int_rn(__fmul_rn(__fmul_rn(a, __fsub_rn(1.f, b)), (float)(1 << W_BITS)));
        int w10 = __float2int_rn(__fmul_rn(__fmul_rn(__fsub_rn(1.f, a), b), (float)(1 << W_BITS)));
        int w11 = (1 << W_BITS) - w00 - w01 - w10;
        float nx = __fsub_rn(nxt_x, half), ny = __fsub_rn(nxt_y, half);

        __syncwarp();
        // ---- issue the next-image region loads first (they are independent of everything below), then stage the
        //      previous-image patch; both global round trips overlap.  The region is a piece of the REFLECT-101-PADDED next
        //      image (what OpenCV's border handling reads), so windows that hang over the image border sample from it like
        //      interior ones: no per-iteration border path.
        bool reg_valid = false, reg_interior = false;
        int rx0 = 0, ry0 = 0;
        uint32_t jr[RN];
        {
            int jx = cvfloor(nx), jy = cvfloor(ny);
            if (jx >= -WIN && jx < Jw && jy >= -WIN && jy < Jh) {
                ry0 = jy - LK_MARGIN; rx0 = (jx - LK_MARGIN) & ~3;
                reg_valid = true;
                reg_interior = rx0 >= 0 && ry0 >= 0 && rx0 + RP <= Jw && ry0 + RH <= Jh;
                if (reg_interior) {      // the common case: aligned word loads into registers, stored after the patch below
#pragma unroll
                    for (int u = 0; u < RN; u++) {
                        int k = lane + 32 * u;
                        jr[u] = (k < RH * RWORDS) ? lk_region_word<RWORDS>(J, Jw, Jh, Js, rx0, ry0, k, true) : 0u;
                    }
                }
            }
        }
        // stage the previous-image patch: mI[j][ioff + i] = I(ix-1+i, iy-1+j), reflect-101 outside.
        // Interior patches (the common case) are staged with aligned 32-bit loads and no border arithmetic.
        int ioff = 0;
        if (ix - 1 >= 0 && iy - 1 >= 0 && ix - 1 + PW <= Iw && iy - 1 + PW <= Ih) {
            ioff = (ix - 1) & 3;
            const uint8_t *base = I + (size_t)(iy - 1) * Is + (ix - 1 - ioff);
            uint32_t ir[IN];
#pragma unroll
            for (int u = 0; u < IN; u++) {
                int k = lane + 32 * u;
                int j = k / IWORDS, q = k - j * IWORDS;
                ir[u] = (k < PW * IWORDS) ? __ldg(reinterpret_cast<const uint32_t *>(base + (size_t)j * Is) + q) : 0u;
            }
#pragma unroll
            for (int u = 0; u < IN; u++) {
                int k = lane + 32 * u;
                if (k < PW * IWORDS) reinterpret_cast<uint32_t *>(mI)[k] = ir[u];
            }
        } else {        // the patch hangs over the border: the same word layout, rows / edge bytes reflected
            ioff = (ix - 1) & 3;
#pragma unroll 1
            for (int k = lane; k < PW * IWORDS; k += 32)
                reinterpret_cast<uint32_t *>(mI)[k] = lk_region_word<IWORDS>(I, Iw, Ih, Is, (ix - 1) & ~3, iy - 1, k, false);
        }
        if (reg_interior) {
#pragma unroll
            for (int u = 0; u < RN; u++) {
                int k = lane + 32 * u;
                if (k < RH * RWORDS) reinterpret_cast<uint32_t *>(mJ)[k] = jr[u];
            }
        } else if (reg_valid) {          // the region hangs over the image border: reflected rows / bytes (rolled: rare, keeps the code small)
#pragma unroll 1
            for (int k = lane; k < RH * RWORDS; k += 32)
                reinterpret_cast<uint32_t *>(mJ)[k] = lk_region_word<RWORDS>(J, Jw, Jh, Js, rx0, ry0, k, false);
        }
        __syncwarp();
        // Scharr derivative patch (zero outside the image: BORDER_CONSTANT on the derivative buffer), separable:
        //   pass 1  (PW columns x DW rows)  c = 3 (p[j] + p[j+2]) + 10 p[j+1]   (vertical smoothing, for dx)
        //                                   d = p[j+2] - p[j]                  (vertical difference, for dy)
        //   pass 2  (DW x DW)               dx = c[i+2] - c[i],  dy = 3 (d[i] + d[i+2]) + 10 d[i+1]
        // exact integers either way; ~1/4 of the instructions of the direct 3x3 form
        // lane = (half-warp hw, column i): each half-warp walks half of the rows with a sliding 3-row register window,
        // so there is no index arithmetic in the loops
        {
            constexpr int HR = (DW + 1) / 2;                  // rows per half-warp
            const int hw = lane >> 4, i = lane & 15;
            const int j0 = hw * HR, j1 = min(DW, j0 + HR);
            for (int ib = i; ib < PW; ib += 16) {              // one pass for WIN <= 13 (PW <= 16), column blocks beyond
                const uint8_t *p = mI + j0 * IP + ib + ioff;
                int p0 = p[0], p1 = p[IP];
                for (int j = j0; j < j1; j++) {
                    int p2 = p[2 * IP];
                    mT[j * PW + ib] = make_short2((short)(3 * (p0 + p2) + 10 * p1), (short)(p2 - p0));
                    p0 = p1; p1 = p2; p += IP;
                }
            }
            __syncwarp();
            const bool inside = ix >= 0 && iy >= 0 && ix + DW <= Iw && iy + DW <= Ih;   // every derivative sample is in the image
            for (int ib = i; ib < DW; ib += 16) {
                const short2 *t = mT + j0 * PW + ib;
                for (int j = j0; j < j1; j++) {
                    short2 t0 = t[0], t1 = t[1], t2 = t[2];
                    short2 d = make_short2((short)(t2.x - t0.x), (short)(3 * (t0.y + t2.y) + 10 * t1.y));
                    if (!inside) {
                        int X = ix + ib, Y = iy + j;
                        if (!(X >= 0 && X < Iw && Y >= 0 && Y < Ih)) d = make_short2(0, 0);
                    }
                    mD[j * DW + ib] = d;
                    t += PW;
                }
            }
        }
        __syncwarp();
        int Iv[NPL], Ixv[NPL], Iyv[NPL];
        int pA11 = 0, pA12 = 0, pA22 = 0;
#pragma unroll
        for (int q = 0; q < NPL; q++) {
            Iv[q] = 0; Ixv[q] = 0; Iyv[q] = 0;
            if (lk_sample_valid<WIN>(lane, q)) {
                const uint8_t *p = mI + offI[q] + ioff;
                int ival = descale(p[0] * w00 + p[1] * w01 + p[IP] * w10 + p[IP + 1] * w11, W_BITS - 5);
                const short2 *dp = mD + offD[q];
                short2 d00 = dp[0], d01 = dp[1], d10 = dp[DW], d11 = dp[DW + 1];
                int ixv = descale(d00.x * w00 + d01.x * w01 + d10.x * w10 + d11.x * w11, W_BITS);
                int iyv = descale(d00.y * w00 + d01.y * w01 + d10.y * w10 + d11.y * w11, W_BITS);
                Iv[q] = (short)ival; Ixv[q] = (short)ixv; Iyv[q] = (short)iyv;
                pA11 += Ixv[q] * Ixv[q]; pA12 += Ixv[q] * Iyv[q]; pA22 += Iyv[q] * Iyv[q];
            }
        }
        float A11 = warp_sum_scaled_exact(pA11), A12 = warp_sum_scaled_exact(pA12), A22 = warp_sum_scaled_exact(pA22);
        float D = __fsub_rn(__fmul_rn(A11, A22), __fmul_rn(A12, A12));
        float dd = __fsub_rn(A11, A22);
        float rad = __fadd_rn(__fmul_rn(dd, dd), __fmul_rn(__fmul_rn(4.f, A12), A12));
        float minEig = __fdiv_rn(__fsub_rn(__fadd_rn(A22, A11), __fsqrt_rn(rad)), (float)(2 * WIN * WIN));
        if ((double)minEig < 1e-4 || D < FLT_EPSILON) { if (level == 0) st = false; continue; }
        D = __fdiv_rn(1.f, D);
        float pdx = 0.f, pdy = 0.f;
        // window origins (jx, jy) that are in OpenCV's range ([-WIN, size)) AND covered by the staged region: one unsigned
        // range test per axis
        int fx_lo = 0, fx_n = -1, fy_lo = 0, fy_n = -1;
        auto set_fast_range = [&]() {
            fx_lo = max(-WIN, rx0); fx_n = min(Jw - 1, rx0 + RP - DW) - fx_lo;
            fy_lo = max(-WIN, ry0); fy_n = min(Jh - 1, ry0 + RH - DW) - fy_lo;
            if (fx_n < 0 || fy_n < 0) fx_n = fy_n = -1;
        };
        if (reg_valid) set_fast_range();
        for (int j = 0; j < max_iter; j++) {
            int jx = cvfloor(nx), jy = cvfloor(ny);
            const bool fast = fx_n >= 0 && (unsigned)(jx - fx_lo) <= (unsigned)fx_n && (unsigned)(jy - fy_lo) <= (unsigned)fy_n;
            if (!fast) {
                if (jx < -WIN || jx >= Jw || jy < -WIN || jy >= Jh) { if (level == 0) st = false; break; }
                // the window left the staged region (or there is none yet): restage around the current window
                __syncwarp();
                ry0 = jy - LK_MARGIN; rx0 = (jx - LK_MARGIN) & ~3;
                const bool interior = rx0 >= 0 && ry0 >= 0 && rx0 + RP <= Jw && ry0 + RH <= Jh;
#pragma unroll 1
                for (int k = lane; k < RH * RWORDS; k += 32)
                    reinterpret_cast<uint32_t *>(mJ)[k] = lk_region_word<RWORDS>(J, Jw, Jh, Js, rx0, ry0, k, interior);
                reg_valid = true;
                set_fast_range();
                __syncwarp();
            }
            a = __fsub_rn(nx, (float)jx); b = __fsub_rn(ny, (float)jy);
            w00 = __float2int_rn(__fmul_rn(__fmul_rn(__fsub_rn(1.f, a), __fsub_rn(1.f, b)), (float)(1 << W_BITS)));
            w01 = __float2int_rn(__fmul_rn(__fmul_rn(a, __fsub_rn(1.f, b)), (float)(1 << W_BITS)));
            w10 = __float2int_rn(__fmul_rn(__fmul_rn(__fsub_rn(1.f, a), b), (float)(1 << W_BITS)));
            w11 = (1 << W_BITS) - w00 - w01 - w10;
            const uint8_t *pJ = mJ + (jy - ry0) * RP + (jx - rx0);
            int pb1 = 0, pb2 = 0;
#pragma unroll
            for (int q = 0; q < NPL; q++) {
                if (lk_sample_valid<WIN>(lane, q)) {
                    const uint8_t *p = pJ + offJ[q];
                    int diff = descale(p[0] * w00 + p[1] * w01 + p[RP] * w10 + p[RP + 1] * w11, W_BITS - 5) - Iv[q];
                    pb1 += diff * Ixv[q]; pb2 += diff * Iyv[q];
                }
            }
            float b1 = warp_sum_scaled_exact(pb1), b2 = warp_sum_scaled_exact(pb2);
            float dx = __fmul_rn(__fsub_rn(__fmul_rn(A12, b2), __fmul_rn(A22, b1)), D);
            float dy = __fmul_rn(__fsub_rn(__fmul_rn(A12, b1), __fmul_rn(A11, b2)), D);
            nx = __fadd_rn(nx, dx); ny = __fadd_rn(ny, dy);
            nxt_x = __fadd_rn(nx, half); nxt_y = __fadd_rn(ny, half);
            // OpenCV: delta.ddot(delta) <= eps^2 in double.  A float estimate decides everything except a 1e-4-wide
            // band around the threshold, where the exact double expression is evaluated (values are warp-uniform).
            float d2f = __fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy));
            bool conv = d2f < eps_lo;
            if (!conv && d2f <= eps_hi) conv = (double)dx * (double)dx + (double)dy * (double)dy <= eps2;
            if (conv) break;
            // |(double)(dx + pdx)| < 0.01 for a float argument is exactly |.| <= 0.01f (0.01f < 0.01 < nextafter(0.01f))
            if (j > 0 && fabsf(__fadd_rn(dx, pdx)) <= 0.01f && fabsf(__fadd_rn(dy, pdy)) <= 0.01f) {
                nxt_x = __fsub_rn(nxt_x, __fmul_rn(dx, 0.5f)); nxt_y = __fsub_rn(nxt_y, __fmul_rn(dy, 0.5f));
                break;
            }
            pdx = dx; pdy = dy;
        }
        if (st && level == 0) {   // OpenCV's error pass re-checks the final position
            int jx = cvfloor(__fsub_rn(nxt_x, half)), jy = cvfloor(__fsub_rn(nxt_y, half));
            if (jx < -WIN || jx >= Jw || jy < -WIN || jy >= Jh) st = false;
        }
    }
    if (lane == 0) {
        next_xy[2 * pt] = nxt_x; next_xy[2 * pt + 1] = nxt_y;
        status[pt] = st ? 1 : 0;
    }
}

int svs_i_lk(svs_ctx *c, const PyrDesc &prev, const PyrDesc &next, const int32_t *pt_img, const float *prev_xy,
             float *next_xy, int n_pts, int win, int max_iter, double eps, uint8_t *status)
{
    if (n_pts <= 0) return SVS_OK;
    if (max_iter > 100) max_iter = 100;
    if (max_iter < 0) max_iter = 0;
    if (eps < 0) eps = 0;
    if (eps > 10) eps = 10;
    double eps2 = eps * eps;
    // the kernel decides convergence from a float estimate outside [eps_lo, eps_hi] and evaluates OpenCV's double expression
    // inside that band: the band edges are launch constants
    const float eps2f = (float)eps2, eps_lo = eps2f * 0.9999f, eps_hi = eps2f * 1.0001f;
    int blocks = (n_pts + LK_WARPS - 1) / LK_WARPS;
#define LK_CASE(WN)                                                                                          \
    case WN:                                                                                                 \
        k_lk_track<WN><<<blocks, LK_WARPS * 32, 0, c->stream>>>(prev, next, pt_img, prev_xy, next_xy, n_pts, \
                                                                max_iter, eps2, eps_lo, eps_hi, status);     \
        break;
    svs_i_prof_begin(c, KID_LK);
    switch (win) {
        LK_CASE(5) LK_CASE(7) LK_CASE(9) LK_CASE(11) LK_CASE(13) LK_CASE(15) LK_CASE(21)
    default:
        svs_i_prof_end(c);
        SVS_FAIL(c, SVS_ERR_ARG, "lk: window size must be one of 5,7,9,11,13,15,21");
    }
#undef LK_CASE
    svs_i_prof_end(c);
    SVS_LAUNCH_CHECK(c);
    return SVS_OK;
}
