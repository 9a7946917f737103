// GFTT corner detection (row a1): what cv::GFTTDetector::detect computes at reference
// src/frontend.cpp:51 (detector created at :24), plus the tracked-feature mask of :42-47.
// Bit-exact against cv::cornerMinEigenVal / cv::goodFeaturesToTrack.
//
//   k_mask_boxes        the cv::rectangle mask
//   k_corner_response   Sobel-3 structure tensor -> 3x3 box sums (f64 running column sums, the
//                       order OpenCV's ColumnSum uses — it is history dependent, so each column is
//                       marched top to bottom by one lane) -> min eigenvalue (f32) + masked max
//   k_corner_select     threshold at quality*max, 3x3 non-max suppression, compaction to sort keys
//   k_corner_greedy     one CTA per image: bitonic sort (value desc, index desc) + greedy min-distance
//
// Compile this file with --fmad=false: every float operation below must round exactly once where
// OpenCV rounds; the fused operations are spelled __fmaf_rn explicitly.
#include "svs_internal.h"
#include "tma.cuh"
#include <algorithm>
#include <climits>
#include <cstdlib>

__device__ __forceinline__ int g_refl101(int i, int n)
{
    if (n == 1) return 0;
    while (i < 0 || i >= n) i = (i < 0) ? -i : 2 * (n - 1) - i;
    return i;
}
__device__ __forceinline__ unsigned f2ord(float f)
{
    unsigned b = __float_as_uint(f);
    return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}
__device__ __forceinline__ float ord2f(unsigned u)
{
    return __uint_as_float((u & 0x80000000u) ? (u & 0x7FFFFFFFu) : ~u);
}

// ------------------------------------------------------------------------------------------
__global__ void k_mask_boxes(uint8_t *__restrict__ mask, int w, int h, const int32_t *__restrict__ occ_off,
                             int n_img, const float *__restrict__ occ_xy)
{
    int box = blockIdx.x;
    // image owning this box: last i with occ_off[i] <= box
    int lo = 0, hi = n_img;
    while (hi - lo > 1) { int mid = (lo + hi) >> 1; if (occ_off[mid] <= box) lo = mid; else hi = mid; }
    float x = occ_xy[2 * box], y = occ_xy[2 * box + 1];
    int x0 = __float2int_rn(__fsub_rn(x, 10.f)), x1 = __float2int_rn(__fadd_rn(x, 10.f));
    int y0 = __float2int_rn(__fsub_rn(y, 10.f)), y1 = __float2int_rn(__fadd_rn(y, 10.f));
    if (x1 < 0 || y1 < 0 || x0 >= w || y0 >= h) return;
    x0 = max(x0, 0); y0 = max(y0, 0); x1 = min(x1, w - 1); y1 = min(y1, h - 1);
    int bw = x1 - x0 + 1, bh = y1 - y0 + 1;
    uint8_t *m = mask + (size_t)lo * w * h;
    for (int i = threadIdx.x; i < bw * bh; i += blockDim.x) m[(size_t)(y0 + i / bw) * w + x0 + i % bw] = 0;
}

// ------------------------------------------------------------------------------------------
// One warp marches a 28-column strip (32 lanes = 28 outputs + 2 halo columns each side) from the top
// row to the bottom row.  Per row and lane: one byte load, 8 warp shuffles, ~40 f32/f64 ops, one f32 store.
#define CR_OUT 28
struct RowF { float rx, ry; };

__global__ void __launch_bounds__(128)
k_corner_response(const uint8_t *__restrict__ img_base, size_t img_pitch, const int32_t *__restrict__ img_ids,
                  int w, int h, int stride, const uint8_t *__restrict__ mask, float *__restrict__ eig,
                  unsigned *__restrict__ maxbits, float ks, float k2, int body, const int *__restrict__ only_if)
{
    const unsigned FULL = 0xffffffffu;
    int lane = threadIdx.x & 31;
    int strip = blockIdx.x * 4 + (threadIdx.x >> 5);
    int n_strips = (w + CR_OUT - 1) / CR_OUT;
    if (strip >= n_strips) return;
    int slot = blockIdx.y;
    if (only_if && !only_if[slot]) return;      // fallback pass of the tiled kernel: only images it flagged as inexact
    int img = img_ids ? img_ids[slot] : slot;
    const uint8_t *src = img_base + (size_t)img * img_pitch;
    float *out = eig + (size_t)slot * w * h;
    const uint8_t *msk = mask ? mask + (size_t)slot * w * h : nullptr;

    int c = strip * CR_OUT - 2 + lane;         // this lane's column (may lie outside the image)
    int cl = g_refl101(max(min(c, w + 1), -2), w);  // column actually loaded
    bool real = (c >= 0 && c < w);
    // lanes that supply cov(c-1) / cov(c+1) under BORDER_REFLECT_101 of the covariance image
    int lane_l = (c - 1 >= 0) ? lane - 1 : lane + 1;
    int lane_r = (c + 1 <= w - 1) ? lane + 1 : lane - 1;
    lane_l = max(0, min(31, lane_l)); lane_r = max(0, min(31, lane_r));
    bool fused = c < body;
    bool writer = real && lane >= 2 && lane < 2 + CR_OUT;

    auto rowfilter = [&](int y) -> RowF {
        int pv = __ldg(src + (size_t)y * stride + cl);
        int pl = __shfl_sync(FULL, pv, max(lane - 1, 0));
        int pr = __shfl_sync(FULL, pv, min(lane + 1, 31));
        RowF r;
        float fl = (float)pl, fc = (float)pv, fr = (float)pr;
        r.rx = __fsub_rn(fr, fl);
        float t = __fmul_rn(fl, ks);
        if (fused) { t = __fmaf_rn(fc, k2, t); t = __fmaf_rn(fr, ks, t); }
        else { t = __fadd_rn(t, __fmul_rn(fc, k2)); t = __fadd_rn(t, __fmul_rn(fr, ks)); }
        r.ry = t;
        return r;
    };

    RowF R0 = rowfilter(0);
    RowF R1 = rowfilter(min(1, h - 1));
    RowF prevR = R1, curR = R0, nextR = R1;
    double SUMxx = 0, SUMxy = 0, SUMyy = 0;
    double Axx = 0, Axy = 0, Ayy = 0, Bxx = 0, Bxy = 0, Byy = 0;   // rs(r-2), rs(r-1)
    float vmax = -INFINITY;
    bool any = false;

    auto emit = [&](int y, double nxx, double nxy, double nyy, double oxx, double oxy, double oyy) {
        double sxx = SUMxx + nxx, sxy = SUMxy + nxy, syy = SUMyy + nyy;
        float a = __fmul_rn((float)sxx, 0.5f), b = (float)sxy, cc = __fmul_rn((float)syy, 0.5f);
        float t = __fsub_rn(a, cc);
        float rad = __fadd_rn(__fmul_rn(t, t), __fmul_rn(b, b));
        float lam = __fsub_rn(__fadd_rn(a, cc), __fsqrt_rn(rad));
        if (writer) {
            out[(size_t)y * w + c] = lam;
            if (!msk || msk[(size_t)y * w + c]) { vmax = fmaxf(vmax, lam); any = true; }
        }
        SUMxx = sxx - oxx; SUMxy = sxy - oxy; SUMyy = syy - oyy;
    };

    for (int r = 0; r < h; r++) {
        if (r > 0) { prevR = curR; curR = nextR; }
        if (r + 1 <= h - 1) { if (r + 1 >= 2) nextR = rowfilter(r + 1); else nextR = R1; }
        else nextR = prevR;                                  // row h reflects to row h-2
        float dx = __fmaf_rn(__fadd_rn(prevR.rx, nextR.rx), ks, __fmul_rn(curR.rx, k2));
        float dy = __fsub_rn(nextR.ry, prevR.ry);
        float cxx = __fmul_rn(dx, dx), cxy = __fmul_rn(dx, dy), cyy = __fmul_rn(dy, dy);
        float lxx = __shfl_sync(FULL, cxx, lane_l), rxx = __shfl_sync(FULL, cxx, lane_r);
        float lxy = __shfl_sync(FULL, cxy, lane_l), rxy = __shfl_sync(FULL, cxy, lane_r);
        float lyy = __shfl_sync(FULL, cyy, lane_l), ryy = __shfl_sync(FULL, cyy, lane_r);
        double Cxx = ((double)lxx + (double)cxx) + (double)rxx;
        double Cxy = ((double)lxy + (double)cxy) + (double)rxy;
        double Cyy = ((double)lyy + (double)cyy) + (double)ryy;
        if (r == 0) {
            Bxx = Cxx; Bxy = Cxy; Byy = Cyy;
        } else {
            if (r == 1) {   // SUM = (0 + rs(-1)) + rs(0) with rs(-1) = rs(1); old row of y = 0 is rs(1) too
                SUMxx = Cxx + Bxx; SUMxy = Cxy + Bxy; SUMyy = Cyy + Byy;
                emit(0, Cxx, Cxy, Cyy, Cxx, Cxy, Cyy);
            } else {
                emit(r - 1, Cxx, Cxy, Cyy, Axx, Axy, Ayy);
            }
            Axx = Bxx; Axy = Bxy; Ayy = Byy;
            Bxx = Cxx; Bxy = Cxy; Byy = Cyy;
        }
    }
    if (h >= 2) emit(h - 1, Axx, Axy, Ayy, 0, 0, 0);   // rs(h) = rs(h-2)

    unsigned mb = any ? f2ord(vmax) : 0u;
    for (int o = 16; o > 0; o >>= 1) mb = max(mb, __shfl_xor_sync(FULL, mb, o));
    if (lane == 0 && mb) atomicMax(maxbits + slot, mb);
}

// ------------------------------------------------------------------------------------------
// Tiled, TMA-staged corner response (the batched path).  The sequential march above exists because OpenCV's box filter keeps
// a RUNNING f64 column sum (SUM += new row; out = SUM; SUM -= old row), whose rounding depends on the history of the column.
// But its intermediate values are only ever T_y = rs(y) + rs(y+1) and O_y = T_(y-1) + rs(y+1) (rs = f64 row sums of three f32
// products): if every one of those sums is EXACTLY representable, every add / subtract of the running scheme is exact and the
// result is the exact sum, whatever the order.  That holds for most pixels of an 8-bit image (products between 2^-47 and 2^-1)
// but not all: a cancelling Sobel sum leaves a ~1e-10 residual whose square sits 60+ bits below a strong neighbour.  So: tiles compute out = (rs(y-1) + rs(y)) + rs(y+1) fully in parallel and CHECK exactness with TwoSum; a
// tile that sees one inexact sum flags its image, and the flagged images are recomputed by the
// sequential kernel — bit-exactness is kept unconditionally, and the grid grows from 66 CTAs to one tile per 64x16 outputs.
// Staging: the 96 x 20 source box (16 bytes left of the tile, see SVS_CR_BOX_W) arrives by cp.async.bulk.tensor (UTMALDG) into one of two buffers while the
// previous tile is computed; zero-filled out-of-image bytes are patched to reflect-101 in shared memory.
#define CT_W 64
#define CT_H 16
#define CT_BYTES (SVS_CR_BOX_W * SVS_CR_BOX_H)
__device__ __forceinline__ bool twosum_exact(double a, double b, double s)
{   // s = fl(a + b); exact iff the TwoSum error term is zero (no contraction: this file is built with --fmad=false)
    const double bb = s - a;
    return ((a - (s - bb)) + (b - bb)) == 0.0;
}
__global__ void __launch_bounds__(256)
k_corner_response_tma(const __grid_constant__ CUtensorMap tm, const int32_t *__restrict__ img_ids, int n_slots, int w, int h,
                      const uint8_t *__restrict__ mask, float *__restrict__ eig, unsigned *__restrict__ tile_max, int *__restrict__ inexact,
                      float ks, float k2, int body, int tiles_x, int tiles_y)
{
    __shared__ __align__(128) uint8_t tile_s[2][1920];
    __shared__ float rxs[SVS_CR_BOX_H][CT_W + 3], rys[SVS_CR_BOX_H][CT_W + 3];
    __shared__ float cxx[CT_H + 2][CT_W + 3], cxy[CT_H + 2][CT_W + 3], cyy[CT_H + 2][CT_W + 3];
    __shared__ __align__(8) uint64_t bar[2];
    __shared__ unsigned s_max[8];
    __shared__ int s_bad;
    const int tid = threadIdx.x, lane = tid & 31;
    const int per_img = tiles_x * tiles_y, total = n_slots * per_img;
    if (tid == 0) { tma::mbar_init(&bar[0], 1); tma::mbar_init(&bar[1], 1); tma::fence_init(); }
    __syncthreads();
    auto issue = [&](int t, int buf) {
        const int n = t / per_img, r = t - n * per_img, ty = r / tiles_x, tx = r - ty * tiles_x;
        tma::mbar_expect_tx(&bar[buf], CT_BYTES);
        tma::load_3d(tile_s[buf], &tm, &bar[buf], tx * CT_W - 16, ty * CT_H - 2, img_ids ? img_ids[n] : n);
    };
    int t = blockIdx.x;
    if (tid == 0 && t < total) issue(t, 0);
    for (int it = 0; t < total; t += gridDim.x, it++) {
        const int buf = it & 1, tn = t + gridDim.x;
        if (tid == 0 && tn < total) { tma::fence_proxy_async(); issue(tn, buf ^ 1); }
        const int slot = t / per_img, r0 = t - slot * per_img, ty0 = r0 / tiles_x, tx0 = r0 - ty0 * tiles_x;
        const int ox0 = tx0 * CT_W, oy0 = ty0 * CT_H, bx0 = ox0 - 16, by0 = oy0 - 2;
        uint8_t *tb = tile_s[buf];
        if (tid == 0) s_bad = 0;
        tma::mbar_wait(&bar[buf], (it >> 1) & 1);
        // image reflect-101 of the one row / column outside the image a Sobel window can touch (rows first, then columns)
        if (by0 < 0 || by0 + SVS_CR_BOX_H > h) {
            for (int i = tid; i < 2 * (SVS_CR_BOX_W / 4); i += 256) {
                const int k = i / (SVS_CR_BOX_W / 4), q = i - k * (SVS_CR_BOX_W / 4);
                const int y = k ? h : -1, ys = k ? h - 2 : 1, rd = y - by0, rs = ys - by0;
                if (rd >= 0 && rd < SVS_CR_BOX_H && rs >= 0 && rs < SVS_CR_BOX_H)
                    reinterpret_cast<uint32_t *>(tb + rd * SVS_CR_BOX_W)[q] = reinterpret_cast<const uint32_t *>(tb + rs * SVS_CR_BOX_W)[q];
            }
            __syncthreads();
        }
        if (bx0 < 0 || bx0 + SVS_CR_BOX_W > w) {
            for (int i = tid; i < 2 * SVS_CR_BOX_H; i += 256) {
                const int k = i / SVS_CR_BOX_H, r = i - k * SVS_CR_BOX_H;
                const int x = k ? w : -1, xs = k ? w - 2 : 1, cd = x - bx0, cs = xs - bx0;
                if (cd >= 0 && cd < SVS_CR_BOX_W && cs >= 0 && cs < SVS_CR_BOX_W) tb[r * SVS_CR_BOX_W + cd] = tb[r * SVS_CR_BOX_W + cs];
            }
        }
        __syncthreads();
        // stage 1: Sobel row pass at tile rows 0..19 (image y = by0 + ty), columns image x = ox0 - 1 + tc, tc = 0..65
        for (int i = tid; i < SVS_CR_BOX_H * (CT_W + 2); i += 256) {
            const int ty = i / (CT_W + 2), tc = i - ty * (CT_W + 2);
            const int x = ox0 - 1 + tc;
            const uint8_t *p = tb + ty * SVS_CR_BOX_W + tc + 15;      // tile column of image x is x - bx0 = tc + 15
            const float fl = (float)p[-1], fc = (float)p[0], fr = (float)p[1];
            float tt = __fmul_rn(fl, ks);
            if (x < body) { tt = __fmaf_rn(fc, k2, tt); tt = __fmaf_rn(fr, ks, tt); }
            else { tt = __fadd_rn(tt, __fmul_rn(fc, k2)); tt = __fadd_rn(tt, __fmul_rn(fr, ks)); }
            rxs[ty][tc] = __fsub_rn(fr, fl);
            rys[ty][tc] = tt;
        }
        __syncthreads();
        // stage 2: covariances at image rows oy0 - 1 + cy (cy = 0..17), same columns
        for (int i = tid; i < (CT_H + 2) * (CT_W + 2); i += 256) {
            const int cy = i / (CT_W + 2), tc = i - cy * (CT_W + 2);
            const float dx = __fmaf_rn(__fadd_rn(rxs[cy][tc], rxs[cy + 2][tc]), ks, __fmul_rn(rxs[cy + 1][tc], k2));
            const float dy = __fsub_rn(rys[cy + 2][tc], rys[cy][tc]);
            cxx[cy][tc] = __fmul_rn(dx, dx); cxy[cy][tc] = __fmul_rn(dx, dy); cyy[cy][tc] = __fmul_rn(dy, dy);
        }
        __syncthreads();
        // stage 3: 3x3 box sums with reflect-101 of the COVARIANCE image, min eigenvalue, masked maximum
        float vmax = -INFINITY;
        bool any = false, bad = false;
        {
            const int xo = tid & 63, seg = tid >> 6, ox = ox0 + xo;
            if (ox < w) {
                const int xl = (ox - 1 >= 0) ? ox - 1 : 1, xr = (ox + 1 <= w - 1) ? ox + 1 : w - 2;
                const int cl = xl - ox0 + 1, cc = xo + 1, cr = xr - ox0 + 1;
                auto rowsum = [&](int yy, double &sxx, double &sxy, double &syy) {
                    const int yr = yy < 0 ? -yy : (yy > h - 1 ? 2 * (h - 1) - yy : yy);
                    const int cy = yr - oy0 + 1;
                    sxx = ((double)cxx[cy][cl] + (double)cxx[cy][cc]) + (double)cxx[cy][cr];
                    sxy = ((double)cxy[cy][cl] + (double)cxy[cy][cc]) + (double)cxy[cy][cr];
                    syy = ((double)cyy[cy][cl] + (double)cyy[cy][cc]) + (double)cyy[cy][cr];
                };
                const int y0 = oy0 + 4 * seg;
                double axx, axy, ayy, bxx, bxy, byy, nxx, nxy, nyy;
                if (y0 < h) { rowsum(y0 - 1, axx, axy, ayy); rowsum(y0, bxx, bxy, byy); }
                for (int j = 0; j < 4; j++) {
                    const int y = y0 + j;
                    if (y >= h) break;
                    rowsum(y + 1, nxx, nxy, nyy);
                    const double txx = axx + bxx, txy = axy + bxy, tyy = ayy + byy;
                    const double sxx = txx + nxx, sxy = txy + nxy, syy = tyy + nyy;
                    bad |= !(twosum_exact(axx, bxx, txx) && twosum_exact(txx, nxx, sxx) && twosum_exact(axy, bxy, txy) &&
                             twosum_exact(txy, nxy, sxy) && twosum_exact(ayy, byy, tyy) && twosum_exact(tyy, nyy, syy));
                    const float a = __fmul_rn((float)sxx, 0.5f), b = (float)sxy, c2 = __fmul_rn((float)syy, 0.5f);
                    const float tt = __fsub_rn(a, c2);
                    const float rad = __fadd_rn(__fmul_rn(tt, tt), __fmul_rn(b, b));
                    const float lam = __fsub_rn(__fadd_rn(a, c2), __fsqrt_rn(rad));
                    const size_t o = (size_t)slot * w * h + (size_t)y * w + ox;
                    eig[o] = lam;
                    if (!mask || mask[o]) { vmax = fmaxf(vmax, lam); any = true; }
                    axx = bxx; axy = bxy; ayy = byy; bxx = nxx; bxy = nxy; byy = nyy;
                }
            }
        }
        unsigned mb = any ? f2ord(vmax) : 0u;
        for (int o = 16; o > 0; o >>= 1) mb = max(mb, __shfl_xor_sync(0xffffffffu, mb, o));
        if (lane == 0) s_max[tid >> 5] = mb;
        if (bad) s_bad = 1;
        __syncthreads();
        if (tid == 0) {
            unsigned m = 0;
            for (int i = 0; i < 8; i++) m = max(m, s_max[i]);
            tile_max[t] = m;
            if (s_bad) inexact[slot] = 1;
        }
        __syncthreads();      // all reads of this buffer / the stage arrays are done before the next iteration reuses them
    }
}
// per image: maximum over its tiles, unless the image was flagged (then the sequential fallback computes it)
__global__ void k_corner_tile_max(const unsigned *__restrict__ tile_max, int per_img, const int *__restrict__ inexact, unsigned *__restrict__ maxbits)
{
    const int slot = blockIdx.x;
    if (inexact[slot]) return;
    unsigned m = 0;
    for (int i = threadIdx.x; i < per_img; i += 32) m = max(m, tile_max[(size_t)slot * per_img + i]);
    for (int o = 16; o > 0; o >>= 1) m = max(m, __shfl_xor_sync(0xffffffffu, m, o));
    if (threadIdx.x == 0) maxbits[slot] = m;
}

// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
k_corner_select(const float *__restrict__ eig, const uint8_t *__restrict__ mask, int w, int h,
                const unsigned *__restrict__ maxbits, double quality, unsigned long long *__restrict__ cand,
                size_t cand_cap, int *__restrict__ cand_count)
{
    int slot = blockIdx.z;
    unsigned mb = maxbits[slot];
    if (mb == 0) return;
    int x = blockIdx.x * blockDim.x + threadIdx.x + 1;
    int y = blockIdx.y * blockDim.y + threadIdx.y + 1;
    if (x >= w - 1 || y >= h - 1) return;
    float thr = (float)((double)ord2f(mb) * quality);
    const float *e = eig + (size_t)slot * w * h;
    float v = e[(size_t)y * w + x];
    if (!(v > thr) || v == 0.f) return;
    if (mask && mask[(size_t)slot * w * h + (size_t)y * w + x] == 0) return;
#pragma unroll
    for (int dy = -1; dy <= 1; dy++)
#pragma unroll
        for (int dx = -1; dx <= 1; dx++) {
            if (dx == 0 && dy == 0) continue;
            float n = e[(size_t)(y + dy) * w + x + dx];
            n = (n > thr) ? n : 0.f;
            if (n > v) return;
        }
    int pos = atomicAdd(cand_count + slot, 1);
    if ((size_t)pos < cand_cap)
        cand[(size_t)slot * cand_cap + pos] = ((unsigned long long)f2ord(v) << 32) | (unsigned)(y * w + x);
}

// ------------------------------------------------------------------------------------------
#define GR_T 256
#define GR_SMEM_KEYS 8192
__global__ void __launch_bounds__(GR_T)
k_corner_greedy(unsigned long long *__restrict__ cand, size_t cand_cap, const int *__restrict__ cand_count,
                int w, int h, int max_corners, double min_distance, float *__restrict__ out_xy,
                float *__restrict__ out_resp, int32_t *__restrict__ out_n, int *__restrict__ overflow)
{
    extern __shared__ unsigned long long sm[];
    __shared__ int s_first;
    int slot = blockIdx.x, tid = threadIdx.x, lane = tid & 31;
    int n = cand_count[slot];
    if ((size_t)n > cand_cap) { if (tid == 0) { *overflow = 1; out_n[slot] = 0; } return; }
    unsigned long long *g = cand + (size_t)slot * cand_cap;
    int npad = 1;
    while (npad < n) npad <<= 1;
    bool in_smem = npad <= GR_SMEM_KEYS;
    unsigned long long *keys = in_smem ? sm : g;
    unsigned *bitmap = reinterpret_cast<unsigned *>(sm + GR_SMEM_KEYS);
    int nwords = (w * h + 31) / 32;
    if (in_smem) for (int i = tid; i < npad; i += GR_T) keys[i] = (i < n) ? g[i] : 0ull;
    else for (int i = n + tid; i < npad; i += GR_T) keys[i] = 0ull;   // cand_cap is a power of two >= npad
    for (int i = tid; i < nwords; i += GR_T) bitmap[i] = 0u;
    __syncthreads();
    // bitonic sort, descending
    for (int k = 2; k <= npad; k <<= 1)
        for (int j = k >> 1; j > 0; j >>= 1) {
            for (int i = tid; i < npad; i += GR_T) {
                int p = i ^ j;
                if (p > i) {
                    unsigned long long a = keys[i], b = keys[p];
                    bool desc = (i & k) == 0;
                    if ((a < b) == desc) { keys[i] = b; keys[p] = a; }
                }
            }
            __syncthreads();
        }
    float *oxy = out_xy + (size_t)slot * 2 * max_corners;
    float *orp = out_resp + (size_t)slot * max_corners;
    if (min_distance < 1.0) {
        int m = min(n, max_corners);
        for (int i = tid; i < m; i += GR_T) {
            unsigned idx = (unsigned)keys[i];
            oxy[2 * i] = (float)(idx % w); oxy[2 * i + 1] = (float)(idx / w);
            orp[i] = ord2f((unsigned)(keys[i] >> 32));
        }
        if (tid == 0) out_n[slot] = m;
        return;
    }
    // greedy: a pixel is "blocked" iff an accepted corner lies in one of the 3x3 neighbouring grid
    // cells (cell = cvRound(minDistance)) at squared distance < minDistance^2 — OpenCV's exact test.
    int cell = __double2int_rn(min_distance);
    double md2 = min_distance * min_distance;
    int R = (int)ceil(min_distance);
    int pos = 0, nacc = 0;
    while (pos < n && nacc < max_corners) {
        if (tid == 0) s_first = INT_MAX;
        __syncthreads();
        int ci = pos + tid;
        bool ok = false;
        if (ci < n) { unsigned idx = (unsigned)keys[ci]; ok = !((bitmap[idx >> 5] >> (idx & 31)) & 1u); }
        unsigned bal = __ballot_sync(0xffffffffu, ok);
        if (lane == 0 && bal) atomicMin(&s_first, pos + (tid & ~31) + __ffs(bal) - 1);
        __syncthreads();
        int f = s_first;
        if (f == INT_MAX) { pos += GR_T; __syncthreads(); continue; }
        unsigned long long kf = keys[f];
        int qx = (int)((unsigned)kf % w), qy = (int)((unsigned)kf / w);
        if (tid == 0) {
            oxy[2 * nacc] = (float)qx; oxy[2 * nacc + 1] = (float)qy;
            orp[nacc] = ord2f((unsigned)(kf >> 32));
        }
        int cx = qx / cell, cy = qy / cell;
        int bx0 = max(max(qx - R, 0), (cx - 1) * cell), bx1 = min(min(qx + R, w - 1), (cx + 2) * cell - 1);
        int by0 = max(max(qy - R, 0), (cy - 1) * cell), by1 = min(min(qy + R, h - 1), (cy + 2) * cell - 1);
        int bw = bx1 - bx0 + 1, bh = by1 - by0 + 1;
        for (int i = tid; i < bw * bh; i += GR_T) {
            int x = bx0 + i % bw, y = by0 + i / bw;
            int dx = x - qx, dy = y - qy;
            if ((double)(dx * dx + dy * dy) < md2) { int id = y * w + x; atomicOr(&bitmap[id >> 5], 1u << (id & 31)); }
        }
        nacc++;
        pos = f + 1;
        __syncthreads();
    }
    if (tid == 0) out_n[slot] = nacc;
}

// ------------------------------------------------------------------------------------------
static size_t next_pow2(size_t v) { size_t p = 1; while (p < v) p <<= 1; return p; }

int svs_i_gftt(svs_ctx *c, const uint8_t *img, int w, int h, int stride, size_t img_pitch, int n_img,
               const int32_t *img_ids, const uint8_t *mask, const int32_t *occ_off, const float *occ_xy,
               int n_occ_total, int max_corners, double quality, double min_distance, int granule,
               float *out_xy, float *out_resp, int32_t *out_n, float *eig_out, const CUtensorMap *tm)
{
    if (n_img <= 0) return SVS_OK;
    if (w < 3 || h < 3) SVS_FAIL(c, SVS_ERR_ARG, "gftt: image must be at least 3x3");
    if (max_corners <= 0 && !eig_out) SVS_FAIL(c, SVS_ERR_ARG, "gftt: max_corners must be > 0");
    size_t P = (size_t)w * h;
    float *eig = eig_out;
    if (!eig) { SVS_CUDA(c, c->d_tmp.reserve(P * n_img * sizeof(float))); eig = c->d_tmp.as<float>(); }
    // per-image state: [maxbits u32][count i32] x n_img, then overflow flag
    SVS_CUDA(c, c->d_tmp2.reserve((size_t)n_img * 8 + 16));
    unsigned *maxbits = c->d_tmp2.as<unsigned>();
    int *count = reinterpret_cast<int *>(maxbits + n_img);
    int *overflow = count + n_img;
    SVS_CUDA(c, cudaMemsetAsync(maxbits, 0, (size_t)n_img * 8 + 4, c->stream));
    uint8_t *mask_dev = const_cast<uint8_t *>(mask);
    if (!mask && occ_xy && n_occ_total > 0) {
        SVS_CUDA(c, c->d_tmp3.reserve(P * n_img));
        mask_dev = c->d_tmp3.as<uint8_t>();
        SVS_CUDA(c, cudaMemsetAsync(mask_dev, 255, P * n_img, c->stream));
        SVS_KERNEL(c, KID_MASK, k_mask_boxes<<<n_occ_total, 128, 0, c->stream>>>(mask_dev, w, h, occ_off, n_img, occ_xy));
    }
    const double s = 1.0 / (4.0 * 3.0 * 255.0);
    float ks = (float)s, k2 = (float)(2.0 * s);
    int body = granule > 0 ? (w / granule) * granule : 0;
    int n_strips = (w + CR_OUT - 1) / CR_OUT;
    dim3 grd((n_strips + 3) / 4, n_img);
    CUtensorMap tm_local;
    const CUtensorMap *tmap = tm;
    if (!tmap && w >= 8 && h >= 8 && (stride & 15) == 0 && (img_pitch & 15) == 0 && (reinterpret_cast<uintptr_t>(img) & 15) == 0 && !img_ids &&
        svs_i_tmap_u8_3d(&tm_local, img, w, h, n_img, (size_t)stride, n_img > 1 ? img_pitch : align_up((size_t)stride * h, 16), SVS_CR_BOX_W, SVS_CR_BOX_H) == 0)
        tmap = &tm_local;
    // A big batch already fills the machine with one warp per 28-column strip (the sequential march), and the tiled kernel's
    // exactness check does fire in practice (a residual gradient of ~1e-10 next to a strong one puts more than 53 bits between
    // the products), so the tiled + fallback pair only pays when there are too few strips to occupy the SMs: small batches,
    // single streams (latency), the full-resolution configuration.
    const bool enough_strips = (long long)grd.x * grd.y >= 2LL * c->sm_count && !getenv("SVS_GFTT_TILED");
    if (tmap && w >= 8 && h >= 8 && !enough_strips) {
        // tiled + TMA-staged kernel with the exactness check; the sequential kernel re-does the images it flags
        const int tx = (w + CT_W - 1) / CT_W, ty = (h + CT_H - 1) / CT_H;
        const long long total = (long long)tx * ty * n_img;
        SVS_CUDA(c, c->d_tmp7.reserve((size_t)total * 4 + (size_t)n_img * 4 + 16));
        unsigned *tile_max = c->d_tmp7.as<unsigned>();
        int *inexact = reinterpret_cast<int *>(tile_max + total);
        SVS_CUDA(c, cudaMemsetAsync(inexact, 0, (size_t)n_img * 4, c->stream));
        const int grid = (int)std::min<long long>(total, (long long)c->sm_count * 4);
        SVS_KERNEL(c, KID_CORNER_RESPONSE, k_corner_response_tma<<<grid, 256, 0, c->stream>>>(*tmap, img_ids, n_img, w, h, mask_dev, eig, tile_max, inexact,
                                                                                               ks, k2, body, tx, ty));
        SVS_KERNEL(c, KID_CORNER_RESPONSE, k_corner_response<<<grd, 128, 0, c->stream>>>(img, img_pitch, img_ids, w, h, stride, mask_dev, eig, maxbits, ks,
                                                                                         k2, body, inexact));
        SVS_KERNEL(c, KID_MISC, k_corner_tile_max<<<n_img, 32, 0, c->stream>>>(tile_max, tx * ty, inexact, maxbits));
    } else {
        SVS_KERNEL(c, KID_CORNER_RESPONSE, k_corner_response<<<grd, 128, 0, c->stream>>>(img, img_pitch, img_ids, w, h, stride, mask_dev, eig, maxbits, ks,
                                                                                         k2, body, nullptr));
    }
    if (max_corners <= 0) return SVS_OK;
    size_t cap = next_pow2(P / 4 + 1);
    SVS_CUDA(c, c->d_tmp4.reserve(cap * n_img * sizeof(unsigned long long)));
    unsigned long long *cand = c->d_tmp4.as<unsigned long long>();
    {
        dim3 blk(64, 4);
        dim3 g2((w - 2 + 63) / 64, (h - 2 + 3) / 4, n_img);
        SVS_KERNEL(c, KID_CORNER_SELECT, k_corner_select<<<g2, blk, 0, c->stream>>>(eig, mask_dev, w, h, maxbits, quality, cand, cap, count));
    }
    {
        size_t smem = GR_SMEM_KEYS * sizeof(unsigned long long) + ((P + 31) / 32) * 4;
        if (smem > 200 * 1024) SVS_FAIL(c, SVS_ERR_CAPACITY, "gftt: image too large for the shared-memory blocked bitmap");
        SVS_CUDA(c, svs_i_opt_in_smem(c, reinterpret_cast<const void *>(k_corner_greedy)));
        SVS_KERNEL(c, KID_CORNER_GREEDY, k_corner_greedy<<<n_img, GR_T, smem, c->stream>>>(cand, cap, count, w, h, max_corners, min_distance, out_xy,
                                                          out_resp, out_n, overflow));
    }
    // overflow is checked by the caller after it synchronises (svs_i_gftt_overflow)
    return SVS_OK;
}

int svs_i_gftt_overflow(svs_ctx *c, int n_img, int *flag_host)
{
    unsigned *maxbits = c->d_tmp2.as<unsigned>();
    int *overflow = reinterpret_cast<int *>(maxbits + n_img) + n_img;
    SVS_CUDA(c, cudaMemcpyAsync(flag_host, overflow, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
    return SVS_OK;
}
