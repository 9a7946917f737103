// Image plumbing kernels: half-resolution nearest resize (row a0) and the Gaussian pyramid
// used by pyramidal LK (rows a2/a3).  Integer arithmetic, bit-exact.
//
//   k_half_nearest   <- cv::resize(.., 0.5, 0.5, INTER_NEAREST)        reference src/dataset.cpp:128-129
//   k_pyr_down       <- cv::pyrDown inside cv::calcOpticalFlowPyrLK    reference src/frontend.cpp:105, :353
//
// Both are HBM-streaming stencils over a batch of images (blockIdx.z = image).
#include "svs_internal.h"
#include "tma.cuh"
#include <algorithm>

int svs_i_make_pyr_desc(PyrDesc *d, int w, int h, int win, int max_level, size_t *bytes_per_image)
{
    if (max_level > SVS_MAX_LEVELS - 1) max_level = SVS_MAX_LEVELS - 1;
    d->base = nullptr;
    size_t off = 0;
    int lw = w, lh = h, n = 0;
    for (int l = 0; l <= max_level; l++) {
        if (l > 0) {
            lw = (lw + 1) / 2; lh = (lh + 1) / 2;
            // buildOpticalFlowPyramid stops before a level not larger than the window
            if (lw <= win || lh <= win) break;
        }
        d->w[l] = lw; d->h[l] = lh;
        d->stride[l] = (int)align_up((size_t)lw, 16);
        d->off[l] = off;
        off += align_up((size_t)d->stride[l] * lh, 256);
        n = l + 1;
    }
    d->nlev = n;
    d->img_pitch = off;
    if (bytes_per_image) *bytes_per_image = off;
    return SVS_OK;
}

// ---------------------------------------------------------------------------------------------
// dst[y][x] = src[min(2y,h-1)][min(2x,w-1)]; each thread produces 4 output pixels.
__global__ void k_half_nearest(const uint8_t *__restrict__ src, const uint8_t *const *__restrict__ src_ptrs, int w, int h,
                               size_t row_stride, size_t img_stride, uint8_t *__restrict__ dst, int dw, int dh, int dst_stride,
                               size_t dst_img_pitch, int vec_ok, int rows_decimated, const int32_t *__restrict__ dst_ids)
{
    int img = blockIdx.z;
    const int dimg = dst_ids ? dst_ids[img] : img;      // destination slot (stream) of source image `img`
    if (src_ptrs) { src = src_ptrs[img]; img_stride = 0; }
    int y = blockIdx.y * blockDim.y + threadIdx.y;
    int x4 = (blockIdx.x * blockDim.x + threadIdx.x) * 4;
    if (y >= dh || x4 >= dw) return;
    int sy = rows_decimated ? y : min(2 * y, h - 1);   // decimated staging already holds rows 0,2,4,...
    const uint8_t *srow = src + (size_t)img * img_stride + (size_t)sy * row_stride;
    uint8_t *drow = dst + (size_t)dimg * dst_img_pitch + (size_t)y * dst_stride;
    if (vec_ok && x4 + 4 <= dw && 2 * x4 + 8 <= w) {
        const uint32_t *s32 = reinterpret_cast<const uint32_t *>(srow + 2 * x4);
        uint32_t a = __ldg(s32), b = __ldg(s32 + 1);
        uint32_t o = (a & 0xFF) | ((a >> 8) & 0xFF00) | ((b & 0xFF) << 16) | ((b << 8) & 0xFF000000u);
        *reinterpret_cast<uint32_t *>(drow + x4) = o;
    } else {
        for (int k = 0; k < 4 && x4 + k < dw; k++) drow[x4 + k] = __ldg(srow + min(2 * (x4 + k), w - 1));
    }
}

int svs_i_half_nearest(svs_ctx *c, const uint8_t *src, int w, int h, size_t row_stride, size_t img_stride,
                       int n, uint8_t *dst, int dw, int dh, int dst_stride, size_t dst_img_pitch,
                       const uint8_t *const *src_ptrs_dev, int ptrs_aligned4, int rows_decimated, const int32_t *dst_ids_dev)
{
    if (n <= 0) return SVS_OK;
    int vec_ok = ((reinterpret_cast<uintptr_t>(src) | ((rows_decimated ? 1 : 2) * row_stride) | img_stride) & 3) == 0 &&
                 ((reinterpret_cast<uintptr_t>(dst) | (size_t)dst_stride | dst_img_pitch) & 3) == 0;
    if (src_ptrs_dev) vec_ok = ptrs_aligned4 && ((2 * row_stride) & 3) == 0 &&
                               ((reinterpret_cast<uintptr_t>(dst) | (size_t)dst_stride | dst_img_pitch) & 3) == 0;
    dim3 blk(32, 8);
    dim3 grd((dw + 4 * 32 - 1) / (4 * 32), (dh + 7) / 8, n);
    SVS_KERNEL(c, KID_HALF, k_half_nearest<<<grd, blk, 0, c->stream>>>(src, src_ptrs_dev, w, h, row_stride, img_stride, dst, dw, dh, dst_stride,
                                               dst_img_pitch, vec_ok, rows_decimated, dst_ids_dev));
    return SVS_OK;
}

// ---------------------------------------------------------------------------------------------
// Zero-copy ingest (frames in pinned, device-addressable HOST memory): the same resize, reading the even source rows
// straight over PCIe.  A PCIe read has ~2 us latency, so what this kernel needs is bytes in flight (~100 KB at the
// link's ~45 GB/s), NOT SM residency: a normal one-thread-per-output grid fills every CTA slot of the device with warps
// parked on PCIe loads and keeps the other contexts' kernels (LK, pose-LM, BA of other stream groups) off the SMs for
// the whole transfer.  So: a SMALL persistent grid, each thread keeps ZC_U x 2 independent 4-byte loads in flight and
// walks (image, row, 4-pixel chunk) items with a grid stride; both eyes in one launch (src_ptrs = [left.. | right..]).
// The grid is sized so that ALL live contexts together keep ~0.5 MB in flight (svs_i_zc_grid): with double-buffered
// ingest every context has such a kernel resident most of the time.
#define ZC_U 8
__global__ void __launch_bounds__(256)
k_half_nearest_zc(const uint8_t *const *__restrict__ src_ptrs, int n_img_per_eye, int w, int h, size_t row_stride,
                  uint8_t *__restrict__ dstL, uint8_t *__restrict__ dstR /* null: one eye only */, int dw, int dh, int dst_stride,
                  size_t dst_img_pitch, int vec_ok, const int32_t *__restrict__ dst_ids)
{
    const int chunks = (dw + 3) >> 2;                    // 4 output pixels per item
    const long long per_img = (long long)chunks * dh;
    const long long total = per_img * (dstR ? 2 : 1) * n_img_per_eye;
    const long long nthr = (long long)gridDim.x * blockDim.x;
    for (long long base = (long long)blockIdx.x * blockDim.x + threadIdx.x; base < total; base += nthr * ZC_U) {
        uint32_t a[ZC_U], b[ZC_U];
        uint8_t *drow[ZC_U];
        const uint8_t *srow[ZC_U];
        int x4[ZC_U];
        bool fast[ZC_U];
#pragma unroll
        for (int u = 0; u < ZC_U; u++) {
            long long it = base + (long long)u * nthr;
            fast[u] = false; drow[u] = nullptr; srow[u] = nullptr; x4[u] = 0; a[u] = b[u] = 0;
            if (it >= total) continue;
            int img = (int)(it / per_img);
            int r = (int)(it - (long long)img * per_img);
            int y = r / chunks;
            x4[u] = (r - y * chunks) * 4;
            srow[u] = src_ptrs[img] + (size_t)min(2 * y, h - 1) * row_stride;
            const int slot = img < n_img_per_eye ? img : img - n_img_per_eye;
            const int dimg = dst_ids ? dst_ids[slot] : slot;
            uint8_t *dst = (img < n_img_per_eye ? dstL : dstR) + (size_t)dimg * dst_img_pitch;
            drow[u] = dst + (size_t)y * dst_stride;
            fast[u] = vec_ok && x4[u] + 4 <= dw && 2 * x4[u] + 8 <= w;
            if (fast[u]) {
                const uint32_t *s32 = reinterpret_cast<const uint32_t *>(srow[u] + 2 * x4[u]);
                a[u] = __ldg(s32); b[u] = __ldg(s32 + 1);
            }
        }
#pragma unroll
        for (int u = 0; u < ZC_U; u++) {
            if (!drow[u]) continue;
            if (fast[u]) {
                uint32_t o = (a[u] & 0xFF) | ((a[u] >> 8) & 0xFF00) | ((b[u] & 0xFF) << 16) | ((b[u] << 8) & 0xFF000000u);
                *reinterpret_cast<uint32_t *>(drow[u] + x4[u]) = o;
            } else {
                for (int k = 0; k < 4 && x4[u] + k < dw; k++) drow[u][x4[u] + k] = __ldg(srow[u] + min(2 * (x4[u] + k), w - 1));
            }
        }
    }
}

int svs_i_half_nearest_zc(svs_ctx *c, const uint8_t *const *src_ptrs_dev, int n_per_eye, int w, int h, size_t row_stride,
                          uint8_t *dstL, uint8_t *dstR, int dw, int dh, int dst_stride, size_t dst_img_pitch, int ptrs_aligned4,
                          const int32_t *dst_ids_dev)
{
    if (n_per_eye <= 0) return SVS_OK;
    int vec_ok = ptrs_aligned4 && ((2 * row_stride) & 3) == 0 &&
                 ((reinterpret_cast<uintptr_t>(dstL) | reinterpret_cast<uintptr_t>(dstR) | (size_t)dst_stride | dst_img_pitch) & 3) == 0;
    long long items = (long long)((dw + 3) / 4) * dh * (dstR ? 2 : 1) * n_per_eye;
    long long want = (items + 256 * ZC_U - 1) / (256 * ZC_U);
    int grid = (int)std::min<long long>(want, svs_i_zc_grid(c));
    SVS_KERNEL(c, KID_HALF, k_half_nearest_zc<<<grid, 256, 0, c->stream>>>(src_ptrs_dev, n_per_eye, w, h, row_stride, dstL, dstR, dw, dh,
                                                                         dst_stride, dst_img_pitch, vec_ok, dst_ids_dev));
    return SVS_OK;
}

__global__ void k_copy2d(const uint8_t *__restrict__ src, const uint8_t *const *__restrict__ src_ptrs, int w, int h,
                         size_t row_stride, size_t img_stride, uint8_t *__restrict__ dst, int dst_stride, size_t dst_img_pitch,
                         const int32_t *__restrict__ dst_ids)
{
    int img = blockIdx.z;
    const int dimg = dst_ids ? dst_ids[img] : img;
    if (src_ptrs) { src = src_ptrs[img]; img_stride = 0; }
    int y = blockIdx.y * blockDim.y + threadIdx.y;
    int x = blockIdx.x * blockDim.x + threadIdx.x;
    if (y >= h || x >= w) return;
    dst[(size_t)dimg * dst_img_pitch + (size_t)y * dst_stride + x] =
        __ldg(src + (size_t)img * img_stride + (size_t)y * row_stride + x);
}

int svs_i_copy_level0(svs_ctx *c, const uint8_t *src, int w, int h, size_t row_stride, size_t img_stride, int n,
                      const PyrDesc &d, const uint8_t *const *src_ptrs_dev, const int32_t *dst_ids_dev)
{
    if (n <= 0) return SVS_OK;
    dim3 blk(64, 4);
    dim3 grd((w + 63) / 64, (h + 3) / 4, n);
    SVS_KERNEL(c, KID_COPY0, k_copy2d<<<grd, blk, 0, c->stream>>>(src, src_ptrs_dev, w, h, row_stride, img_stride, d.base + d.off[0], d.stride[0],
                                         d.img_pitch, dst_ids_dev));
    return SVS_OK;
}

// ---------------------------------------------------------------------------------------------
// pyrDown: separable [1 4 6 4 1], BORDER_REFLECT_101, even rows/cols, (sum + 128) >> 8.
// One CTA = 64x16 output pixels.  The 136-byte x 35-row input tile is staged in shared memory with aligned 32-bit loads
// (per-byte reflection only for words that straddle the image border), the horizontal pass produces two outputs per item
// packed as 16-bit halves of one word (a row sum is <= 16*255 = 4080), and the vertical pass runs on the packed halves
// (<= 16*4080 = 65280 < 2^16, so neither the weighted sum nor the +128 rounding carries across halves) and stores four
// output pixels per thread with one 32-bit store.
__device__ __forceinline__ int refl101(int i, int n)
{
    if (n == 1) return 0;
    while (i < 0 || i >= n) i = (i < 0) ? -i : 2 * (n - 1) - i;
    return i;
}

#define PD_TX 64
#define PD_TY 16
#define PD_ROWS (2 * PD_TY + 3)
#define PD_WORDS (PD_TX / 2 + 2)
__global__ void __launch_bounds__(256)
k_pyr_down(const uint8_t *__restrict__ src_base, int sw, int sh, int sstride, uint8_t *__restrict__ dst_base,
           int dw, int dh, int dstride, size_t img_pitch, const int32_t *__restrict__ img_ids)
{
    __shared__ uint32_t tile[PD_ROWS][PD_WORDS];
    __shared__ __align__(8) uint32_t hrow[PD_ROWS][PD_TX / 2];
    const int img = img_ids ? img_ids[blockIdx.z] : blockIdx.z;     // pyramid slot (stream)
    const uint8_t *src = src_base + (size_t)img * img_pitch;
    uint8_t *dst = dst_base + (size_t)img * img_pitch;
    const int ox0 = blockIdx.x * PD_TX, oy0 = blockIdx.y * PD_TY;
    const int bx0 = 2 * ox0 - 4, iy0 = 2 * oy0 - 2;      // image x of tile byte 0 (4-byte aligned), image y of tile row 0
    const int tid = threadIdx.x;
    for (int i = tid; i < PD_ROWS * PD_WORDS; i += 256) {
        int r = i / PD_WORDS, q = i - r * PD_WORDS;
        const uint8_t *row = src + (size_t)refl101(iy0 + r, sh) * sstride;
        int x = bx0 + 4 * q;
        uint32_t v;
        if (x >= 0 && x + 4 <= sw) v = __ldg(reinterpret_cast<const uint32_t *>(row + x));
        else v = (uint32_t)__ldg(row + refl101(x, sw)) | ((uint32_t)__ldg(row + refl101(x + 1, sw)) << 8) |
                 ((uint32_t)__ldg(row + refl101(x + 2, sw)) << 16) | ((uint32_t)__ldg(row + refl101(x + 3, sw)) << 24);
        tile[r][q] = v;
    }
    __syncthreads();
    // horizontal: outputs ox0+2j (tile bytes 4j+2..4j+6) and ox0+2j+1 (tile bytes 4j+4..4j+8)
    for (int i = tid; i < PD_ROWS * (PD_TX / 2); i += 256) {
        int r = i / (PD_TX / 2), j = i - r * (PD_TX / 2);
        uint32_t w0 = tile[r][j], w1 = tile[r][j + 1], w2 = tile[r][j + 2];
        uint32_t b2 = (w0 >> 16) & 0xFF, b3 = w0 >> 24, b4 = w1 & 0xFF, b5 = (w1 >> 8) & 0xFF, b6 = (w1 >> 16) & 0xFF, b7 = w1 >> 24,
                 b8 = w2 & 0xFF;
        uint32_t A = b2 + 4 * b3 + 6 * b4 + 4 * b5 + b6;
        uint32_t B = b4 + 4 * b5 + 6 * b6 + 4 * b7 + b8;
        hrow[r][j] = A | (B << 16);
    }
    __syncthreads();
    // vertical on packed halves: thread = (output row, 4 adjacent output columns)
    {
        int ty = tid >> 4, jq = tid & 15;
        int oy = oy0 + ty, ox = ox0 + 4 * jq;
        if (oy < dh && ox < dw) {
            const uint2 *h = reinterpret_cast<const uint2 *>(&hrow[2 * ty][2 * jq]);
            const int rp = (PD_TX / 2) / 2;     // row pitch in uint2
            uint2 h0 = h[0], h1 = h[rp], h2 = h[2 * rp], h3 = h[3 * rp], h4 = h[4 * rp];
            uint32_t s0 = h0.x + 4 * h1.x + 6 * h2.x + 4 * h3.x + h4.x;
            uint32_t s1 = h0.y + 4 * h1.y + 6 * h2.y + 4 * h3.y + h4.y;
            uint32_t r0 = ((s0 + 0x00800080u) >> 8) & 0x00FF00FFu, r1 = ((s1 + 0x00800080u) >> 8) & 0x00FF00FFu;
            uint32_t out = (r0 & 0xFF) | ((r0 >> 16) << 8) | ((r1 & 0xFF) << 16) | ((r1 >> 16) << 24);
            uint8_t *d = dst + (size_t)oy * dstride + ox;
            if (ox + 4 <= dw) *reinterpret_cast<uint32_t *>(d) = out;
            else for (int k = 0; k < 4 && ox + k < dw; k++) d[k] = (uint8_t)(out >> (8 * k));
        }
    }
}

// ---------------------------------------------------------------------------------------------
// pyrDown with TMA staging (frame sets): a PERSISTENT grid walks the (image, tile) list of one level; the 160-byte x 35-row
// source box of the NEXT tile is in flight (cp.async.bulk.tensor.3d -> UTMALDG, completion on an mbarrier) while the current
// one is filtered — the staging loop, its address arithmetic and its per-word border tests are gone from the instruction
// stream.  TMA fills out-of-image bytes with zeros; the few border tiles patch their reflect-101 rows / columns in shared
// memory (the reflected source pixel is always inside the same tile).  Arithmetic identical to k_pyr_down (bit-exact).
#define PD_TP SVS_PD_BOX_W                       // tile row pitch in bytes = TMA box width; tile byte c is image x = 2*ox0 - 16 + c
#define PD_TBYTES (PD_ROWS * PD_TP)              // 5600
__global__ void __launch_bounds__(256)
k_pyr_down_tma(const __grid_constant__ CUtensorMap tm, int sw, int sh, uint8_t *__restrict__ dst_base, int dw, int dh, int dstride,
               size_t img_pitch, const int32_t *__restrict__ img_ids, int n_img, int tiles_x, int tiles_y)
{
    __shared__ __align__(128) uint8_t tile_s[2][5632];
    __shared__ __align__(8) uint32_t hrow[PD_ROWS][PD_TX / 2];
    __shared__ __align__(8) uint64_t bar[2];
    const int tid = threadIdx.x;
    const int per_img = tiles_x * tiles_y, total = n_img * per_img;
    if (tid == 0) { tma::mbar_init(&bar[0], 1); tma::mbar_init(&bar[1], 1); tma::fence_init(); }
    __syncthreads();
    auto issue = [&](int t, int buf) {
        const int n = t / per_img, r = t - n * per_img, ty = r / tiles_x, tx = r - ty * tiles_x;
        const int z = img_ids ? img_ids[n] : n;
        tma::mbar_expect_tx(&bar[buf], PD_TBYTES);
        tma::load_3d(tile_s[buf], &tm, &bar[buf], 2 * tx * PD_TX - 16, 2 * ty * PD_TY - 2, z);
    };
    int t = blockIdx.x;
    if (tid == 0 && t < total) issue(t, 0);
    for (int it = 0; t < total; t += gridDim.x, it++) {
        const int buf = it & 1, tn = t + gridDim.x;
        if (tid == 0 && tn < total) { tma::fence_proxy_async(); issue(tn, buf ^ 1); }
        const int n = t / per_img, r0 = t - n * per_img, ty0 = r0 / tiles_x, tx0 = r0 - ty0 * tiles_x;
        const int img = img_ids ? img_ids[n] : n;
        uint8_t *dst = dst_base + (size_t)img * img_pitch;
        const int ox0 = tx0 * PD_TX, oy0 = ty0 * PD_TY, bx0 = 2 * ox0 - 16, iy0 = 2 * oy0 - 2;
        uint8_t *tb = tile_s[buf];
        tma::mbar_wait(&bar[buf], (it >> 1) & 1);
        // reflect-101 patch-up of the zero-filled out-of-image bytes this tile actually reads (rows first, then columns)
        if (iy0 < 0 || iy0 + PD_ROWS > sh) {
            for (int i = tid; i < PD_ROWS * (PD_TP / 4); i += 256) {
                const int r = i / (PD_TP / 4), q = i - r * (PD_TP / 4), iy = iy0 + r;
                if (iy < 0 || iy >= sh) {
                    const int ry = refl101(iy, sh) - iy0;
                    if (ry >= 0 && ry < PD_ROWS) reinterpret_cast<uint32_t *>(tb + r * PD_TP)[q] = reinterpret_cast<const uint32_t *>(tb + ry * PD_TP)[q];
                }
            }
            __syncthreads();
        }
        if (bx0 < 0 || bx0 + PD_TP > sw) {
            for (int i = tid; i < PD_ROWS * 8; i += 256) {
                const int r = i >> 3, k = i & 7;
                const int x = (k < 4) ? k - 4 : sw + (k - 4);          // the columns a 5-tap window can reach outside the image
                const int cdst = x - bx0, csrc = refl101(x, sw) - bx0;
                if (cdst >= 0 && cdst < PD_TP && csrc >= 0 && csrc < PD_TP) tb[r * PD_TP + cdst] = tb[r * PD_TP + csrc];
            }
            __syncthreads();
        }
        for (int i = tid; i < PD_ROWS * (PD_TX / 2); i += 256) {
            const int r = i / (PD_TX / 2), j = i - r * (PD_TX / 2);
            const uint32_t *tw = reinterpret_cast<const uint32_t *>(tb + r * PD_TP);
            uint32_t w0 = tw[j + 3], w1 = tw[j + 4], w2 = tw[j + 5];      // image bytes 2*ox0 + 4j - 4 .. + 7 sit 12 bytes into the box
            uint32_t b2 = (w0 >> 16) & 0xFF, b3 = w0 >> 24, b4 = w1 & 0xFF, b5 = (w1 >> 8) & 0xFF, b6 = (w1 >> 16) & 0xFF, b7 = w1 >> 24,
                     b8 = w2 & 0xFF;
            uint32_t A = b2 + 4 * b3 + 6 * b4 + 4 * b5 + b6;
            uint32_t B = b4 + 4 * b5 + 6 * b6 + 4 * b7 + b8;
            hrow[r][j] = A | (B << 16);
        }
        __syncthreads();
        {
            int ty = tid >> 4, jq = tid & 15;
            int oy = oy0 + ty, ox = ox0 + 4 * jq;
            if (oy < dh && ox < dw) {
                const uint2 *h = reinterpret_cast<const uint2 *>(&hrow[2 * ty][2 * jq]);
                const int rp = (PD_TX / 2) / 2;
                uint2 h0 = h[0], h1 = h[rp], h2 = h[2 * rp], h3 = h[3 * rp], h4 = h[4 * rp];
                uint32_t s0 = h0.x + 4 * h1.x + 6 * h2.x + 4 * h3.x + h4.x;
                uint32_t s1 = h0.y + 4 * h1.y + 6 * h2.y + 4 * h3.y + h4.y;
                uint32_t r0 = ((s0 + 0x00800080u) >> 8) & 0x00FF00FFu, r1 = ((s1 + 0x00800080u) >> 8) & 0x00FF00FFu;
                uint32_t out = (r0 & 0xFF) | ((r0 >> 16) << 8) | ((r1 & 0xFF) << 16) | ((r1 >> 16) << 24);
                uint8_t *d = dst + (size_t)oy * dstride + ox;
                if (ox + 4 <= dw) *reinterpret_cast<uint32_t *>(d) = out;
                else for (int k = 0; k < 4 && ox + k < dw; k++) d[k] = (uint8_t)(out >> (8 * k));
            }
        }
        __syncthreads();      // every read of tile_s[buf] and hrow is done before the copy issued in the NEXT iteration lands
    }
}

int svs_i_tmap_u8_3d(CUtensorMap *m, const void *base, int w, int h, int n, size_t row_stride, size_t img_pitch, int box_w, int box_h)
{
    typedef CUresult (*EncodeFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                 const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                 CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
    static EncodeFn fn = nullptr;
    if (!fn) {
        void *p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess || !p) return -1;
        fn = reinterpret_cast<EncodeFn>(p);
    }
    if ((reinterpret_cast<uintptr_t>(base) & 15) || (row_stride & 15) || (img_pitch & 15) || (box_w & 15) || box_w > 256 || box_h > 256) return -2;
    cuuint64_t gdim[3] = {(cuuint64_t)w, (cuuint64_t)h, (cuuint64_t)n};
    cuuint64_t gstr[2] = {(cuuint64_t)row_stride, (cuuint64_t)img_pitch};
    cuuint32_t box[3] = {(cuuint32_t)box_w, (cuuint32_t)box_h, 1}, estr[3] = {1, 1, 1};
    CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_UINT8, 3, const_cast<void *>(base), gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                    CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS ? 0 : -3;
}

int svs_i_build_pyramid(svs_ctx *c, const PyrDesc &d, int n_images, const int32_t *img_ids_dev, const CUtensorMap *level_maps)
{
    if (n_images <= 0) return SVS_OK;
    for (int l = 1; l < d.nlev; l++) {
        if (level_maps && d.w[l - 1] >= 16 && d.h[l - 1] >= 16) {     // TMA-staged persistent kernel (frame sets)
            const int tx = (d.w[l] + PD_TX - 1) / PD_TX, ty = (d.h[l] + PD_TY - 1) / PD_TY;
            const long long total = (long long)tx * ty * n_images;
            const int grid = (int)std::min<long long>(total, (long long)c->sm_count * 6);
            SVS_KERNEL(c, KID_PYRDOWN, k_pyr_down_tma<<<grid, 256, 0, c->stream>>>(level_maps[l - 1], d.w[l - 1], d.h[l - 1], d.base + d.off[l], d.w[l],
                                                                                     d.h[l], d.stride[l], d.img_pitch, img_ids_dev, n_images, tx, ty));
            continue;
        }
        dim3 blk(256);
        dim3 grd((d.w[l] + PD_TX - 1) / PD_TX, (d.h[l] + PD_TY - 1) / PD_TY, n_images);
        SVS_KERNEL(c, KID_PYRDOWN, k_pyr_down<<<grd, blk, 0, c->stream>>>(d.base + d.off[l - 1], d.w[l - 1], d.h[l - 1], d.stride[l - 1],
                                               d.base + d.off[l], d.w[l], d.h[l], d.stride[l], d.img_pitch, img_ids_dev));
    }
    return SVS_OK;
}
