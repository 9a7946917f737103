// Image plumbing kernels: half-resolution nearest resize (row a0) and the Gaussian pyramid
// used by pyramidal LK (rows a2/a3).  Integer arithmetic, bit-exact.
//
//   k_half_nearest   <- cv::resize(.., 0.5, 0.5, INTER_NEAREST)        reference src/dataset.cpp:128-129
//   k_pyr_down       <- cv::pyrDown inside cv::calcOpticalFlowPyrLK    reference src/frontend.cpp:105, :353
//
// Both are HBM-streaming stencils over a batch of images (blockIdx.z = image).
#include "svs_internal.h"

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
                               size_t dst_img_pitch, int vec_ok, int rows_decimated)
{
    int img = blockIdx.z;
    if (src_ptrs) { src = src_ptrs[img]; img_stride = 0; }
    int y = blockIdx.y * blockDim.y + threadIdx.y;
    int x4 = (blockIdx.x * blockDim.x + threadIdx.x) * 4;
    if (y >= dh || x4 >= dw) return;
    int sy = rows_decimated ? y : min(2 * y, h - 1);   // decimated staging already holds rows 0,2,4,...
    const uint8_t *srow = src + (size_t)img * img_stride + (size_t)sy * row_stride;
    uint8_t *drow = dst + (size_t)blockIdx.z * dst_img_pitch + (size_t)y * dst_stride;
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
                       const uint8_t *const *src_ptrs_dev, int ptrs_aligned4, int rows_decimated)
{
    if (n <= 0) return SVS_OK;
    int vec_ok = ((reinterpret_cast<uintptr_t>(src) | ((rows_decimated ? 1 : 2) * row_stride) | img_stride) & 3) == 0 &&
                 ((reinterpret_cast<uintptr_t>(dst) | (size_t)dst_stride | dst_img_pitch) & 3) == 0;
    if (src_ptrs_dev) vec_ok = ptrs_aligned4 && ((2 * row_stride) & 3) == 0 &&
                               ((reinterpret_cast<uintptr_t>(dst) | (size_t)dst_stride | dst_img_pitch) & 3) == 0;
    dim3 blk(32, 8);
    dim3 grd((dw + 4 * 32 - 1) / (4 * 32), (dh + 7) / 8, n);
    SVS_KERNEL(c, KID_HALF, k_half_nearest<<<grd, blk, 0, c->stream>>>(src, src_ptrs_dev, w, h, row_stride, img_stride, dst, dw, dh, dst_stride,
                                               dst_img_pitch, vec_ok, rows_decimated));
    return SVS_OK;
}

__global__ void k_copy2d(const uint8_t *__restrict__ src, const uint8_t *const *__restrict__ src_ptrs, int w, int h,
                         size_t row_stride, size_t img_stride, uint8_t *__restrict__ dst, int dst_stride, size_t dst_img_pitch)
{
    int img = blockIdx.z;
    if (src_ptrs) { src = src_ptrs[img]; img_stride = 0; }
    int y = blockIdx.y * blockDim.y + threadIdx.y;
    int x = blockIdx.x * blockDim.x + threadIdx.x;
    if (y >= h || x >= w) return;
    dst[(size_t)blockIdx.z * dst_img_pitch + (size_t)y * dst_stride + x] =
        __ldg(src + (size_t)img * img_stride + (size_t)y * row_stride + x);
}

int svs_i_copy_level0(svs_ctx *c, const uint8_t *src, int w, int h, size_t row_stride, size_t img_stride, int n,
                      const PyrDesc &d, const uint8_t *const *src_ptrs_dev)
{
    if (n <= 0) return SVS_OK;
    dim3 blk(64, 4);
    dim3 grd((w + 63) / 64, (h + 3) / 4, n);
    SVS_KERNEL(c, KID_COPY0, k_copy2d<<<grd, blk, 0, c->stream>>>(src, src_ptrs_dev, w, h, row_stride, img_stride, d.base + d.off[0], d.stride[0],
                                         d.img_pitch));
    return SVS_OK;
}

// ---------------------------------------------------------------------------------------------
// pyrDown: separable [1 4 6 4 1], BORDER_REFLECT_101, even rows/cols, (sum + 128) >> 8.
// One CTA = 32x8 output pixels; the 67x19 input tile is staged in shared memory (each input byte is
// read from HBM/L2 once per CTA), then a horizontal pass into shared memory and a vertical pass.
__device__ __forceinline__ int refl101(int i, int n)
{
    if (n == 1) return 0;
    while (i < 0 || i >= n) i = (i < 0) ? -i : 2 * (n - 1) - i;
    return i;
}

#define PD_TX 32
#define PD_TY 8
__global__ void __launch_bounds__(PD_TX *PD_TY)
k_pyr_down(const uint8_t *__restrict__ src_base, int sw, int sh, int sstride, uint8_t *__restrict__ dst_base,
           int dw, int dh, int dstride, size_t img_pitch)
{
    __shared__ uint8_t tile[2 * PD_TY + 3][2 * PD_TX + 3 + 1];
    __shared__ int hrow[2 * PD_TY + 3][PD_TX];
    const uint8_t *src = src_base + (size_t)blockIdx.z * img_pitch;
    uint8_t *dst = dst_base + (size_t)blockIdx.z * img_pitch;
    int ox0 = blockIdx.x * PD_TX, oy0 = blockIdx.y * PD_TY;
    int ix0 = 2 * ox0 - 2, iy0 = 2 * oy0 - 2;
    int tid = threadIdx.y * PD_TX + threadIdx.x;
    for (int i = tid; i < (2 * PD_TY + 3) * (2 * PD_TX + 3); i += PD_TX * PD_TY) {
        int ty = i / (2 * PD_TX + 3), tx = i % (2 * PD_TX + 3);
        tile[ty][tx] = __ldg(src + (size_t)refl101(iy0 + ty, sh) * sstride + refl101(ix0 + tx, sw));
    }
    __syncthreads();
    for (int i = tid; i < (2 * PD_TY + 3) * PD_TX; i += PD_TX * PD_TY) {
        int ty = i / PD_TX, tx = i % PD_TX;
        const uint8_t *t = &tile[ty][2 * tx];
        hrow[ty][tx] = t[0] + 4 * t[1] + 6 * t[2] + 4 * t[3] + t[4];
    }
    __syncthreads();
    int ox = ox0 + threadIdx.x, oy = oy0 + threadIdx.y;
    if (ox < dw && oy < dh) {
        int ty = 2 * threadIdx.y, tx = threadIdx.x;
        int s = hrow[ty][tx] + 4 * hrow[ty + 1][tx] + 6 * hrow[ty + 2][tx] + 4 * hrow[ty + 3][tx] + hrow[ty + 4][tx];
        dst[(size_t)oy * dstride + ox] = (uint8_t)((s + 128) >> 8);
    }
}

int svs_i_build_pyramid(svs_ctx *c, const PyrDesc &d, int n_images)
{
    if (n_images <= 0) return SVS_OK;
    for (int l = 1; l < d.nlev; l++) {
        dim3 blk(PD_TX, PD_TY);
        dim3 grd((d.w[l] + PD_TX - 1) / PD_TX, (d.h[l] + PD_TY - 1) / PD_TY, n_images);
        SVS_KERNEL(c, KID_PYRDOWN, k_pyr_down<<<grd, blk, 0, c->stream>>>(d.base + d.off[l - 1], d.w[l - 1], d.h[l - 1], d.stride[l - 1],
                                               d.base + d.off[l], d.w[l], d.h[l], d.stride[l], d.img_pitch));
    }
    return SVS_OK;
}
