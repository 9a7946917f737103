// Dense stereo path (row a10): cv::StereoBM(ndisp, block)::compute and the disparity -> depth -> world
// back-projection of reference src/dense_reconstruction.cpp:104-173.  Integer arithmetic, bit-exact
// against cv2.StereoBM.
//
//   k_bgr2gray      cv::cvtColor(BGR2GRAY) 8-bit fixed point (shift 15)                      :111-113
//   k_bm_prefilter  PREFILTER_XSOBEL, cap 31, both eyes
//   k_bm_sad        one CTA = 32 output columns x a strip of rows, all disparities:
//                   A) per (column, 16 disparities) vertical running sums of |L'-R'| in registers
//                   B) per (disparity, 8 columns) horizontal running window sums -> SAD tile in shared memory
//                   C) 16 lanes per pixel: arg-min (first minimum in OpenCV's search order), texture and
//                      uniqueness tests, sub-pixel interpolation -> int16 disparity * 16        :114
//   k_backproject   disp/16 -> depth (f32) -> Camera::pixel2world (f64), x-outer / y-inner order :116-173
// This stage is integer-ALU / shared-memory bound, not HBM bound (4 B of compulsory traffic per pixel).
#include "svs_internal.h"
#include "geom_dev.cuh"
#include <cstring>

__global__ void k_bgr2gray(const uint8_t *__restrict__ bgr, uint8_t *__restrict__ gray, size_t n)
{
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    int b = bgr[3 * i], g = bgr[3 * i + 1], r = bgr[3 * i + 2];
    gray[i] = (uint8_t)((b * 3735 + g * 19235 + r * 9798 + (1 << 14)) >> 15);
}

__global__ void k_bm_prefilter(const uint8_t *__restrict__ src, int w, int h, int stride, size_t img_stride,
                               uint8_t *__restrict__ dst, int cap)
{
    int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y * blockDim.y + threadIdx.y;
    if (x >= w || y >= h) return;
    const uint8_t *s = src + (size_t)blockIdx.z * img_stride;
    uint8_t *d = dst + (size_t)blockIdx.z * w * h;
    int v = cap;
    if (x > 0 && x < w - 1 && !((h & 1) && y == h - 1)) {
        int ym = (y > 0) ? y - 1 : (h > 1 ? 1 : 0), yp = (y < h - 1) ? y + 1 : (h > 1 ? h - 2 : 0);
        const uint8_t *r0 = s + (size_t)ym * stride, *r1 = s + (size_t)y * stride, *r2 = s + (size_t)yp * stride;
        int g = (r0[x + 1] - r0[x - 1]) + 2 * (r1[x + 1] - r1[x - 1]) + (r2[x + 1] - r2[x - 1]);
        v = min(max(g, -cap), cap) + cap;
    }
    d[(size_t)y * w + x] = (uint8_t)v;
}

#define BM_T 512
#define BM_TW 32          // output columns per CTA
#define BM_DG 16          // disparities per phase-A thread
#define BM_RS 48          // output rows per CTA strip

__global__ void __launch_bounds__(BM_T)
k_bm_sad(const uint8_t *__restrict__ Lp, const uint8_t *__restrict__ Rp, int w, int h, int ndisp, int r, int cap, int texture,
         int uniq, int16_t *__restrict__ disp)
{
    extern __shared__ unsigned short smu[];
    const int win = 2 * r + 1, CW = BM_TW + 2 * r;           // column sums needed per row
    const int cpitch = CW | 1;                               // u16 pitch of colbuf rows (odd word-ish stride)
    unsigned short *colbuf = smu;                            // [(ndisp + 1)][cpitch]   (row ndisp = texture)
    unsigned short *sadbuf = colbuf + (size_t)(ndisp + 1) * cpitch + ((ndisp + 1) * cpitch & 1);   // [BM_TW][ndisp]
    int *texbuf = reinterpret_cast<int *>(sadbuf + (size_t)BM_TW * ndisp);                          // [BM_TW]
    const int tid = threadIdx.x;
    const int img = blockIdx.z;
    const uint8_t *L = Lp + (size_t)img * w * h, *R = Rp + (size_t)img * w * h;
    int16_t *out = disp + (size_t)img * w * h;
    const int x_lo = ndisp - 1 + r, x_hi = w - r, y_lo = r, y_hi = h - r;
    const int x0 = x_lo + blockIdx.x * BM_TW;                // first output column of this CTA
    const int ys = y_lo + blockIdx.y * BM_RS, ye = min(ys + BM_RS, y_hi);
    if (x0 >= x_hi || ys >= y_hi) return;
    const int ndg = ndisp / BM_DG;
    // phase-A role: column cx (0..CW-1) <-> image column x0 - r + cx, disparity group dg
    const int a_cx = tid % CW, a_dg = tid / CW;
    const bool a_on = a_dg < ndg && (x0 - r + a_cx) < w;
    const bool t_on = a_dg == ndg && (x0 - r + a_cx) < w;   // one extra "group" carries the texture column sums
    const int ax = x0 - r + a_cx;
    int col[BM_DG];
#pragma unroll
    for (int k = 0; k < BM_DG; k++) col[k] = 0;
    int tcol = 0;
    auto add_row = [&](int y, int sgn) {
        if (a_on) {
            int lv = L[(size_t)y * w + ax];
            const uint8_t *rr = R + (size_t)y * w + ax - a_dg * BM_DG;
#pragma unroll
            for (int k = 0; k < BM_DG; k++) col[k] += sgn * abs(lv - (int)rr[-k]);
        } else if (t_on) {
            tcol += sgn * abs((int)L[(size_t)y * w + ax] - cap);
        }
    };
    for (int y = ys - r; y < ys + r; y++) add_row(y, +1);    // first 2r rows of the first window
    for (int y = ys; y < ye; y++) {
        add_row(y + r, +1);
        if (a_on) {
#pragma unroll
            for (int k = 0; k < BM_DG; k++) colbuf[(size_t)(a_dg * BM_DG + k) * cpitch + a_cx] = (unsigned short)col[k];
        } else if (t_on) {
            colbuf[(size_t)ndisp * cpitch + a_cx] = (unsigned short)tcol;
        }
        __syncthreads();
        // ---- phase B: horizontal window sums; task = (disparity row d in 0..ndisp, 8-column group)
        for (int t = tid; t < (ndisp + 1) * (BM_TW / 8); t += BM_T) {
            int d = t % (ndisp + 1), xq = t / (ndisp + 1);
            const unsigned short *cb = colbuf + (size_t)d * cpitch + 8 * xq;
            int s = 0;
            for (int i = 0; i < win; i++) s += cb[i];
#pragma unroll
            for (int j = 0; j < 8; j++) {
                if (d < ndisp) sadbuf[(size_t)(8 * xq + j) * ndisp + d] = (unsigned short)s;
                else texbuf[8 * xq + j] = s;
                s += cb[win + j] - cb[j];
            }
        }
        __syncthreads();
        // ---- phase C: 16 lanes per pixel
        {
            int px = tid >> 4, sub = tid & 15;               // 32 pixels x 16 lanes
            int x = x0 + px;
            const unsigned short *sp = sadbuf + (size_t)px * ndisp;
            int best = 0x7fffffff, bestD = -1;
            // OpenCV scans search index d_s = ndisp-1-D ascending and keeps the first strict minimum
            // => among equal SADs the LARGEST D wins.
            for (int D = sub; D < ndisp; D += 16) {
                int v = sp[D];
                if (v < best || (v == best && D > bestD)) { best = v; bestD = D; }
            }
#pragma unroll
            for (int o = 8; o > 0; o >>= 1) {
                int ov = __shfl_xor_sync(0xffffffffu, best, o), oD = __shfl_xor_sync(0xffffffffu, bestD, o);
                if (ov < best || (ov == best && oD > bestD)) { best = ov; bestD = oD; }
            }
            int thresh = best + (best * uniq) / 100;
            int bad = 0;
            for (int D = sub; D < ndisp; D += 16) bad |= (sp[D] <= thresh && abs(D - bestD) > 1) ? 1 : 0;
#pragma unroll
            for (int o = 8; o > 0; o >>= 1) bad |= __shfl_xor_sync(0xffffffffu, bad, o);
            if (sub == 0 && x < x_hi) {
                int16_t res = (int16_t)(-16);
                if (texbuf[px] >= texture && !bad) {
                    int p = (bestD >= 1) ? sp[bestD - 1] : sp[1];
                    int n = (bestD <= ndisp - 2) ? sp[bestD + 1] : sp[ndisp - 2];
                    int den = p + n - 2 * best + abs(p - n);
                    int q = den ? ((p - n) * 256) / den : 0;
                    res = (int16_t)((bestD * 256 + q + 15) >> 4);
                }
                out[(size_t)y * w + x] = res;
            }
        }
        add_row(y - r, -1);
        __syncthreads();
    }
}

__global__ void k_fill_i16(int16_t *p, size_t n, int16_t v)
{
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) p[i] = v;
}

// ---- back-projection: one thread per image column (x outer / y inner output order)
__global__ void k_bp_count(const int16_t *__restrict__ disp, int w, int h, float fb, int *__restrict__ col_count)
{
    int x = blockIdx.x * blockDim.x + threadIdx.x;
    if (x >= w) return;
    int n = 0;
    for (int y = 0; y < h; y++) {
        float d = __fmul_rn((float)disp[(size_t)y * w + x], 0.0625f);
        float z = (d > 0.f) ? __fdiv_rn(fb, d) : 0.f;
        n += (z < 1.f) ? 0 : 1;
    }
    col_count[x] = n;
}
__global__ void k_bp_scan(int *col_count, int w, int *total)
{   // tiny exclusive scan by one thread block (w <= a few thousand)
    if (threadIdx.x == 0) {
        int run = 0;
        for (int x = 0; x < w; x++) { int c = col_count[x]; col_count[x] = run; run += c; }
        *total = run;
    }
}
__global__ void k_bp_emit(const int16_t *__restrict__ disp, const uint8_t *__restrict__ bgr, int w, int h, float fb, double fx,
                          double fy, double cx, double cy, const double *__restrict__ Tci /* cam pose inverse */,
                          const double *__restrict__ Twc /* inverse of T_cw */, const int *__restrict__ col_off,
                          float *__restrict__ xyz, uint8_t *__restrict__ rgb)
{
    int x = blockIdx.x * blockDim.x + threadIdx.x;
    if (x >= w) return;
    double A[7], B[7];
    for (int i = 0; i < 7; i++) { A[i] = Tci[i]; B[i] = Twc[i]; }
    int o = col_off[x];
    for (int y = 0; y < h; y++) {
        float d = __fmul_rn((float)disp[(size_t)y * w + x], 0.0625f);
        float z = (d > 0.f) ? __fdiv_rn(fb, d) : 0.f;
        if (z < 1.f) continue;
        double depth = (double)z;
        double pc[3] = {((double)x - cx) * depth / fx, ((double)y - cy) * depth / fy, depth}, q[3], pw[3];
        gd::se3_act(A, pc, q);
        gd::se3_act(B, q, pw);
        xyz[3 * (size_t)o] = (float)pw[0]; xyz[3 * (size_t)o + 1] = (float)pw[1]; xyz[3 * (size_t)o + 2] = (float)pw[2];
        const uint8_t *c = bgr + 3 * ((size_t)y * w + x);
        rgb[3 * (size_t)o] = c[2]; rgb[3 * (size_t)o + 1] = c[1]; rgb[3 * (size_t)o + 2] = c[0];
        o++;
    }
}

extern "C" {

int svs_bgr2gray(svs_ctx *c, const uint8_t *bgr, int w, int h, int n, uint8_t *gray)
{
    if (!c || !bgr || !gray || w < 1 || h < 1 || n < 1) return SVS_ERR_ARG;
    SVS_CUDA(c, cudaSetDevice(c->device));
    size_t px = (size_t)w * h * n;
    SVS_CUDA(c, c->d_in.reserve(px * 3));
    SVS_CUDA(c, c->d_out.reserve(px));
    SVS_CUDA(c, cudaMemcpyAsync(c->d_in.p, bgr, px * 3, cudaMemcpyHostToDevice, c->stream));
    SVS_KERNEL(c, KID_BGR2GRAY, k_bgr2gray<<<(unsigned)((px + 255) / 256), 256, 0, c->stream>>>(c->d_in.as<uint8_t>(), c->d_out.as<uint8_t>(), px));
    SVS_CUDA(c, cudaMemcpyAsync(gray, c->d_out.p, px, cudaMemcpyDeviceToHost, c->stream));
    SVS_CUDA(c, cudaStreamSynchronize(c->stream));
    return SVS_OK;
}

int svs_i_stereo_bm(svs_ctx *c, const uint8_t *l_dev, const uint8_t *r_dev, int w, int h, int stride, size_t img_stride, int n,
                    int ndisp, int block, int16_t *disp_dev)
{
    if (ndisp < 16 || ndisp % 16 || ndisp > 256) SVS_FAIL(c, SVS_ERR_ARG, "stereo_bm: ndisp must be a multiple of 16 in [16,256]");
    if (block < 5 || !(block & 1) || block > 31) SVS_FAIL(c, SVS_ERR_ARG, "stereo_bm: block must be odd in [5,31]");
    int r = block / 2, cap = 31;
    int CW = BM_TW + 2 * r;
    if (CW * (ndisp / BM_DG + 1) > BM_T) SVS_FAIL(c, SVS_ERR_CAPACITY, "stereo_bm: (32+block-1)*(ndisp/16+1) exceeds 512 threads");
    size_t px = (size_t)w * h;
    SVS_CUDA(c, c->d_tmp.reserve(2 * px * n));
    uint8_t *Lp = c->d_tmp.as<uint8_t>(), *Rp = Lp + px * n;
    dim3 pb(64, 4), pg((w + 63) / 64, (h + 3) / 4, n);
    SVS_KERNEL(c, KID_BM_PREFILTER, k_bm_prefilter<<<pg, pb, 0, c->stream>>>(l_dev, w, h, stride, img_stride, Lp, cap));
    SVS_KERNEL(c, KID_BM_PREFILTER, k_bm_prefilter<<<pg, pb, 0, c->stream>>>(r_dev, w, h, stride, img_stride, Rp, cap));
    SVS_KERNEL(c, KID_MISC, k_fill_i16<<<(unsigned)((px * n + 255) / 256), 256, 0, c->stream>>>(disp_dev, px * n, (int16_t)-16));
    int nx = w - r - (ndisp - 1 + r), ny = h - 2 * r;
    if (nx <= 0 || ny <= 0) return SVS_OK;
    int cpitch = CW | 1;
    size_t colb = (size_t)(ndisp + 1) * cpitch;
    size_t smem = (colb + (colb & 1) + (size_t)BM_TW * ndisp) * 2 + BM_TW * 4;
    SVS_CUDA(c, svs_i_opt_in_smem(c, reinterpret_cast<const void *>(k_bm_sad)));
    dim3 g((nx + BM_TW - 1) / BM_TW, (ny + BM_RS - 1) / BM_RS, n);
    SVS_KERNEL(c, KID_BM_SAD, k_bm_sad<<<g, BM_T, smem, c->stream>>>(Lp, Rp, w, h, ndisp, r, cap, 10, 15, disp_dev));
    return SVS_OK;
}

int svs_stereo_bm(svs_ctx *c, const uint8_t *left, const uint8_t *right, int w, int h, int stride, int n, size_t img_stride,
                  int ndisp, int block, int16_t *disp_out)
{
    if (!c || !left || !right || !disp_out || w < 3 || h < 3 || n < 1 || stride < w) return SVS_ERR_ARG;
    SVS_CUDA(c, cudaSetDevice(c->device));
    size_t in_b = img_stride * n, px = (size_t)w * h * n;
    SVS_CUDA(c, c->d_in.reserve(in_b));
    SVS_CUDA(c, c->d_in2.reserve(in_b));
    SVS_CUDA(c, c->d_out.reserve(px * 2));
    SVS_CUDA(c, cudaMemcpyAsync(c->d_in.p, left, in_b, cudaMemcpyHostToDevice, c->stream));
    SVS_CUDA(c, cudaMemcpyAsync(c->d_in2.p, right, in_b, cudaMemcpyHostToDevice, c->stream));
    SVS_TRY(svs_i_stereo_bm(c, c->d_in.as<uint8_t>(), c->d_in2.as<uint8_t>(), w, h, stride, img_stride, n, ndisp, block,
                            c->d_out.as<int16_t>()));
    SVS_CUDA(c, cudaMemcpyAsync(disp_out, c->d_out.p, px * 2, cudaMemcpyDeviceToHost, c->stream));
    SVS_CUDA(c, cudaStreamSynchronize(c->stream));
    return SVS_OK;
}

int svs_stereo_bm_dev(svs_ctx *c, const uint8_t *left_dev, const uint8_t *right_dev, int w, int h, int stride, int n, size_t img_stride,
                      int ndisp, int block, int16_t *disp_out_dev)
{
    if (!c || !left_dev || !right_dev || !disp_out_dev || w < 3 || h < 3 || n < 1 || stride < w) return SVS_ERR_ARG;
    SVS_CUDA(c, cudaSetDevice(c->device));
    SVS_TRY(svs_i_stereo_bm(c, left_dev, right_dev, w, h, stride, img_stride, n, ndisp, block, disp_out_dev));
    SVS_CUDA(c, cudaStreamSynchronize(c->stream));
    return SVS_OK;
}

int svs_backproject(svs_ctx *c, const int16_t *disp, const uint8_t *bgr, int w, int h, const double K[4], double baseline,
                    const double cam_pose_inv[7], const double T_cw[7], float *xyz_out, uint8_t *rgb_out, int32_t *n_out)
{
    if (!c || !disp || !bgr || !K || !cam_pose_inv || !T_cw || !xyz_out || !rgb_out || !n_out || w < 1 || h < 1) return SVS_ERR_ARG;
    SVS_CUDA(c, cudaSetDevice(c->device));
    size_t px = (size_t)w * h;
    SVS_CUDA(c, c->d_in.reserve(px * 2));
    SVS_CUDA(c, c->d_in2.reserve(px * 3));
    SVS_CUDA(c, c->d_tmp2.reserve((size_t)w * 4 + 16 + 14 * 8));
    SVS_CUDA(c, c->d_out.reserve(px * 12));
    SVS_CUDA(c, c->d_out2.reserve(px * 3));
    // T_wc = T_cw^-1 on the host (Sophus::SE3d::inverse: conjugate quaternion, -R^T t)
    double Twc[7] = {-T_cw[0], -T_cw[1], -T_cw[2], T_cw[3], 0, 0, 0};
    {
        const double *q = Twc, *p = T_cw + 4;
        double ux = 2.0 * (q[1] * p[2] - q[2] * p[1]), uy = 2.0 * (q[2] * p[0] - q[0] * p[2]), uz = 2.0 * (q[0] * p[1] - q[1] * p[0]);
        Twc[4] = -(p[0] + q[3] * ux + (q[1] * uz - q[2] * uy));
        Twc[5] = -(p[1] + q[3] * uy + (q[2] * ux - q[0] * uz));
        Twc[6] = -(p[2] + q[3] * uz + (q[0] * uy - q[1] * ux));
    }
    int *colc = c->d_tmp2.as<int>();
    int *total = colc + w;
    double *dT = reinterpret_cast<double *>(c->d_tmp2.as<uint8_t>() + align_up((size_t)(w + 1) * 4, 16));
    double hT[14];
    memcpy(hT, cam_pose_inv, 56); memcpy(hT + 7, Twc, 56);
    SVS_CUDA(c, cudaMemcpyAsync(dT, hT, sizeof(hT), cudaMemcpyHostToDevice, c->stream));
    SVS_CUDA(c, cudaMemcpyAsync(c->d_in.p, disp, px * 2, cudaMemcpyHostToDevice, c->stream));
    SVS_CUDA(c, cudaMemcpyAsync(c->d_in2.p, bgr, px * 3, cudaMemcpyHostToDevice, c->stream));
    float fb = (float)K[0] * (float)baseline;   // (focal_length * baseline) in f32, src/dense_reconstruction.cpp:120-136
    SVS_KERNEL(c, KID_BACKPROJECT, k_bp_count<<<(w + 127) / 128, 128, 0, c->stream>>>(c->d_in.as<int16_t>(), w, h, fb, colc));
    SVS_KERNEL(c, KID_BACKPROJECT, k_bp_scan<<<1, 32, 0, c->stream>>>(colc, w, total));
    SVS_KERNEL(c, KID_BACKPROJECT, k_bp_emit<<<(w + 127) / 128, 128, 0, c->stream>>>(c->d_in.as<int16_t>(), c->d_in2.as<uint8_t>(), w, h, fb, K[0], K[1], K[2], K[3],
                                                      dT, dT + 7, colc, c->d_out.as<float>(), c->d_out2.as<uint8_t>()));
    int tot = 0;
    SVS_CUDA(c, cudaMemcpyAsync(&tot, total, 4, cudaMemcpyDeviceToHost, c->stream));
    SVS_CUDA(c, cudaStreamSynchronize(c->stream));
    if (tot > 0) {
        SVS_CUDA(c, cudaMemcpyAsync(xyz_out, c->d_out.p, (size_t)tot * 12, cudaMemcpyDeviceToHost, c->stream));
        SVS_CUDA(c, cudaMemcpyAsync(rgb_out, c->d_out2.p, (size_t)tot * 3, cudaMemcpyDeviceToHost, c->stream));
        SVS_CUDA(c, cudaStreamSynchronize(c->stream));
    }
    *n_out = tot;
    return SVS_OK;
}

}  // extern "C"
