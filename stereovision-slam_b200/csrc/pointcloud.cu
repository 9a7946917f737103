// Dense-map post-processing: what DenseReconstruction::DenseReconstruct does with PCL right after the
// back-projection, reference src/dense_reconstruction.cpp:175-209 —
//   pcl::StatisticalOutlierRemoval, setMeanK(50), setStddevMulThresh(1.0), per keyframe (:179-184) and on the merged map
//   (:194-200): mean distance of every point to its 50 nearest neighbours, keep mean <= mu + 1.0 * sigma
//   pcl::VoxelGrid, leaf 0.02 m (:203-209): one centroid (xyz and colour) per occupied voxel, ascending voxel index
// PCL is not vendored (and absent here): the arithmetic below restates PCL 1.12's published filters (float points, FLANN's
// float L2 for the neighbour distances, double statistics; float voxel indices and float centroid sums) — parity unpinned,
// checked against the NumPy / scipy restatement the tests hold (tests/test_gpu_pointcloud.py).
//
// k-NN: a uniform grid over the bounding box (points counting-sorted by cell, cell size adapted to the occupancy) and one
// thread per point walking Chebyshev rings of cells with a 51-entry max-heap in local memory; a ring r search is complete as
// soon as the (k+1)-th best distance is <= r * cell (nothing outside the searched cube can be closer), so the result is the
// EXACT neighbour set.  Voxel grid: keys as PCL computes them, stable radix sort (CUB), one thread per voxel summing its
// points in index order.  CUB's device radix sort is used as a library primitive for the two sorts; everything else is here.
#include "svs_internal.h"
#include <cub/cub.cuh>
#include <cfloat>
#include <climits>
#include <cmath>
#include <cstring>
#include <vector>

#define PC_KMAX 64

struct PcGrid { float ox, oy, oz, cell; int nx, ny, nz; };

__global__ void k_pc_bounds(const float *__restrict__ xyz, int n, float *__restrict__ mnmx /* 6: min xyz, max xyz (ordered ints) */)
{
    __shared__ float s[6][256];
    float mn[3] = {FLT_MAX, FLT_MAX, FLT_MAX}, mx[3] = {-FLT_MAX, -FLT_MAX, -FLT_MAX};
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const float x = xyz[3 * (size_t)i], y = xyz[3 * (size_t)i + 1], z = xyz[3 * (size_t)i + 2];
        if (!(fabsf(x) <= FLT_MAX && fabsf(y) <= FLT_MAX && fabsf(z) <= FLT_MAX)) continue;
        mn[0] = fminf(mn[0], x); mn[1] = fminf(mn[1], y); mn[2] = fminf(mn[2], z);
        mx[0] = fmaxf(mx[0], x); mx[1] = fmaxf(mx[1], y); mx[2] = fmaxf(mx[2], z);
    }
    for (int a = 0; a < 3; a++) { s[a][threadIdx.x] = mn[a]; s[3 + a][threadIdx.x] = mx[a]; }
    __syncthreads();
    for (int st = 128; st > 0; st >>= 1) {
        if (threadIdx.x < st)
            for (int a = 0; a < 3; a++) {
                s[a][threadIdx.x] = fminf(s[a][threadIdx.x], s[a][threadIdx.x + st]);
                s[3 + a][threadIdx.x] = fmaxf(s[3 + a][threadIdx.x], s[3 + a][threadIdx.x + st]);
            }
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        // float atomic min / max through the order-preserving integer image
        for (int a = 0; a < 6; a++) {
            const float v = s[a][0];
            int iv = __float_as_int(v);
            iv = iv >= 0 ? iv : iv ^ 0x7FFFFFFF;
            if (a < 3) atomicMin(reinterpret_cast<int *>(mnmx) + a, iv); else atomicMax(reinterpret_cast<int *>(mnmx) + a, iv);
        }
    }
}
__host__ __device__ inline float pc_ordered_to_float(int iv) { iv = iv >= 0 ? iv : iv ^ 0x7FFFFFFF; float f; memcpy(&f, &iv, 4); return f; }

__device__ __forceinline__ int pc_cell_of(const PcGrid &g, float x, float y, float z, int &cx, int &cy, int &cz)
{
    cx = min(max((int)floorf((x - g.ox) / g.cell), 0), g.nx - 1);
    cy = min(max((int)floorf((y - g.oy) / g.cell), 0), g.ny - 1);
    cz = min(max((int)floorf((z - g.oz) / g.cell), 0), g.nz - 1);
    return (cz * g.ny + cy) * g.nx + cx;
}
__global__ void k_pc_cell_keys(const float *__restrict__ xyz, int n, PcGrid g, unsigned *__restrict__ keys, int *__restrict__ idx)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float x = xyz[3 * (size_t)i], y = xyz[3 * (size_t)i + 1], z = xyz[3 * (size_t)i + 2];
    int cx, cy, cz;
    const bool fin = fabsf(x) <= FLT_MAX && fabsf(y) <= FLT_MAX && fabsf(z) <= FLT_MAX;
    keys[i] = fin ? (unsigned)pc_cell_of(g, x, y, z, cx, cy, cz) : 0xFFFFFFFFu;      // non-finite points sort to the end
    idx[i] = i;
}
__global__ void k_pc_cell_ranges(const unsigned *__restrict__ keys, int n, int n_cells, int *__restrict__ cstart, int *__restrict__ cend)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const unsigned k = keys[i];
    if (k >= (unsigned)n_cells) return;
    if (i == 0 || keys[i - 1] != k) cstart[k] = i;
    if (i == n - 1 || keys[i + 1] != k) cend[k] = i + 1;
}
__global__ void k_pc_occupancy(const unsigned *__restrict__ keys, int n, int *__restrict__ out /* occupied cells, finite points */)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    int head = 0, fin = 0;
    if (i < n) { fin = keys[i] != 0xFFFFFFFFu; head = fin && (i == 0 || keys[i - 1] != keys[i]); }
    head = __reduce_add_sync(0xffffffffu, head); fin = __reduce_add_sync(0xffffffffu, fin);
    if ((threadIdx.x & 31) == 0) { if (head) atomicAdd(out, head); if (fin) atomicAdd(out + 1, fin); }
}
__global__ void k_pc_gather(const float *__restrict__ xyz, const int *__restrict__ idx, int n, float4 *__restrict__ sorted)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int s = idx[i];
    sorted[i] = make_float4(xyz[3 * (size_t)s], xyz[3 * (size_t)s + 1], xyz[3 * (size_t)s + 2], __int_as_float(s));
}

// mean distance to the k nearest neighbours (PCL StatisticalOutlierRemoval::applyFilterIndices): the (k+1)-NN query includes
// the point itself, which is dropped; squared distances in float exactly as FLANN's L2_Simple accumulates them
// (((dx*dx) + dy*dy) + dz*dz, no contraction: this file is built with --fmad=false), sqrtf, sum in double, result float.
__global__ void __launch_bounds__(128)
k_pc_knn_mean(const float4 *__restrict__ pts, int n_valid, PcGrid g, const int *__restrict__ cstart, const int *__restrict__ cend, int k,
              float *__restrict__ mean_dist /* by ORIGINAL index */)
{
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n_valid) return;
    const float4 q = pts[t];
    float heap[PC_KMAX + 1];          // max-heap of the k+1 smallest squared distances seen so far
    int hn = 0;
    const int K1 = k + 1;
    int cx, cy, cz;
    pc_cell_of(g, q.x, q.y, q.z, cx, cy, cz);
    const int rmax = max(max(max(cx, g.nx - 1 - cx), max(cy, g.ny - 1 - cy)), max(cz, g.nz - 1 - cz));
    auto visit = [&](int c) {
        const int s0 = cstart[c];
        if (s0 < 0) return;
        const int s1 = cend[c];
        for (int j = s0; j < s1; j++) {
            const float4 p = pts[j];
            const float dx = q.x - p.x, dy = q.y - p.y, dz = q.z - p.z;
            const float d2 = __fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz));
            if (hn < K1) {                    // sift up
                int i = hn++;
                while (i > 0) { const int pr = (i - 1) >> 1; if (heap[pr] >= d2) break; heap[i] = heap[pr]; i = pr; }
                heap[i] = d2;
            } else if (d2 < heap[0]) {        // replace the maximum, sift down
                int i = 0;
                for (;;) {
                    int ch = 2 * i + 1;
                    if (ch >= K1) break;
                    if (ch + 1 < K1 && heap[ch + 1] > heap[ch]) ch++;
                    if (heap[ch] <= d2) break;
                    heap[i] = heap[ch]; i = ch;
                }
                heap[i] = d2;
            }
        }
    };
    for (int r = 0; r <= rmax; r++) {       // the shell of cells at Chebyshev distance exactly r
        const int z0 = max(cz - r, 0), z1 = min(cz + r, g.nz - 1), y0 = max(cy - r, 0), y1 = min(cy + r, g.ny - 1);
        const int xa = max(cx - r, 0), xb = min(cx + r, g.nx - 1);
        for (int z = z0; z <= z1; z++)
            for (int y = y0; y <= y1; y++) {
                const int rowc = (z * g.ny + y) * g.nx;
                if (z == cz - r || z == cz + r || y == cy - r || y == cy + r) {
                    for (int x = xa; x <= xb; x++) visit(rowc + x);
                } else {
                    if (cx - r >= 0) visit(rowc + cx - r);
                    if (cx + r <= g.nx - 1) visit(rowc + cx + r);
                }
            }
        if (hn == K1 && r >= 1) {
            // every unsearched point lies outside the (2r+1)^3 cube of cells around the query's cell: at least r cells away
            // (1e-3 of a cell of slack for the float rounding of the cell assignment)
            const double reach = ((double)r - 1e-3) * (double)g.cell;
            if ((double)heap[0] <= reach * reach) break;
        }
    }
    // sum of sqrt over the k+1 neighbours minus the nearest one (the point itself, distance 0), in ascending order like PCL's
    // sorted result: heap-sort in place
    for (int m = hn - 1; m > 0; m--) {
        const float top = heap[0], v = heap[m];
        heap[m] = top;
        int i = 0;
        for (;;) {
            int ch = 2 * i + 1;
            if (ch >= m) break;
            if (ch + 1 < m && heap[ch + 1] > heap[ch]) ch++;
            if (heap[ch] <= v) break;
            heap[i] = heap[ch]; i = ch;
        }
        heap[i] = v;
    }
    double sum = 0.0;
    for (int i = 1; i < hn; i++) sum += (double)__fsqrt_rn(heap[i]);
    mean_dist[__float_as_int(q.w)] = (float)(sum / (double)k);
}

__global__ void k_pc_stats(const float *__restrict__ d, int n, double *__restrict__ part /* 2 per CTA */)
{
    __shared__ double s0[256], s1[256];
    double a = 0, b = 0;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) { const float v = d[i]; a += (double)v; b += (double)__fmul_rn(v, v); }
    s0[threadIdx.x] = a; s1[threadIdx.x] = b;
    __syncthreads();
    for (int st = 128; st > 0; st >>= 1) { if (threadIdx.x < st) { s0[threadIdx.x] += s0[threadIdx.x + st]; s1[threadIdx.x] += s1[threadIdx.x + st]; } __syncthreads(); }
    if (threadIdx.x == 0) { part[2 * blockIdx.x] = s0[0]; part[2 * blockIdx.x + 1] = s1[0]; }
}
__global__ void k_pc_keep(const float *__restrict__ d, int n, double thr, uint8_t *__restrict__ keep)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) keep[i] = !((double)d[i] > thr);
}

// ------------------------------------------------------------------------------------------------ voxel grid
struct VoxParams { float inv[3]; int minb[3]; int mul[3]; };
__global__ void k_vox_keys(const float *__restrict__ xyz, int n, VoxParams p, unsigned *__restrict__ keys, int *__restrict__ idx)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float x = xyz[3 * (size_t)i], y = xyz[3 * (size_t)i + 1], z = xyz[3 * (size_t)i + 2];
    idx[i] = i;
    if (!(fabsf(x) <= FLT_MAX && fabsf(y) <= FLT_MAX && fabsf(z) <= FLT_MAX)) { keys[i] = 0xFFFFFFFFu; return; }
    // pcl::VoxelGrid::applyFilter: static_cast<int>(std::floor(p.x * inverse_leaf_size_[0]) - static_cast<float>(min_b_[0]))
    const int i0 = (int)__fsub_rn(floorf(__fmul_rn(x, p.inv[0])), (float)p.minb[0]);
    const int i1 = (int)__fsub_rn(floorf(__fmul_rn(y, p.inv[1])), (float)p.minb[1]);
    const int i2 = (int)__fsub_rn(floorf(__fmul_rn(z, p.inv[2])), (float)p.minb[2]);
    keys[i] = (unsigned)(i0 * p.mul[0] + i1 * p.mul[1] + i2 * p.mul[2]);
}
__global__ void k_vox_heads(const unsigned *__restrict__ keys, int n, int *__restrict__ head_flag)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    head_flag[i] = (keys[i] != 0xFFFFFFFFu && (i == 0 || keys[i - 1] != keys[i])) ? 1 : 0;
}
__global__ void k_vox_centroids(const float *__restrict__ xyz, const uint8_t *__restrict__ rgb, const unsigned *__restrict__ keys, const int *__restrict__ idx,
                                const int *__restrict__ head_flag, const int *__restrict__ head_rank /* exclusive scan */, int n,
                                float *__restrict__ xyz_out, uint8_t *__restrict__ rgb_out)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n || !head_flag[i]) return;
    const unsigned key = keys[i];
    // pcl::CentroidPoint: float accumulators, divided by the point count as float; colours truncated
    float sx = 0.f, sy = 0.f, sz = 0.f, sr = 0.f, sg = 0.f, sb = 0.f;
    int cnt = 0;
    for (int j = i; j < n && keys[j] == key; j++) {
        const int s = idx[j];
        sx = __fadd_rn(sx, xyz[3 * (size_t)s]); sy = __fadd_rn(sy, xyz[3 * (size_t)s + 1]); sz = __fadd_rn(sz, xyz[3 * (size_t)s + 2]);
        if (rgb) { sr = __fadd_rn(sr, (float)rgb[3 * (size_t)s]); sg = __fadd_rn(sg, (float)rgb[3 * (size_t)s + 1]); sb = __fadd_rn(sb, (float)rgb[3 * (size_t)s + 2]); }
        cnt++;
    }
    const int o = head_rank[i];
    const float fn = (float)cnt;
    xyz_out[3 * (size_t)o] = __fdiv_rn(sx, fn); xyz_out[3 * (size_t)o + 1] = __fdiv_rn(sy, fn); xyz_out[3 * (size_t)o + 2] = __fdiv_rn(sz, fn);
    if (rgb && rgb_out) {
        rgb_out[3 * (size_t)o] = (uint8_t)(unsigned)__fdiv_rn(sr, fn); rgb_out[3 * (size_t)o + 1] = (uint8_t)(unsigned)__fdiv_rn(sg, fn);
        rgb_out[3 * (size_t)o + 2] = (uint8_t)(unsigned)__fdiv_rn(sb, fn);
    }
}

static int pc_sort_pairs(svs_ctx *c, unsigned *keys_in, unsigned *keys_out, int *val_in, int *val_out, int n, int end_bit)
{
    size_t tmp_bytes = 0;
    SVS_CUDA(c, cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, keys_in, keys_out, val_in, val_out, n, 0, end_bit, c->stream));
    SVS_CUDA(c, c->d_tmp6.reserve(tmp_bytes + 256));
    SVS_CUDA(c, cub::DeviceRadixSort::SortPairs(c->d_tmp6.p, tmp_bytes, keys_in, keys_out, val_in, val_out, n, 0, end_bit, c->stream));
    c->launches += 4;
    return SVS_OK;
}

extern "C" {

int svs_pointcloud_sor(svs_ctx *c, const float *xyz, int n, int mean_k, double stddev_mul, uint8_t *keep_out, float *mean_dist_out, int *n_kept)
{
    if (!c || !xyz || !keep_out || n < 0 || mean_k < 1 || mean_k > PC_KMAX - 1) return SVS_ERR_ARG;
    if (n_kept) *n_kept = 0;
    if (n == 0) return SVS_OK;
    if (n <= mean_k) SVS_FAIL(c, SVS_ERR_ARG, "pointcloud_sor: fewer points than mean_k + 1");
    SVS_CUDA(c, cudaSetDevice(c->device));
    const size_t N = (size_t)n;
    // device layout
    size_t off = 0;
    auto take = [&](size_t b) { size_t o = off; off = align_up(off + b, 256); return o; };
    const size_t o_xyz = take(N * 12), o_k0 = take(N * 4), o_k1 = take(N * 4), o_i0 = take(N * 4), o_i1 = take(N * 4), o_sorted = take(N * 16),
                 o_md = take(N * 4), o_keep = take(N), o_mm = take(64), o_part = take(2 * 1024 * 8);
    SVS_CUDA(c, c->d_tmp.reserve(off));
    uint8_t *db = c->d_tmp.as<uint8_t>();
    float *d_xyz = (float *)(db + o_xyz), *d_md = (float *)(db + o_md), *d_mm = (float *)(db + o_mm);
    unsigned *k0 = (unsigned *)(db + o_k0), *k1 = (unsigned *)(db + o_k1);
    int *i0 = (int *)(db + o_i0), *i1 = (int *)(db + o_i1);
    float4 *sorted = (float4 *)(db + o_sorted);
    double *part = (double *)(db + o_part);
    SVS_CUDA(c, cudaMemcpyAsync(d_xyz, xyz, N * 12, cudaMemcpyHostToDevice, c->stream));
    // bounding box of the finite points
    {
        int init[6];
        float big = FLT_MAX, small = -FLT_MAX;
        int ib, is;
        memcpy(&ib, &big, 4); memcpy(&is, &small, 4);
        is = is ^ 0x7FFFFFFF;
        for (int a = 0; a < 3; a++) { init[a] = ib; init[3 + a] = is; }
        SVS_CUDA(c, cudaMemcpyAsync(d_mm, init, sizeof(init), cudaMemcpyHostToDevice, c->stream));
        SVS_KERNEL(c, KID_MISC, k_pc_bounds<<<std::min(1024, (n + 255) / 256), 256, 0, c->stream>>>(d_xyz, n, d_mm));
    }
    int mm_i[6];
    SVS_CUDA(c, cudaMemcpyAsync(mm_i, d_mm, sizeof(mm_i), cudaMemcpyDeviceToHost, c->stream));
    SVS_CUDA(c, cudaStreamSynchronize(c->stream));
    float mn[3], mx[3];
    for (int a = 0; a < 3; a++) { mn[a] = pc_ordered_to_float(mm_i[a]); mx[a] = pc_ordered_to_float(mm_i[3 + a]); }
    if (!(mn[0] <= mx[0])) SVS_FAIL(c, SVS_ERR_ARG, "pointcloud_sor: no finite point");
    // grid: cubic cells, adapted so that an occupied cell holds ~ mean_k / 4 .. mean_k points (surfaces, not volumes)
    PcGrid g;
    const float ex = mx[0] - mn[0], ey = mx[1] - mn[1], ez = mx[2] - mn[2];
    const float emax = std::max(std::max(ex, ey), std::max(ez, 1e-6f));
    double cell = std::cbrt((double)std::max(ex, emax * 1e-3f) * std::max(ey, emax * 1e-3f) * std::max(ez, emax * 1e-3f) * (double)mean_k / (double)n);
    int n_valid = 0, n_cells = 0;
    for (int attempt = 0; attempt < 6; attempt++) {
        cell = std::max(cell, (double)emax / 128.0);
        g.ox = mn[0]; g.oy = mn[1]; g.oz = mn[2]; g.cell = (float)cell;
        g.nx = std::max(1, (int)std::floor(ex / g.cell) + 1); g.ny = std::max(1, (int)std::floor(ey / g.cell) + 1); g.nz = std::max(1, (int)std::floor(ez / g.cell) + 1);
        n_cells = g.nx * g.ny * g.nz;
        SVS_CUDA(c, c->d_tmp2.reserve((size_t)n_cells * 8 + 64));
        int *cstart = c->d_tmp2.as<int>(), *cend = cstart + n_cells;
        SVS_CUDA(c, cudaMemsetAsync(cstart, 0xFF, (size_t)n_cells * 4, c->stream));
        SVS_KERNEL(c, KID_MISC, k_pc_cell_keys<<<(n + 255) / 256, 256, 0, c->stream>>>(d_xyz, n, g, k0, i0));
        SVS_TRY(pc_sort_pairs(c, k0, k1, i0, i1, n, 32));
        SVS_KERNEL(c, KID_MISC, k_pc_cell_ranges<<<(n + 255) / 256, 256, 0, c->stream>>>(k1, n, n_cells, cstart, cend));
        // occupancy of the grid (counted on the device): occupied cells and finite points
        int *d_occ = reinterpret_cast<int *>(d_mm) + 8;
        SVS_CUDA(c, cudaMemsetAsync(d_occ, 0, 8, c->stream));
        SVS_KERNEL(c, KID_MISC, k_pc_occupancy<<<(n + 255) / 256, 256, 0, c->stream>>>(k1, n, d_occ));
        int h_occ[2] = {0, 0};
        SVS_CUDA(c, cudaMemcpyAsync(h_occ, d_occ, 8, cudaMemcpyDeviceToHost, c->stream));
        SVS_CUDA(c, cudaStreamSynchronize(c->stream));
        const int occ = h_occ[0];
        n_valid = h_occ[1];
        const double per_cell = (double)n_valid / std::max(1, occ);
        if (per_cell > 1.5 * mean_k && cell > (double)emax / 128.0 * 1.01) { cell *= 0.6; continue; }
        if (per_cell < 0.2 * mean_k && n_cells > 1) { cell *= 1.7; continue; }
        break;
    }
    if (n_valid <= mean_k) SVS_FAIL(c, SVS_ERR_ARG, "pointcloud_sor: fewer finite points than mean_k + 1");
    int *cstart = c->d_tmp2.as<int>(), *cend = cstart + n_cells;
    SVS_KERNEL(c, KID_MISC, k_pc_gather<<<(n + 255) / 256, 256, 0, c->stream>>>(d_xyz, i1, n, sorted));
    SVS_CUDA(c, cudaMemsetAsync(d_md, 0, N * 4, c->stream));          // non-finite points keep distance 0 like PCL
    SVS_KERNEL(c, KID_MISC, k_pc_knn_mean<<<(n_valid + 127) / 128, 128, 0, c->stream>>>(sorted, n_valid, g, cstart, cend, mean_k, d_md));
    const int sb = std::min(1024, (n + 255) / 256);
    SVS_KERNEL(c, KID_MISC, k_pc_stats<<<sb, 256, 0, c->stream>>>(d_md, n, part));
    std::vector<double> hp(2 * (size_t)sb);
    SVS_CUDA(c, cudaMemcpyAsync(hp.data(), part, hp.size() * 8, cudaMemcpyDeviceToHost, c->stream));
    SVS_CUDA(c, cudaStreamSynchronize(c->stream));
    double sum = 0, sq = 0;
    for (int i = 0; i < sb; i++) { sum += hp[2 * i]; sq += hp[2 * i + 1]; }
    const double nv = (double)n_valid;
    const double mean = sum / nv, variance = (sq - sum * sum / nv) / (nv - 1.0), stddev = std::sqrt(variance);
    const double thr = mean + stddev_mul * stddev;
    uint8_t *d_keep = db + o_keep;
    SVS_KERNEL(c, KID_MISC, k_pc_keep<<<(n + 255) / 256, 256, 0, c->stream>>>(d_md, n, thr, d_keep));
    SVS_CUDA(c, cudaMemcpyAsync(keep_out, d_keep, N, cudaMemcpyDeviceToHost, c->stream));
    if (mean_dist_out) SVS_CUDA(c, cudaMemcpyAsync(mean_dist_out, d_md, N * 4, cudaMemcpyDeviceToHost, c->stream));
    SVS_CUDA(c, cudaStreamSynchronize(c->stream));
    if (n_kept) { int k = 0; for (int i = 0; i < n; i++) k += keep_out[i]; *n_kept = k; }
    return SVS_OK;
}

int svs_voxel_grid(svs_ctx *c, const float *xyz, const uint8_t *rgb, int n, double leaf, float *xyz_out, uint8_t *rgb_out, int *n_out)
{
    if (!c || !xyz || !xyz_out || !n_out || n < 0 || !(leaf > 0)) return SVS_ERR_ARG;
    *n_out = 0;
    if (n == 0) return SVS_OK;
    SVS_CUDA(c, cudaSetDevice(c->device));
    const size_t N = (size_t)n;
    size_t off = 0;
    auto take = [&](size_t b) { size_t o = off; off = align_up(off + b, 256); return o; };
    const size_t o_xyz = take(N * 12), o_rgb = take(N * 3), o_k0 = take(N * 4), o_k1 = take(N * 4), o_i0 = take(N * 4), o_i1 = take(N * 4),
                 o_hf = take(N * 4), o_hr = take(N * 4), o_xo = take(N * 12), o_ro = take(N * 3), o_mm = take(64);
    SVS_CUDA(c, c->d_tmp.reserve(off));
    uint8_t *db = c->d_tmp.as<uint8_t>();
    float *d_xyz = (float *)(db + o_xyz), *d_mm = (float *)(db + o_mm);
    uint8_t *d_rgb = rgb ? db + o_rgb : nullptr;
    unsigned *k0 = (unsigned *)(db + o_k0), *k1 = (unsigned *)(db + o_k1);
    int *i0 = (int *)(db + o_i0), *i1 = (int *)(db + o_i1), *hf = (int *)(db + o_hf), *hr = (int *)(db + o_hr);
    SVS_CUDA(c, cudaMemcpyAsync(d_xyz, xyz, N * 12, cudaMemcpyHostToDevice, c->stream));
    if (rgb) SVS_CUDA(c, cudaMemcpyAsync(d_rgb, rgb, N * 3, cudaMemcpyHostToDevice, c->stream));
    {
        int init[6];
        float big = FLT_MAX, small = -FLT_MAX;
        int ib, is;
        memcpy(&ib, &big, 4); memcpy(&is, &small, 4);
        is = is ^ 0x7FFFFFFF;
        for (int a = 0; a < 3; a++) { init[a] = ib; init[3 + a] = is; }
        SVS_CUDA(c, cudaMemcpyAsync(d_mm, init, sizeof(init), cudaMemcpyHostToDevice, c->stream));
        SVS_KERNEL(c, KID_MISC, k_pc_bounds<<<std::min(1024, (n + 255) / 256), 256, 0, c->stream>>>(d_xyz, n, d_mm));
    }
    int mm_i[6];
    SVS_CUDA(c, cudaMemcpyAsync(mm_i, d_mm, sizeof(mm_i), cudaMemcpyDeviceToHost, c->stream));
    SVS_CUDA(c, cudaStreamSynchronize(c->stream));
    float mn[3], mx[3];
    for (int a = 0; a < 3; a++) { mn[a] = pc_ordered_to_float(mm_i[a]); mx[a] = pc_ordered_to_float(mm_i[3 + a]); }
    auto passthrough = [&]() {          // PCL: "Leaf size is too small for the input dataset. Integer indices would overflow." -> output = input
        memcpy(xyz_out, xyz, N * 12);
        if (rgb && rgb_out) memcpy(rgb_out, rgb, N * 3);
        *n_out = n;
        return SVS_OK;
    };
    if (!(mn[0] <= mx[0])) return passthrough();
    // pcl::VoxelGrid::applyFilter (float leaf, float inverse), voxel_grid.hpp
    const float leaf_f = (float)leaf, inv = 1.0f / leaf_f;
    const long long dx = (long long)((mx[0] - mn[0]) * inv) + 1, dy = (long long)((mx[1] - mn[1]) * inv) + 1, dz = (long long)((mx[2] - mn[2]) * inv) + 1;
    if (dx * dy * dz > (long long)INT_MAX) return passthrough();
    VoxParams p;
    int maxb[3], divb[3];
    for (int a = 0; a < 3; a++) { p.inv[a] = inv; p.minb[a] = (int)std::floor(mn[a] * inv); maxb[a] = (int)std::floor(mx[a] * inv); divb[a] = maxb[a] - p.minb[a] + 1; }
    p.mul[0] = 1; p.mul[1] = divb[0]; p.mul[2] = divb[0] * divb[1];
    SVS_KERNEL(c, KID_MISC, k_vox_keys<<<(n + 255) / 256, 256, 0, c->stream>>>(d_xyz, n, p, k0, i0));
    SVS_TRY(pc_sort_pairs(c, k0, k1, i0, i1, n, 32));          // stable: points of a voxel stay in index order
    SVS_KERNEL(c, KID_MISC, k_vox_heads<<<(n + 255) / 256, 256, 0, c->stream>>>(k1, n, hf));
    {
        size_t tmp_bytes = 0;
        SVS_CUDA(c, cub::DeviceScan::ExclusiveSum(nullptr, tmp_bytes, hf, hr, n, c->stream));
        SVS_CUDA(c, c->d_tmp6.reserve(tmp_bytes + 256));
        SVS_CUDA(c, cub::DeviceScan::ExclusiveSum(c->d_tmp6.p, tmp_bytes, hf, hr, n, c->stream));
        c->launches += 2;
    }
    float *d_xo = (float *)(db + o_xo);
    uint8_t *d_ro = db + o_ro;
    SVS_KERNEL(c, KID_MISC, k_vox_centroids<<<(n + 255) / 256, 256, 0, c->stream>>>(d_xyz, d_rgb, k1, i1, hf, hr, n, d_xo, d_ro));
    int last_rank = 0, last_flag = 0;
    SVS_CUDA(c, cudaMemcpyAsync(&last_rank, hr + (n - 1), 4, cudaMemcpyDeviceToHost, c->stream));
    SVS_CUDA(c, cudaMemcpyAsync(&last_flag, hf + (n - 1), 4, cudaMemcpyDeviceToHost, c->stream));
    SVS_CUDA(c, cudaStreamSynchronize(c->stream));
    const int m = last_rank + last_flag;
    SVS_CUDA(c, cudaMemcpyAsync(xyz_out, d_xo, (size_t)m * 12, cudaMemcpyDeviceToHost, c->stream));
    if (rgb && rgb_out) SVS_CUDA(c, cudaMemcpyAsync(rgb_out, d_ro, (size_t)m * 3, cudaMemcpyDeviceToHost, c->stream));
    SVS_CUDA(c, cudaStreamSynchronize(c->stream));
    *n_out = m;
    return SVS_OK;
}

}  // extern "C"
