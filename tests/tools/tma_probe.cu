// Probe (GPU box): which cp.async.bulk.tensor box / coordinate combinations does the hardware accept for u8 tensors?
// nvcc -gencode arch=compute_100a,code=sm_100a -o /tmp/tma_probe tests/tools/tma_probe.cu && /tmp/tma_probe
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
#include <cstring>
#include <vector>
#include "../../stereovision-slam_b200/csrc/tma.cuh"

int svs_i_tmap_u8_3d(CUtensorMap *m, const void *base, int w, int h, int n, size_t row_stride, size_t img_pitch, int box_w, int box_h)
{
    typedef CUresult (*EncodeFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                 const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                 CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
    void *p = nullptr;
    cudaDriverEntryPointQueryResult q;
    cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q);
    EncodeFn fn = (EncodeFn)p;
    cuuint64_t gdim[3] = {(cuuint64_t)w, (cuuint64_t)h, (cuuint64_t)n};
    cuuint64_t gstr[2] = {(cuuint64_t)row_stride, (cuuint64_t)img_pitch};
    cuuint32_t box[3] = {(cuuint32_t)box_w, (cuuint32_t)box_h, 1}, estr[3] = {1, 1, 1};
    CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_UINT8, 3, const_cast<void *>(base), gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                    CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS ? 0 : -3;
}

__global__ void k(const __grid_constant__ CUtensorMap tm, int x, int y, int z, int bytes, uint8_t *out)
{
    extern __shared__ __align__(128) uint8_t sm[];
    __shared__ __align__(8) uint64_t bar;
    if (threadIdx.x == 0) { tma::mbar_init(&bar, 1); tma::fence_init(); }
    __syncthreads();
    if (threadIdx.x == 0) { tma::mbar_expect_tx(&bar, bytes); tma::load_3d(sm, &tm, &bar, x, y, z); }
    tma::mbar_wait(&bar, 0);
    for (int i = threadIdx.x; i < bytes; i += blockDim.x) out[i] = sm[i];
}

int main()
{
    const int W = 620, H = 188, N = 3, S = 624;
    size_t pitch = (size_t)S * H;
    pitch = (pitch + 255) / 256 * 256;
    std::vector<uint8_t> h(pitch * N);
    for (size_t i = 0; i < h.size(); i++) h[i] = (uint8_t)(i * 7 + i / S);
    uint8_t *d, *o;
    cudaMalloc(&d, h.size()); cudaMalloc(&o, 65536);
    cudaMemcpy(d, h.data(), h.size(), cudaMemcpyHostToDevice);
    int boxes[][2] = {{160, 35}, {96, 20}, {256, 35}};
    int xs[] = {0, 16, 128, -16, 608, 112};
    for (auto &b : boxes) {
        CUtensorMap tm;
        int e = svs_i_tmap_u8_3d(&tm, d, W, H, N, S, pitch, b[0], b[1]);
        printf("box %dx%d encode %d\n", b[0], b[1], e);
        if (e) continue;
        for (int x : xs) {
            int y = 30, z = 1, bytes = b[0] * b[1];
            cudaMemset(o, 0xEE, 65536);
            k<<<1, 128, bytes + 128>>>(tm, x, y, z, bytes, o);
            cudaError_t err = cudaDeviceSynchronize();
            if (err != cudaSuccess) { printf("  x=%d: %s\n", x, cudaGetErrorString(err)); return 1; }
            std::vector<uint8_t> r(bytes);
            cudaMemcpy(r.data(), o, bytes, cudaMemcpyDeviceToHost);
            int bad = 0;
            for (int j = 0; j < b[1]; j++) for (int i = 0; i < b[0]; i++) {
                int xx = x + i, yy = y + j;
                uint8_t want = (xx >= 0 && xx < W && yy >= 0 && yy < H) ? h[z * pitch + (size_t)yy * S + xx] : 0;
                bad += r[j * b[0] + i] != want;
            }
            printf("  x=%d ok, mismatches %d\n", x, bad);
        }
    }
    return 0;
}
