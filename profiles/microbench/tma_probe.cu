// Probe: 3-D uint8 TMA tile load with byte-granular, possibly negative box origin (zero fill) on sm_100a.
#include <cstdio>
#include <cstdint>
#include <vector>
#include <cuda.h>
#include <cuda_runtime.h>
#include "../../basic_video_codec_b200/csrc/bvc_common.cuh"
using namespace bvc;
__global__ void probe(const __grid_constant__ CUtensorMap map, uint8_t* out, int bw, int bh, int x, int y, int z, int variant) {
    extern __shared__ __align__(128) uint8_t smem[];
    __shared__ uint64_t bar;
    if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_mbar_init(); }
    __syncthreads();
    if (threadIdx.x == 0) {
        mbar_arrive_expect_tx(&bar, bw * bh);
        tma_load_3d(smem, &map, &bar, x, y, z);
    }
    mbar_wait(&bar, 0);
    for (int i = threadIdx.x; i < bw * bh; i += blockDim.x) out[i] = smem[i];
}
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
int main() {
    const int W = 96, H = 64, P = 3, pitch = 96; size_t pb = (size_t)pitch * H;
    std::vector<uint8_t> h(pb * P);
    for (size_t i = 0; i < h.size(); i++) h[i] = (uint8_t)(i * 7 + i / 96);
    uint8_t* d; cudaMalloc(&d, h.size()); cudaMemcpy(d, h.data(), h.size(), cudaMemcpyHostToDevice);
    void* fn = nullptr; cudaDriverEntryPointQueryResult q;
    cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q);
    printf("entry %p q=%d\n", fn, (int)q);
    for (int bwid : {80, 128}) {
        const int bh = 32;
        CUtensorMap map;
        cuuint64_t dims[3] = {W, H, P}; cuuint64_t strides[2] = {(cuuint64_t)pitch, pb};
        cuuint32_t box[3] = {(cuuint32_t)bwid, bh, 1}; cuuint32_t es[3] = {1, 1, 1};
        CUresult r = ((EncodeTiledFn)fn)(&map, CU_TENSOR_MAP_DATA_TYPE_UINT8, 3, d, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                         CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        printf("encode box %d -> %d\n", bwid, (int)r);
        uint8_t* o; cudaMalloc(&o, bwid * bh);
        for (int x : {0, 16, -16, -32, 48, 80, -8}) {
            int y = -8, z = 1;
            probe<<<1, 128, bwid * bh + 128>>>(map, o, bwid, bh, x, y, z, 0);
            cudaError_t e = cudaDeviceSynchronize();
            std::vector<uint8_t> ho(bwid * bh);
            cudaMemcpy(ho.data(), o, ho.size(), cudaMemcpyDeviceToHost);
            int bad = 0;
            for (int r2 = 0; r2 < bh; r2++) for (int c = 0; c < bwid; c++) {
                int gx = x + c, gy = y + r2; uint8_t want = (gx < 0 || gy < 0 || gx >= W || gy >= H) ? 0 : h[z * pb + gy * pitch + gx];
                if (ho[r2 * bwid + c] != want) bad++;
            }
            printf("  x=%d: err=%s mismatches=%d\n", x, cudaGetErrorString(e), bad);
            if (e != cudaSuccess) return 1;
        }
    }
    return 0;
}
