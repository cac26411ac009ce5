// peaks.cu -- bvc_measure_peaks: the instruction-issue ceilings the kernels are measured against, taken on the device and
// at the clocks of the run that quotes them (the search roofline is VABSDIFF4.U8.ACC issue on the ALU pipe, the
// transform's is DFMA issue; NVIDIA publishes neither rate for sm_100).  Dependent-free unrolled chains, 8 CTAs x 256
// threads per SM, best of several launches timed with CUDA events.
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/bvc.h"

namespace {

constexpr int PK_ITERS = 4096, PK_CH = 8;

template <int OP>
__global__ void __launch_bounds__(256) peak_kernel(uint32_t* out, uint32_t seed) {
    uint32_t a[PK_CH], acc[PK_CH];
    const uint32_t b = seed ^ threadIdx.x * 2654435761u;
    double d[PK_CH];
#pragma unroll
    for (int j = 0; j < PK_CH; j++) { a[j] = b + j * 0x01020304u; acc[j] = j; d[j] = (double)(b & 1023) * 1e-3 + j; }
    const double dm = 1.0000001, da = 1e-9;
#pragma unroll 1
    for (int it = 0; it < PK_ITERS; it++) {
#pragma unroll
        for (int j = 0; j < PK_CH; j++) {
            if (OP == 0) asm volatile("vabsdiff4.u32.u32.u32.add %0, %1, %2, %0;" : "+r"(acc[j]) : "r"(a[j]), "r"(b));
            if (OP == 1) asm volatile("fma.rn.f64 %0, %0, %1, %2;" : "+d"(d[j]) : "d"(dm), "d"(da));
        }
    }
    uint32_t s = 0;
    double ds = 0;
#pragma unroll
    for (int j = 0; j < PK_CH; j++) { s += acc[j] + a[j]; ds += d[j]; }
    out[blockIdx.x * blockDim.x + threadIdx.x] = s + (uint32_t)ds;
}

template <int OP>
cudaError_t measure(int nsm, uint32_t* out, double* ops_per_s) {
    const int grid = nsm * 8;
    cudaEvent_t e0, e1;
    cudaError_t e;
    if ((e = cudaEventCreate(&e0)) != cudaSuccess) return e;
    if ((e = cudaEventCreate(&e1)) != cudaSuccess) { cudaEventDestroy(e0); return e; }
    double best_ms = 1e30;
    for (int rep = 0; rep < 6; rep++) {   // rep 0 warms up (clocks, instruction cache)
        cudaEventRecord(e0);
        peak_kernel<OP><<<grid, 256>>>(out, 123u + rep);
        cudaEventRecord(e1);
        if ((e = cudaEventSynchronize(e1)) != cudaSuccess) break;
        float ms = 0;
        cudaEventElapsedTime(&ms, e0, e1);
        if (rep > 0 && ms < best_ms) best_ms = ms;
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    if (e != cudaSuccess) return e;
    *ops_per_s = (double)grid * 256.0 * PK_ITERS * PK_CH / (best_ms * 1e-3);
    return cudaGetLastError();
}

}  // namespace

extern "C" int bvc_measure_peaks(int device, double* vabsdiff4_thread_ops_per_s, double* dfma_thread_ops_per_s, int* sm_count) {
    if (cudaSetDevice(device) != cudaSuccess) return BVC_ERR_CUDA;
    int nsm = 0;
    if (cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, device) != cudaSuccess || nsm < 1) return BVC_ERR_CUDA;
    uint32_t* out = nullptr;
    if (cudaMalloc((void**)&out, (size_t)nsm * 8 * 256 * sizeof(uint32_t)) != cudaSuccess) return BVC_ERR_CUDA;
    double v = 0, f = 0;
    cudaError_t e = measure<0>(nsm, out, &v);
    if (e == cudaSuccess) e = measure<1>(nsm, out, &f);
    cudaFree(out);
    if (e != cudaSuccess) return BVC_ERR_CUDA;
    if (vabsdiff4_thread_ops_per_s) *vabsdiff4_thread_ops_per_s = v;
    if (dfma_thread_ops_per_s) *dfma_thread_ops_per_s = f;
    if (sm_count) *sm_count = nsm;
    return BVC_OK;
}
