// cosched.cu -- does a small fp64 kernel run in the shadow of an ALU-bound kernel on B200?
// Kernel A mimics the motion search footprint (256 thr, ~96 regs, 75 KB smem, VABSDIFF4 bound, 2 CTAs/SM);
// kernel B mimics the transform kernel (128 thr, ~110 regs, 34 KB smem, fp64 + latency bound).
// We time A alone, B alone and A || B on two streams under three priority settings, and log where/when B's CTAs ran.
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <algorithm>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); exit(1); } } while (0)

__device__ __forceinline__ unsigned sad4(unsigned a, unsigned b, unsigned c) { unsigned d; asm("vabsdiff4.u32.u32.u32.add %0,%1,%2,%3;" : "=r"(d) : "r"(a), "r"(b), "r"(c)); return d; }
__device__ __forceinline__ unsigned long long gtime() { unsigned long long t; asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t)); return t; }
__device__ __forceinline__ unsigned smid() { unsigned s; asm volatile("mov.u32 %0, %smid;" : "=r"(s)); return s; }

template <int NT>
__global__ void __launch_bounds__(NT, 2) alu_hog(unsigned* out, int iters) {
    extern __shared__ unsigned sm[];
    unsigned cur[64], acc[16];
#pragma unroll
    for (int i = 0; i < 64; i++) cur[i] = threadIdx.x * 7 + i + blockIdx.x;
#pragma unroll
    for (int i = 0; i < 16; i++) acc[i] = 0;
    sm[threadIdx.x] = threadIdx.x;
    __syncthreads();
    for (int it = 0; it < iters; it++) {
        unsigned w = sm[(threadIdx.x + it) & 255];
#pragma unroll
        for (int i = 0; i < 64; i++) acc[i & 15] = sad4(w + i, cur[i], acc[i & 15]);
    }
    unsigned s = 0;
#pragma unroll
    for (int i = 0; i < 16; i++) s += acc[i];
    if (s == 0xdeadbeef) out[0] = s;
}

__global__ void __launch_bounds__(128, 4) dp_task(double* out, int iters, unsigned long long* log, int do_log) {
    extern __shared__ double smd[];
    const unsigned long long t0 = gtime();
    double a[40];
#pragma unroll
    for (int i = 0; i < 40; i++) a[i] = threadIdx.x * 0.001 + i;
    smd[threadIdx.x] = threadIdx.x;
    __syncthreads();
    for (int it = 0; it < iters; it++) {
        double w = smd[(threadIdx.x + it) & 127];   // shared-memory latency in the chain, like the transform's line exchange
#pragma unroll
        for (int i = 0; i < 40; i++) a[i] = fma(a[i], 1.0000001, w);
        smd[threadIdx.x] = a[it % 40 == 0 ? 0 : 1];
        __syncwarp();
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < 40; i++) s += a[i];
    if (s == 1.2345) out[0] = s;
    if (do_log && threadIdx.x == 0) { log[3 * blockIdx.x] = t0; log[3 * blockIdx.x + 1] = gtime(); log[3 * blockIdx.x + 2] = smid(); }
}

int main(int argc, char** argv) {
    const int nt = argc > 1 ? atoi(argv[1]) : 256;      // threads of kernel A: 256 or 320
    int lo, hi;
    CK(cudaDeviceGetStreamPriorityRange(&lo, &hi));
    unsigned* dout; double* ddout; unsigned long long* dlog;
    const int gridA = 148 * 2 * 10, gridB = 148 * 35;
    CK(cudaMalloc(&dout, 1024)); CK(cudaMalloc(&ddout, 1024)); CK(cudaMalloc(&dlog, gridB * 24));
    const int smA = 70 * 1024, smB = 33 * 1024;
    auto kA = nt == 256 ? alu_hog<256> : alu_hog<320>;
    CK(cudaFuncSetAttribute(kA, cudaFuncAttributeMaxDynamicSharedMemorySize, smA));
    CK(cudaFuncSetAttribute(dp_task, cudaFuncAttributeMaxDynamicSharedMemorySize, smB));
    CK(cudaFuncSetAttribute(kA, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
    CK(cudaFuncSetAttribute(dp_task, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
    cudaFuncAttributes fa; CK(cudaFuncGetAttributes(&fa, kA)); printf("A: %d thr, %d regs\n", nt, fa.numRegs);
    CK(cudaFuncGetAttributes(&fa, dp_task)); printf("B: 128 thr, %d regs\n", fa.numRegs);
    const int itA = 4000, itB = 300;
    cudaEvent_t e0, e1, e2, e3;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1)); CK(cudaEventCreate(&e2)); CK(cudaEventCreate(&e3));
    cudaStream_t sa, sb;
    for (int mode = -2; mode < 3; mode++) {
        // -2: A alone, -1: B alone, 0: B high prio, 1: A high prio, 2: equal
        CK(cudaStreamCreateWithPriority(&sa, cudaStreamNonBlocking, mode == 1 ? hi : lo));
        CK(cudaStreamCreateWithPriority(&sb, cudaStreamNonBlocking, mode == 0 ? hi : lo));
        for (int rep = 0; rep < 2; rep++) {
            CK(cudaDeviceSynchronize());
            if (mode != -1) { CK(cudaEventRecord(e0, sa)); kA<<<gridA, nt, smA, sa>>>(dout, itA); CK(cudaEventRecord(e1, sa)); }
            if (mode != -2) { CK(cudaEventRecord(e2, sb)); dp_task<<<gridB, 128, smB, sb>>>(ddout, itB, dlog, 1); CK(cudaEventRecord(e3, sb)); }
            CK(cudaDeviceSynchronize());
        }
        float ta = 0, tb = 0;
        if (mode != -1) CK(cudaEventElapsedTime(&ta, e0, e1));
        if (mode != -2) CK(cudaEventElapsedTime(&tb, e2, e3));
        printf("mode %2d: A %.3f ms  B %.3f ms", mode, ta, tb);
        if (mode != -2) {
            std::vector<unsigned long long> h(gridB * 3);
            CK(cudaMemcpy(h.data(), dlog, gridB * 24, cudaMemcpyDeviceToHost));
            unsigned long long tmin = ~0ull, tmax = 0; double dur = 0; std::vector<int> per_sm(256, 0);
            for (int i = 0; i < gridB; i++) { tmin = std::min(tmin, h[3 * i]); tmax = std::max(tmax, h[3 * i + 1]); dur += (double)(h[3 * i + 1] - h[3 * i]); per_sm[h[3 * i + 2] & 255]++; }
            // max concurrent B CTAs on SM 0: sweep
            std::vector<std::pair<unsigned long long, int>> ev;
            for (int i = 0; i < gridB; i++) if ((h[3 * i + 2] & 255) == 0) { ev.push_back({h[3 * i], 1}); ev.push_back({h[3 * i + 1], -1}); }
            std::sort(ev.begin(), ev.end());
            int curc = 0, maxc = 0; for (auto& x : ev) { curc += x.second; maxc = std::max(maxc, curc); }
            printf("  | B span %.3f ms, mean CTA %.1f us, max concurrent B CTAs on SM0 = %d", (tmax - tmin) * 1e-6, dur / gridB * 1e-3, maxc);
        }
        printf("\n");
        CK(cudaStreamDestroy(sa)); CK(cudaStreamDestroy(sb));
    }
    return 0;
}
