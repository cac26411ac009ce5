// chain_probe.cu -- what bounds the FastME chain kernel (fastme.cu: fastme_chain_kernel)?
// One CTA per frame walks nblk blocks; every block has a 2192-byte transfer table, staged through a cp.async ring;
// the walk is one dependent 16-bit shared-memory look-up per block.  Variants: full, fetch only, walk only, and
// the ring depth / batch size / thread count.   nvcc -O3 -gencode arch=compute_100a,code=sm_100a chain_probe.cu
#include <cstdio>
#include <cstdint>
#include <vector>
#include <cuda_runtime.h>
constexpr int STRIDE = 1096;
__device__ __forceinline__ void cp16(void* d, const void* s) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((uint32_t)__cvta_generic_to_shared(d)), "l"(s) : "memory");
}
template <int THREADS, int BATCH, int NBUF, int MODE>   // MODE 0 full, 1 fetch only, 2 walk only, 3 full with warp 0 not copying
__global__ void __launch_bounds__(THREADS) chain(const uint16_t* tables, int nblk, uint32_t* out) {
    extern __shared__ __align__(16) uint16_t s_tab[];
    const int tid = threadIdx.x, fl = blockIdx.x;
    const uint16_t* tab_lane = tables + (size_t)fl * nblk * STRIDE;
    const int nbatch = (nblk + BATCH - 1) / BATCH;
    auto prefetch = [&](int bt) {
        if (MODE != 2 && bt < nbatch) {
            const int nb = min(BATCH, nblk - bt * BATCH);
            const uint16_t* src = tab_lane + (size_t)bt * BATCH * STRIDE;
            uint16_t* dst = s_tab + (size_t)(bt % NBUF) * BATCH * STRIDE;
            if (MODE == 3) { for (int v = tid - 32; v < nb * (STRIDE / 8); v += THREADS - 32) cp16(dst + v * 8, src + v * 8); }
            else for (int v = tid; v < nb * (STRIDE / 8); v += THREADS) cp16(dst + v * 8, src + v * 8);
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
    };
    int q = 544;
    uint32_t* o = out + (size_t)fl * nblk;
    const bool copier = MODE != 3 || tid >= 32;
    if (copier) for (int bt = 0; bt < NBUF - 1; bt++) prefetch(bt);
    for (int bt = 0; bt < nbatch; bt++) {
        if (copier) asm volatile("cp.async.wait_group %0;" ::"n"(NBUF - 2) : "memory");
        __syncthreads();
        if (copier) prefetch(bt + NBUF - 1);
        if (tid < 32 && MODE != 1) {
            const uint16_t* tb = s_tab + (size_t)(bt % NBUF) * BATCH * STRIDE;
            const int b0 = bt * BATCH, nb = min(BATCH, nblk - b0);
            for (int i = 0; i < nb; i++, tb += STRIDE) {
                if (tid == 0) o[b0 + i] = q;
                q = tb[q] % 1089;
            }
        }
    }
    if (tid == 0) o[0] = q;
}
template <int THREADS, int BATCH, int NBUF, int MODE>
void run(const char* name, const uint16_t* d, int lanes, int nblk, uint32_t* out) {
    const size_t smem = (size_t)NBUF * BATCH * STRIDE * 2;
    cudaFuncSetAttribute(chain<THREADS, BATCH, NBUF, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    float best = 1e9;
    for (int it = 0; it < 5; it++) {
        cudaEventRecord(e0);
        chain<THREADS, BATCH, NBUF, MODE><<<lanes, THREADS, smem>>>(d, nblk, out);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1); best = ms < best ? ms : best;
    }
    printf("%-44s %7.1f us  (%.0f ns per block, %.1f GB/s per CTA)  %s\n", name, best * 1000, best * 1e6 / nblk,
           MODE == 2 ? 0.0 : nblk * STRIDE * 2.0 / best / 1e6, cudaGetErrorString(cudaGetLastError()));
}
int main() {
    const int lanes = 38, nblk = 396;
    std::vector<uint16_t> h((size_t)lanes * nblk * STRIDE);
    uint32_t x = 12345;
    for (auto& v : h) { x = x * 1664525u + 1013904223u; v = (x >> 8) % 1089; }
    uint16_t* d; uint32_t* out;
    cudaMalloc(&d, h.size() * 2); cudaMalloc(&out, (size_t)lanes * nblk * 4);
    cudaMemcpy(d, h.data(), h.size() * 2, cudaMemcpyHostToDevice);
    run<128, 16, 4, 0>("128 thr, batch 16, 4 buffers: full", d, lanes, nblk, out);
    run<128, 16, 4, 1>("128 thr, batch 16, 4 buffers: fetch only", d, lanes, nblk, out);
    run<128, 16, 4, 2>("128 thr, batch 16, 4 buffers: walk only", d, lanes, nblk, out);
    run<256, 16, 4, 0>("256 thr, batch 16, 4 buffers: full", d, lanes, nblk, out);
    run<512, 16, 4, 0>("512 thr, batch 16, 4 buffers: full", d, lanes, nblk, out);
    run<512, 16, 6, 0>("512 thr, batch 16, 6 buffers: full", d, lanes, nblk, out);
    run<512, 8, 8, 0>("512 thr, batch 8, 8 buffers: full", d, lanes, nblk, out);
    run<512, 32, 3, 0>("512 thr, batch 32, 3 buffers: full", d, lanes, nblk, out);
    run<160, 16, 4, 3>("160 thr, batch 16, 4 buffers: warp 0 walks only", d, lanes, nblk, out);
    run<288, 16, 4, 3>("288 thr, batch 16, 4 buffers: warp 0 walks only", d, lanes, nblk, out);
    run<1024, 16, 6, 1>("1024 thr, batch 16, 6 buffers: fetch only", d, lanes, nblk, out);
    return 0;
}
