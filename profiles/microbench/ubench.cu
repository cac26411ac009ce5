// Instruction-throughput microbenchmark for the SAD / transform pipes on sm_100a.
// Measures ops/clk/SM for the candidate inner-loop instructions of the motion-estimation
// kernel (VABSDIFF4.U8.ACC & friends) and the fp64 pipe.  Run: ./ubench
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

#define ITERS 4096
#define NCH 8

template <int OP>
__global__ void __launch_bounds__(256) k(uint32_t* out, uint32_t seed, unsigned long long* cyc) {
    uint32_t a[NCH], acc[NCH];
    uint32_t b = seed ^ threadIdx.x * 2654435761u;
#pragma unroll
    for (int j = 0; j < NCH; j++) { a[j] = b + j * 0x01020304u; acc[j] = j; }
    double d[NCH]; 
#pragma unroll
    for (int j = 0; j < NCH; j++) d[j] = (double)(b & 1023) * 1e-3 + j;
    double dm = 1.0000001, da = 1e-9;
    __syncthreads();
    long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < ITERS; it++) {
#pragma unroll
        for (int j = 0; j < NCH; j++) {
            if (OP == 0) asm volatile("vabsdiff4.u32.u32.u32.add %0, %1, %2, %0;" : "+r"(acc[j]) : "r"(a[j]), "r"(b));
            if (OP == 1) asm volatile("vabsdiff4.u32.u32.u32 %0, %1, %2, %0;" : "+r"(acc[j]) : "r"(a[j]), "r"(b));
            if (OP == 2) asm volatile("sad.u32 %0, %1, %2, %0;" : "+r"(acc[j]) : "r"(a[j]), "r"(b));
            if (OP == 3) asm volatile("add.u32 %0, %0, %1;" : "+r"(acc[j]) : "r"(a[j]));
            if (OP == 4) asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(acc[j]) : "r"(a[j]), "r"(b));
            if (OP == 5) asm volatile("prmt.b32 %0, %0, %1, %2;" : "+r"(acc[j]) : "r"(a[j]), "r"(b));
            if (OP == 6) asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(acc[j]) : "r"(a[j]), "r"(b));
            if (OP == 7) asm volatile("fma.rn.f64 %0, %0, %1, %2;" : "+d"(d[j]) : "d"(dm), "d"(da));
            if (OP == 8) asm volatile("add.rn.f64 %0, %0, %1;" : "+d"(d[j]) : "d"(da));
            if (OP == 9) { // vabsdiff4.add + independent IMAD (fma pipe) interleaved: do they dual-issue?
                asm volatile("vabsdiff4.u32.u32.u32.add %0, %1, %2, %0;" : "+r"(acc[j]) : "r"(a[j]), "r"(b));
                asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(a[j]) : "r"(b), "r"(seed));
            }
            if (OP == 10) { // vabsdiff4.add + PRMT (alu pipe) interleaved
                asm volatile("vabsdiff4.u32.u32.u32.add %0, %1, %2, %0;" : "+r"(acc[j]) : "r"(a[j]), "r"(b));
                asm volatile("prmt.b32 %0, %0, %1, %2;" : "+r"(a[j]) : "r"(b), "r"(seed));
            }
            if (OP == 11) asm volatile("shf.r.wrap.b32 %0, %0, %1, %2;" : "+r"(acc[j]) : "r"(a[j]), "r"(seed));
            if (OP == 12) asm volatile("vsub4.u32.u32.u32.sat %0, %1, %2, %0;" : "+r"(acc[j]) : "r"(a[j]), "r"(b));
            if (OP == 13) { // dp4a
                asm volatile("dp4a.u32.u32 %0, %1, %2, %0;" : "+r"(acc[j]) : "r"(a[j]), "r"(b));
            }
            if (OP == 14) { // vabsdiff4.add + shf (funnel shift) interleaved
                asm volatile("vabsdiff4.u32.u32.u32.add %0, %1, %2, %0;" : "+r"(acc[j]) : "r"(a[j]), "r"(b));
                asm volatile("shf.r.wrap.b32 %0, %0, %1, %2;" : "+r"(a[j]) : "r"(b), "r"(seed));
            }
            if (OP == 15) asm volatile("vabsdiff2.u32.u32.u32.add %0, %1, %2, %0;" : "+r"(acc[j]) : "r"(a[j]), "r"(b));
        }
    }
    long long t1 = clock64();
    uint32_t s = 0; double ds = 0;
#pragma unroll
    for (int j = 0; j < NCH; j++) { s += acc[j] + a[j]; ds += d[j]; }
    out[blockIdx.x * blockDim.x + threadIdx.x] = s + (uint32_t)ds;
    if (threadIdx.x == 0) cyc[blockIdx.x] = (unsigned long long)(t1 - t0);
}

template <int OP>
void run(const char* name, int ops_per_iter, int ctas_per_sm) {
    int nsm; cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, 0);
    int grid = nsm * ctas_per_sm;
    uint32_t* out; unsigned long long* cyc;
    cudaMalloc(&out, grid * 256 * 4); cudaMalloc(&cyc, grid * 8);
    k<OP><<<grid, 256>>>(out, 123, cyc);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    k<OP><<<grid, 256>>>(out, 123, cyc);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    unsigned long long* h = (unsigned long long*)malloc(grid * 8);
    cudaMemcpy(h, cyc, grid * 8, cudaMemcpyDeviceToHost);
    double avg = 0; for (int i = 0; i < grid; i++) avg += h[i]; avg /= grid;
    double ops_per_sm = (double)ctas_per_sm * 256.0 * ITERS * NCH * ops_per_iter;
    printf("%-34s ctas/SM=%d  cycles/CTA=%.0f  thread-ops/clk/SM=%.2f  wall=%.3f ms  total=%.3e thread-ops/s err=%s\n",
           name, ctas_per_sm, avg, ops_per_sm / avg, ms, ops_per_sm * nsm / (ms * 1e-3), cudaGetErrorString(cudaGetLastError()));
    cudaFree(out); cudaFree(cyc); free(h);
}

int main() {
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    printf("device %s SMs=%d clock=%d kHz\n", p.name, p.multiProcessorCount, p.clockRate);
    for (int c = 4; c <= 8; c += 4) {
        run<0>("vabsdiff4.add (VABSDIFF4.U8.ACC)", 1, c);
        run<1>("vabsdiff4 (no acc)", 1, c);
        run<2>("sad.u32", 1, c);
        run<3>("add.u32 (IADD3)", 1, c);
        run<4>("lop3", 1, c);
        run<5>("prmt", 1, c);
        run<6>("mad.lo.u32 (IMAD)", 1, c);
        run<7>("fma.rn.f64 (DFMA)", 1, c);
        run<8>("add.rn.f64 (DADD)", 1, c);
        run<9>("vabsdiff4.add + IMAD", 2, c);
        run<10>("vabsdiff4.add + PRMT", 2, c);
        run<11>("shf.r.wrap", 1, c);
        run<12>("vsub4.sat", 1, c);
        run<13>("dp4a", 1, c);
        run<14>("vabsdiff4.add + SHF", 2, c);
        run<15>("vabsdiff2.add", 1, c);
    }
    return 0;
}
