// bvc_common.cuh -- shared device helpers (sm_100a): TMA / mbarrier PTX wrappers, byte-SIMD SAD,
// exp-Golomb helpers, and the host-side structures shared by the kernels and the C-ABI layer.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

#define BVC_EOB_MARKER 8190  // Frame.EOB_MARKER (reference encoder/Frame.py:23)
#define BVC_MAX_REFS 8       // deque(maxlen=nRefFrames) window the kernels support
#define BVC_SAD_POISON (1u << 22)

namespace bvc {

// cudaFuncSetAttribute is per device: launchers remember what they configured for each device of the process, so a
// process that drives several GPUs (one context each) configures every kernel on every device.
constexpr int BVC_MAX_DEVICES = 64;
inline int current_device_slot() {
    int d = 0;
    cudaGetDevice(&d);
    return (d >= 0 && d < BVC_MAX_DEVICES) ? d : 0;
}

// ---------------------------------------------------------------------------------------------
// Geometry of one stream of frames (all planes of a context share it).
struct Geom {
    int W, H;      // padded luma size, multiples of bs
    int pitch;     // bytes per plane row on the device (multiple of 16, for TMA)
    int bs;        // block size i
    int bw, bh;    // blocks per row / rows of blocks
    int nblk;      // bw*bh
    size_t plane_bytes;  // pitch*H rounded up to 256
};

// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_mbar_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void fence_proxy_async_smem() {
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "BVC_WAIT:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra BVC_DONE;\n"
        "bra BVC_WAIT;\n"
        "BVC_DONE:\n"
        "}\n" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}
// TMA: 3-D tiled tensor load (x = byte column, y = row, z = plane) global -> shared, completion on mbarrier.
// Out-of-bounds elements (negative or past-the-end coordinates) are zero-filled by the hardware.
__device__ __forceinline__ void tma_load_3d(void* dst, const CUtensorMap* map, uint64_t* bar, int x, int y, int z) {
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
        ::"r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar)), "r"(x), "r"(y), "r"(z)
        : "memory");
}
// the same, pinned in program order relative to volatile loads (software-pipelined search bodies)
__device__ __forceinline__ uint32_t sad4_keep(uint32_t a, uint32_t b, uint32_t acc) {
    uint32_t d;
    asm volatile("vabsdiff4.u32.u32.u32.add %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(acc));
    return d;
}
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* map) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(map) : "memory");
}

// acc + sum_i |a.b[i] - b.b[i]| over the four packed bytes: one VABSDIFF4.U8.ACC (ALU pipe, 16 lanes/clk/SMSP)
__device__ __forceinline__ uint32_t sad4(uint32_t a, uint32_t b, uint32_t acc) {
    uint32_t d;
    asm("vabsdiff4.u32.u32.u32.add %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(acc));
    return d;
}

// Wavefront kernels (one warp per CTA): the CTA's position in the wavefront is the order in which it started.  The last
// CTA of the grid puts the counter back to zero for the next launch on the stream.
__device__ __forceinline__ int wavefront_ticket(int* counter, int lane) {
    int tk = 0;
    if (lane == 0) {
        tk = atomicAdd(counter, 1);
        if (tk == (int)gridDim.x - 1) atomicExch(counter, 0);
    }
    return __shfl_sync(0xffffffffu, tk, 0);
}

// signed exp-Golomb (reference encoder/entropy_encoder.py:8-29): code value e = map(v)+1 written in
// nb = bitlen(e) bits preceded by nb-1 zeros  => total 2*nb-1 bits, numerically just `e`.
__device__ __forceinline__ uint32_t eg_code(int v) { return (v <= 0 ? (uint32_t)(-2 * v) : (uint32_t)(2 * v - 1)) + 1u; }
__device__ __forceinline__ int eg_len_of_code(uint32_t e) { return 2 * (32 - __clz(e)) - 1; }

}  // namespace bvc
