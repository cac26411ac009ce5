// halfpel.cu -- K2: half-pel interpolation of a reconstructed plane.
//
// Replaces build_pre_interpolated_buffer (reference encoder/block_predictor.py:145-177):
//   P[2y][2x]     = f                      P[2y][2x+1]   = ceil((a+b)/2)
//   P[2y+1][2x]   = ceil((a+c)/2)          P[2y+1][2x+1] = ceil((a+b+c+d)/4)
// with the last row / column of P left 0.  Instead of the interleaved (2H x 2W) plane we keep four
// W x H phase planes (phase = (x&1) | (y&1)<<1): phase 0 is the plane itself, so only three planes are
// written (1 B read, 3 B written per pixel -- HBM bound).  Packed byte / 16-bit SIMD.
#include "bvc_kernels.h"

namespace bvc {
namespace {

// 16 pixels x HP_ROWS rows per thread: 128-bit loads and stores, the lower row of one output row is the upper row of the
// next (round 1 moved one 32-bit word per thread and one row per CTA: 1.9 TB/s; this one is bound by the 3 B written
// per pixel).
constexpr int HP_ROWS = 4;

__device__ __forceinline__ uint32_t bytes_mask(int n) { return n >= 4 ? 0xFFFFFFFFu : (n <= 0 ? 0u : ((1u << (8 * n)) - 1u)); }

struct Row16 { uint32_t w[5]; };   // 16 pixels + the word right of them
__device__ __forceinline__ Row16 load_row16(const uint8_t* row, int gx, int gpr_pitch) {
    Row16 r;
    const uint4 v = reinterpret_cast<const uint4*>(row)[gx];
    r.w[0] = v.x; r.w[1] = v.y; r.w[2] = v.z; r.w[3] = v.w;
    r.w[4] = (gx + 1 < gpr_pitch) ? reinterpret_cast<const uint32_t*>(row)[4 * (gx + 1)] : 0u;
    return r;
}

__global__ void __launch_bounds__(256) halfpel_kernel(const uint8_t* const* src_planes, uint8_t* const* dst_planes, int W,
                                                      int H, int pitch, size_t plane_bytes) {
    const int gpr = (W + 15) >> 4;             // 16-byte groups that hold pixels
    const int gx = blockIdx.x * blockDim.x + threadIdx.x;
    const int y0 = (blockIdx.y * blockDim.y + threadIdx.y) * HP_ROWS;
    if (gx >= gpr || y0 >= H) return;
    const int plane = blockIdx.z;
    const uint8_t* src = src_planes[plane];
    uint8_t* dst = dst_planes[plane];  // three consecutive phase planes: 1, 2, 3
    const int gpr_pitch = pitch >> 4;
    // per word: how many of its bytes are pixels (v), and how many have a right neighbour (h, g): the last column of
    // the 2x plane stays 0 (block_predictor.py:145-177), bytes between W and the pitch stay 0 as well
    uint32_t mv[4], mh[4];
#pragma unroll
    for (int j = 0; j < 4; j++) {
        const int x = 16 * gx + 4 * j;
        mv[j] = bytes_mask(W - x);
        mh[j] = bytes_mask(W - 1 - x);
    }
    Row16 A = load_row16(src + (size_t)y0 * pitch, gx, gpr_pitch);
#pragma unroll
    for (int r = 0; r < HP_ROWS; r++) {
        const int y = y0 + r;
        if (y >= H) break;
        const bool has_c = (y + 1 < H);
        Row16 C;
        if (has_c) C = load_row16(src + (size_t)(y + 1) * pitch, gx, gpr_pitch);
        else { C.w[0] = C.w[1] = C.w[2] = C.w[3] = C.w[4] = 0u; }
        uint32_t h[4], v[4], g[4];
#pragma unroll
        for (int j = 0; j < 4; j++) {
            const uint32_t a = A.w[j], c = C.w[j];
            const uint32_t b = __funnelshift_r(a, A.w[j + 1], 8);  // pixels x+1
            const uint32_t d = __funnelshift_r(c, C.w[j + 1], 8);
            h[j] = __vavgu4(a, b) & mh[j];                          // (a+b+1)>>1 per byte
            const uint32_t m = 0x00FF00FFu;
            const uint32_t se = (a & m) + (b & m) + (c & m) + (d & m) + 0x00030003u;
            const uint32_t so = ((a >> 8) & m) + ((b >> 8) & m) + ((c >> 8) & m) + ((d >> 8) & m) + 0x00030003u;
            // the last row has no lower neighbour: v and g stay 0
            v[j] = has_c ? (__vavgu4(a, c) & mv[j]) : 0u;
            g[j] = has_c ? ((((se >> 2) & m) | (((so >> 2) & m) << 8)) & mh[j]) : 0u;
        }
        uint8_t* o = dst + (size_t)y * pitch;
        reinterpret_cast<uint4*>(o)[gx] = make_uint4(h[0], h[1], h[2], h[3]);
        reinterpret_cast<uint4*>(o + plane_bytes)[gx] = make_uint4(v[0], v[1], v[2], v[3]);
        reinterpret_cast<uint4*>(o + 2 * plane_bytes)[gx] = make_uint4(g[0], g[1], g[2], g[3]);
        A = C;
    }
}

__global__ void __launch_bounds__(256) halfpel_interleave_kernel(const uint8_t* phases, int W, int H, int pitch,
                                                                 size_t plane_bytes, uint8_t* out2x) {
    const int x2 = blockIdx.x * blockDim.x + threadIdx.x, y2 = blockIdx.y;
    if (x2 >= 2 * W || y2 >= 2 * H) return;
    const int ph = (x2 & 1) | ((y2 & 1) << 1);
    out2x[(size_t)y2 * (2 * W) + x2] = phases[(size_t)ph * plane_bytes + (size_t)(y2 >> 1) * pitch + (x2 >> 1)];
}

}  // namespace

cudaError_t launch_halfpel(const uint8_t* const* src_planes, uint8_t* const* dst_planes, int nplanes, int W, int H,
                           int pitch, size_t plane_bytes, cudaStream_t st) {
    const int gpr = (W + 15) >> 4;
    const int tx = gpr >= 128 ? 128 : (gpr >= 64 ? 64 : 32), ty = 256 / tx;
    const int row_groups = (H + HP_ROWS - 1) / HP_ROWS;
    dim3 grid((gpr + tx - 1) / tx, (row_groups + ty - 1) / ty, nplanes);
    if (grid.y > 65535u) return cudaErrorInvalidValue;
    halfpel_kernel<<<grid, dim3(tx, ty), 0, st>>>(src_planes, dst_planes, W, H, pitch, plane_bytes);
    return cudaGetLastError();
}

cudaError_t launch_halfpel_interleave(const uint8_t* phases, int W, int H, int pitch, size_t plane_bytes, uint8_t* out2x,
                                      cudaStream_t st) {
    dim3 grid((2 * W + 255) / 256, 2 * H);
    halfpel_interleave_kernel<<<grid, 256, 0, st>>>(phases, W, H, pitch, plane_bytes, out2x);
    return cudaGetLastError();
}

}  // namespace bvc
