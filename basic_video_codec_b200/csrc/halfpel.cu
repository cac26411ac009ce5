// halfpel.cu -- K2: half-pel interpolation of a reconstructed plane.
//
// Replaces build_pre_interpolated_buffer (reference encoder/block_predictor.py:145-177):
//   P[2y][2x]     = f                      P[2y][2x+1]   = ceil((a+b)/2)
//   P[2y+1][2x]   = ceil((a+c)/2)          P[2y+1][2x+1] = ceil((a+b+c+d)/4)
// with the last row / column of P left 0.  Instead of the interleaved (2H x 2W) plane we keep four
// W x H phase planes (phase = (x&1) | (y&1)<<1): phase 0 is the plane itself, so only three planes are
// written (1 B read, 3 B written per pixel -- HBM bound).  Four pixels per thread, packed 16-bit SIMD.
#include "bvc_kernels.h"

namespace bvc {
namespace {

__global__ void __launch_bounds__(256) halfpel_kernel(const uint8_t* const* src_planes, uint8_t* const* dst_planes, int W,
                                                      int H, int pitch, size_t plane_bytes) {
    const int wpr = W >> 2;  // words per row (W is a multiple of 4)
    const int plane = blockIdx.z;
    const uint8_t* src = src_planes[plane];
    uint8_t* dst = dst_planes[plane];  // three consecutive phase planes: 1, 2, 3
    for (int y = blockIdx.y; y < H; y += gridDim.y) {
        const uint32_t* r0 = reinterpret_cast<const uint32_t*>(src + (size_t)y * pitch);
        const uint32_t* r1 = reinterpret_cast<const uint32_t*>(src + (size_t)(y + 1) * pitch);
        const bool has_c = (y + 1 < H);
        for (int wx = blockIdx.x * blockDim.x + threadIdx.x; wx < wpr; wx += gridDim.x * blockDim.x) {
            const uint32_t a = r0[wx];
            const uint32_t an = (wx + 1 < wpr) ? r0[wx + 1] : 0u;
            const uint32_t b = __funnelshift_r(a, an, 8);  // pixels x+1
            const uint32_t c = has_c ? r1[wx] : 0u;
            const uint32_t cn = (has_c && wx + 1 < wpr) ? r1[wx + 1] : 0u;
            const uint32_t d = __funnelshift_r(c, cn, 8);
            uint32_t h = __vavgu4(a, b);  // (a+b+1)>>1 per byte
            uint32_t v = __vavgu4(a, c);
            const uint32_t m = 0x00FF00FFu;
            const uint32_t se = (a & m) + (b & m) + (c & m) + (d & m) + 0x00030003u;
            const uint32_t so = ((a >> 8) & m) + ((b >> 8) & m) + ((c >> 8) & m) + ((d >> 8) & m) + 0x00030003u;
            uint32_t g = ((se >> 2) & m) | (((so >> 2) & m) << 8);
            // the last column has no right neighbour, the last row no lower one: stay 0
            if (wx + 1 == wpr) { h &= 0x00FFFFFFu; g &= 0x00FFFFFFu; }
            if (!has_c) { v = 0; g = 0; }
            reinterpret_cast<uint32_t*>(dst + (size_t)y * pitch)[wx] = h;
            reinterpret_cast<uint32_t*>(dst + plane_bytes + (size_t)y * pitch)[wx] = v;
            reinterpret_cast<uint32_t*>(dst + 2 * plane_bytes + (size_t)y * pitch)[wx] = g;
        }
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
    const int wpr = W >> 2;
    dim3 grid((wpr + 255) / 256, H < 1024 ? H : 1024, nplanes);
    halfpel_kernel<<<grid, 256, 0, st>>>(src_planes, dst_planes, W, H, pitch, plane_bytes);
    return cudaGetLastError();
}

cudaError_t launch_halfpel_interleave(const uint8_t* phases, int W, int H, int pitch, size_t plane_bytes, uint8_t* out2x,
                                      cudaStream_t st) {
    dim3 grid((2 * W + 255) / 256, 2 * H);
    halfpel_interleave_kernel<<<grid, 256, 0, st>>>(phases, W, H, pitch, plane_bytes, out2x);
    return cudaGetLastError();
}

}  // namespace bvc
