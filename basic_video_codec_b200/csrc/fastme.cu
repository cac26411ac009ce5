// fastme.cu -- K4: predictor-centred fast motion estimation.
//
// Replaces find_fast_me_block (reference encoder/block_predictor.py:11-58) and the raster-order MVP
// chain of PFrame.process_block (encoder/PFrame.py:34,44,105-110).  Per level the candidates are
//   origin (0,0), pmv_origin (mvp), top (mvp+(0,-1)), right (mvp+(1,0)), bottom (mvp+(0,1)), left (mvp+(-1,0))
// evaluated for every reference; the winner is the first strict minimum scanning references in
// ascending order and the six keys in that order.  Because of the reference's late-binding closures
// (:20-47) the *reported* reference index is always 0 while the minimum is taken over all references,
// and in iteration k the keys of refs 0..k are all (re)evaluated, so the comparison counter grows by
// nvalid * n(n+1)/2 per level.  Stop when the winner is origin/pmv_origin or |mv| >= 16 (:50-56).
//
// The MVP of a block is the MV of the previous block in raster order, so a frame is one serial
// chain: latency bound by construction.  One CTA walks one frame; each warp evaluates
// one (reference, candidate) SAD per level; independent frames (GOP lanes) run in parallel CTAs.
#include "bvc_kernels.h"

namespace bvc {
namespace {

constexpr int FM_THREADS = 512;  // 16 warps
constexpr int FM_MAXC = 6 * BVC_MAX_REFS;

__global__ void __launch_bounds__(FM_THREADS) fastme_kernel(MeArgs a, const uint8_t* ref_base, size_t ref_plane_bytes,
                                                           int ref_pitch, long long* cmp_out) {
    __shared__ int s_sad[FM_MAXC];
    __shared__ int s_mvp[2];
    __shared__ __align__(16) uint8_t s_cur[32 * 32];
    const int bs = a.bs;
    const int lane_id = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
    const int fl = blockIdx.x;
    const MeLane& L = a.lanes[fl];
    const uint8_t* curp = a.cur_base + (size_t)L.cur_plane * a.cur_plane_bytes;
    const int nref = L.nref;
    const int ncand = 6 * nref;
    long long cmp_total = 0;
    if (threadIdx.x == 0) { s_mvp[0] = 0; s_mvp[1] = 0; }  // mv_field = {(0,0): [0,0]}  (PFrame.py:34)
    __syncthreads();

    for (int b = 0; b < a.nblk; b++) {
        const int ox = (b % a.bw) * bs, oy = (b / a.bw) * bs;
        for (int i = threadIdx.x; i < bs * bs; i += blockDim.x) s_cur[i] = curp[(size_t)(oy + i / bs) * a.cur_pitch + ox + i % bs];
        __syncthreads();
        int best_sad = 0, mvx = 0, mvy = 0;
        for (;;) {
            const int px0 = s_mvp[0], py0 = s_mvp[1];
            for (int c = warp; c < ncand; c += nwarps) {
                const int k = c / 6, p = c - 6 * k;
                const int cx = p == 0 ? 0 : px0 + (p == 3) - (p == 5);
                const int cy = p == 0 ? 0 : py0 + (p == 4) - (p == 2);
                int plane = L.ref_plane[k], dx = cx, dy = cy, phx = 0, phy = 0;
                if (a.sc == 2) { phx = cx & 1; phy = cy & 1; plane += phx | (phy << 1); dx = cx >> 1; dy = cy >> 1; }
                // is_out_of_range (block_predictor.py:116-143), expressed on the phase plane
                const bool ok = (ox + dx >= 0) && (oy + dy >= 0) && (ox + dx + bs <= a.W - phx) && (oy + dy + bs <= a.H - phy);
                int s = 0;
                if (ok) {
                    const uint8_t* rp = ref_base + (size_t)plane * ref_plane_bytes + (size_t)(oy + dy) * ref_pitch + (ox + dx);
                    for (int i = lane_id; i < bs * bs; i += 32) {
                        const int y = i / bs, x = i - y * bs;
                        s += abs((int)s_cur[i] - (int)rp[(size_t)y * ref_pitch + x]);
                    }
#pragma unroll
                    for (int d = 16; d > 0; d >>= 1) s += __shfl_xor_sync(0xffffffffu, s, d);
                } else {
                    s = -1;
                }
                if (lane_id == 0) s_sad[c] = s;
            }
            __syncthreads();
            // first strict minimum in (ref ascending, key order); every thread computes it redundantly
            int best = 0x7fffffff, best_p = 0, nvalid = 0;
            for (int c = 0; c < ncand; c++) {
                const int s = s_sad[c];
                if (s < 0) continue;
                if (c < 6) nvalid++;
                if (s < best) { best = s; best_p = c % 6; }
            }
            cmp_total += (long long)nvalid * (nref * (nref + 1) / 2);
            mvx = best_p == 0 ? 0 : px0 + (best_p == 3) - (best_p == 5);
            mvy = best_p == 0 ? 0 : py0 + (best_p == 4) - (best_p == 2);
            best_sad = best;
            const bool stop = (best_p <= 1) || abs(mvx) >= 16 || abs(mvy) >= 16;
            __syncthreads();
            if (threadIdx.x == 0) { s_mvp[0] = mvx; s_mvp[1] = mvy; }  // recursion mvp / next block's mvp
            __syncthreads();
            if (stop) break;
        }
        if (threadIdx.x == 0) a.out[(size_t)fl * a.nblk + b] = make_int4(mvx, mvy, 0, best_sad);
    }
    if (threadIdx.x == 0 && cmp_out) cmp_out[fl] = cmp_total;
}

}  // namespace

cudaError_t launch_fastme(const MeArgs& a, int lanes, const uint8_t* ref_base, size_t ref_plane_bytes, int ref_pitch,
                          long long* cmp_out, cudaStream_t st) {
    if (a.bs > 32) return cudaErrorInvalidValue;
    fastme_kernel<<<lanes, FM_THREADS, 0, st>>>(a, ref_base, ref_plane_bytes, ref_pitch, cmp_out);
    return cudaGetLastError();
}

}  // namespace bvc
