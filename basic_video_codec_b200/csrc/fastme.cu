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

// ---------------------------------------------------------------------------------------------
// FastME on a SAD map.  The candidates of every level lie within one MV unit of the running predictor, and the walk
// stops once a component reaches 16, so nearly every SAD it can ask for lies within +-16 MV units of the block.  The
// tiled full-search kernel computes all of those at VABSDIFF4 speed (a.sad_map, radius a.R plane units); the serial MVP
// chain of a frame is then one warp doing table look-ups: the table of the next block is fetched into shared memory
// with cp.async while the current block is walked (double buffer), lane c takes candidate c = 6*ref + key, and the
// first strict minimum in (ref, key) order is a warp min over (SAD << 8 | c).  A candidate outside the map (possible
// when the predictor has drifted beyond 16) is evaluated directly by the whole warp.  Same results, same comparison
// count as fastme_kernel above.  grid = lanes, one warp each; dynamic smem = 2 * nref_max * nphase * map_stride * 2 B.
__device__ __forceinline__ void cp_async16(void* dst, const void* src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(dst)), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }

__global__ void __launch_bounds__(32) fastme_walk_kernel(MeArgs a, const uint8_t* ref_base, size_t ref_plane_bytes, int ref_pitch,
                                                       long long* cmp_out) {
    extern __shared__ __align__(16) uint16_t s_map[];   // [2][nref][nphase][map_stride]
    const int bs = a.bs, lane = threadIdx.x, fl = blockIdx.x;
    const MeLane& L = a.lanes[fl];
    const uint8_t* curp = a.cur_base + (size_t)L.cur_plane * a.cur_plane_bytes;
    const int nref = L.nref, ncand = 6 * nref, R = a.R, n1 = 2 * R + 1;
    const int chunk = a.map_stride;                       // elements per (ref, phase) of one block
    const int buf_elems = a.max_refs * a.nphase * chunk;
    const int vec_per_chunk = chunk / 8, nchunks = nref * a.nphase;
    const size_t blk_stride = (size_t)chunk, kp_stride = (size_t)a.nblk * chunk;
    const uint16_t* map_lane = a.sad_map + (size_t)fl * a.max_refs * a.nphase * kp_stride;
    auto prefetch = [&](int b, int slot) {
        // chunks of block b: (ref k, phase ph) = kp -> s_map[slot][kp][*]; 16 bytes per lane per step
        const uint16_t* src = map_lane + (size_t)b * blk_stride;
        uint16_t* dst = s_map + (size_t)slot * buf_elems;
        for (int kp = 0; kp < nchunks; kp++, src += kp_stride, dst += chunk)
            for (int v = lane; v < vec_per_chunk; v += 32) cp_async16(dst + v * 8, src + v * 8);
        cp_async_commit();
    };
    // per-lane candidate constants: lane handles candidates c = lane (and lane + 32 when there are more than 32)
    int ck[2], cdx[2], cdy[2];
    bool con[2], corg[2], cfirst[2];
#pragma unroll
    for (int j = 0; j < 2; j++) {
        const int c = lane + 32 * j, p = c % 6;
        con[j] = c < ncand;
        ck[j] = con[j] ? c / 6 : 0;
        corg[j] = p == 0;
        cdx[j] = (p == 3) - (p == 5);
        cdy[j] = (p == 4) - (p == 2);
        cfirst[j] = c < 6;
    }
    const int tri = nref * (nref + 1) / 2;
    long long cmp_total = 0;
    int mvpx = 0, mvpy = 0;   // mv_field = {(0,0): [0,0]}  (PFrame.py:34); carried from block to block
    prefetch(0, 0);
    for (int b = 0; b < a.nblk; b++) {
        const int ox = (b % a.bw) * bs, oy = (b / a.bw) * bs;
        cp_async_wait_all();
        __syncwarp();
        if (b + 1 < a.nblk) prefetch(b + 1, (b + 1) & 1);
        const uint16_t* tab = s_map + (size_t)(b & 1) * buf_elems;
        int mvx = 0, mvy = 0, best_sad = 0;
        for (;;) {
            uint32_t key = 0xffffffffu;
            int nvalid = 0;
#pragma unroll
            for (int j = 0; j < 2; j++) {
                if (j == 1 && ncand <= 32) break;   // warp-uniform
                const bool on = con[j];
                const int k = ck[j];
                const int cx = corg[j] ? 0 : mvpx + cdx[j];
                const int cy = corg[j] ? 0 : mvpy + cdy[j];
                int ph = 0, px = 0, py = 0, dx = cx, dy = cy;
                if (a.sc == 2) { px = cx & 1; py = cy & 1; ph = px | (py << 1); dx = cx >> 1; dy = cy >> 1; }
                const bool inmap = dx >= -R && dy >= -R && dx <= R - px && dy <= R - py;
                int s = -1;
                if (on && inmap) {
                    const uint16_t v = tab[(k * a.nphase + ph) * chunk + (dy + R) * n1 + (dx + R)];
                    s = v == 0xffffu ? -1 : (int)v;
                }
                // candidates outside the map: the warp evaluates them one by one (is_out_of_range block_predictor.py:116-143)
                uint32_t need = __ballot_sync(0xffffffffu, on && !inmap);
                while (need) {
                    const int src = __ffs(need) - 1;
                    need &= need - 1;
                    const int kk = __shfl_sync(0xffffffffu, k, src), pp = __shfl_sync(0xffffffffu, ph, src);
                    const int ddx = __shfl_sync(0xffffffffu, dx, src), ddy = __shfl_sync(0xffffffffu, dy, src);
                    const int phx = pp & 1, phy = pp >> 1;
                    const bool ok = (ox + ddx >= 0) && (oy + ddy >= 0) && (ox + ddx + bs <= a.W - phx) && (oy + ddy + bs <= a.H - phy);
                    int t = -1;
                    if (ok) {
                        const uint8_t* rp = ref_base + (size_t)(L.ref_plane[kk] + pp) * ref_plane_bytes + (size_t)(oy + ddy) * ref_pitch + (ox + ddx);
                        t = 0;
                        for (int i = lane; i < bs * bs; i += 32) {
                            const int y = i / bs, x = i - y * bs;
                            t += abs((int)curp[(size_t)(oy + y) * a.cur_pitch + ox + x] - (int)rp[(size_t)y * ref_pitch + x]);
                        }
#pragma unroll
                        for (int d = 16; d > 0; d >>= 1) t += __shfl_xor_sync(0xffffffffu, t, d);
                    }
                    if (lane == src) s = t;
                }
                if (on && s >= 0) key = min(key, ((uint32_t)s << 8) | (uint32_t)(lane + 32 * j));
                nvalid += __popc(__ballot_sync(0xffffffffu, on && s >= 0 && cfirst[j]));
            }
#pragma unroll
            for (int d = 16; d > 0; d >>= 1) key = min(key, __shfl_xor_sync(0xffffffffu, key, d));
            cmp_total += (long long)(nvalid * tri);
            const int best_p = (key == 0xffffffffu) ? 0 : (int)(key & 255u) % 6;
            best_sad = (key == 0xffffffffu) ? 0x7fffffff : (int)(key >> 8);
            mvx = best_p == 0 ? 0 : mvpx + (best_p == 3) - (best_p == 5);
            mvy = best_p == 0 ? 0 : mvpy + (best_p == 4) - (best_p == 2);
            const bool stop = (best_p <= 1) || abs(mvx) >= 16 || abs(mvy) >= 16;
            mvpx = mvx;
            mvpy = mvy;
            if (stop) break;
        }
        if (lane == 0) a.out[(size_t)fl * a.nblk + b] = make_int4(mvx, mvy, 0, best_sad);
        __syncwarp();   // everyone is done with this block's table before the prefetch after next overwrites it
    }
    if (lane == 0 && cmp_out) cmp_out[fl] = cmp_total;
}

}  // namespace

cudaError_t launch_fastme(const MeArgs& a, int lanes, const uint8_t* ref_base, size_t ref_plane_bytes, int ref_pitch,
                          long long* cmp_out, cudaStream_t st) {
    if (a.bs > 32) return cudaErrorInvalidValue;
    fastme_kernel<<<lanes, FM_THREADS, 0, st>>>(a, ref_base, ref_plane_bytes, ref_pitch, cmp_out);
    return cudaGetLastError();
}

cudaError_t launch_fastme_walk(const MeArgs& a, int lanes, const uint8_t* ref_base, size_t ref_plane_bytes, int ref_pitch,
                               long long* cmp_out, cudaStream_t st) {
    if (!a.sad_map || a.bs > 32 || a.map_stride % 8) return cudaErrorInvalidValue;
    const size_t smem = 2 * (size_t)a.max_refs * a.nphase * a.map_stride * sizeof(uint16_t);
    static size_t configured_dev[BVC_MAX_DEVICES] = {};
    size_t& configured = configured_dev[current_device_slot()];
    if (smem > configured) {
        cudaError_t e = cudaFuncSetAttribute(fastme_walk_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        configured = smem;
    }
    fastme_walk_kernel<<<lanes, 32, smem, st>>>(a, ref_base, ref_plane_bytes, ref_pitch, cmp_out);
    return cudaGetLastError();
}

}  // namespace bvc
