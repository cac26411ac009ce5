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
// tiled full-search kernel computes all of those at VABSDIFF4 speed (a.sad_map, radius a.R plane units).  What is left
// of find_fast_me_block is table look-ups, organised in one of two ways:
//
//   serial walk (fastme_walk_kernel, bvc_set_fastme_direct(ctx, 2)): one warp per frame walks the blocks in raster
//   order; the table of the next block is fetched into shared memory with cp.async while the current one is walked.
//
//   transfer table (default): the walk of a block is a pure function  F_b : predictor -> vector  of the block's SAD
//   map, and one level of it is a pure function  step_b : predictor -> (predictor' | stop).  fastme_table_kernel (one
//   CTA per block) evaluates step_b for all 31 x 31 predictors with components in [-15, 15] (their levels only touch
//   the map: a component of +-16 stops the walk), then follows the step pointers to the fixed point, which gives F_b for
//   every such predictor -- all blocks and predictors in parallel.  The serial MVP chain of a frame
//   (mvp_{b+1} = F_b(mvp_b), PFrame.py:34,44,105-110) is then one 2-byte look-up per block (fastme_chain_kernel, tables
//   staged through a cp.async ring); a predictor outside the table (drifted past 15) is walked on the spot.  Finally
//   fastme_finish_kernel (one warp per block, all blocks in parallel) replays each block's walk from its now known
//   predictor for the vector, its SAD and the comparison count.
//
// A candidate outside the map is evaluated directly by the warp.  Same results and comparison count as fastme_kernel.
__device__ __forceinline__ void cp_async16(void* dst, const void* src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(dst)), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// per-lane candidate constants: a lane handles candidates c = lane (and lane + 32 when there are more than 32);
// c = 6 * reference + key, keys in the order origin, pmv_origin, top, right, bottom, left (block_predictor.py:20-47)
struct WalkLane {
    int ck[2], cdx[2], cdy[2];
    bool con[2], corg[2], cfirst[2];
    __device__ __forceinline__ void init(int lane, int ncand) {
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
    }
};

// The walk of one block (find_fast_me_block) by one warp, starting from predictor (mvpx, mvpy).  tab = the block's SAD
// map, chunk (reference k, phase ph) at tab + (k * nphase + ph) * kp_stride (shared or global memory).
// Returns the vector in (mvpx, mvpy), its SAD, and adds the comparisons to cmp_total.
__device__ __forceinline__ void walk_block(const MeArgs& a, const MeLane& L, const WalkLane& wl, const uint8_t* curp,
                                           const uint8_t* ref_base, size_t ref_plane_bytes, int ref_pitch,
                                           const uint16_t* tab, size_t kp_stride, int b, int lane, int& mvpx, int& mvpy,
                                           int& best_sad, long long& cmp_total) {
    const int bs = a.bs, R = a.R, n1 = 2 * R + 1;
    const int nref = L.nref, ncand = 6 * nref, tri = nref * (nref + 1) / 2;
    const int ox = (b % a.bw) * bs, oy = (b / a.bw) * bs;
    for (;;) {
        uint32_t key = 0xffffffffu;
        int nvalid = 0;
#pragma unroll
        for (int j = 0; j < 2; j++) {
            if (j == 1 && ncand <= 32) break;   // warp-uniform
            const bool on = wl.con[j];
            const int k = wl.ck[j];
            const int cx = wl.corg[j] ? 0 : mvpx + wl.cdx[j];
            const int cy = wl.corg[j] ? 0 : mvpy + wl.cdy[j];
            int ph = 0, px = 0, py = 0, dx = cx, dy = cy;
            if (a.sc == 2) { px = cx & 1; py = cy & 1; ph = px | (py << 1); dx = cx >> 1; dy = cy >> 1; }
            const bool inmap = dx >= -R && dy >= -R && dx <= R - px && dy <= R - py;
            int s = -1;
            if (on && inmap) {
                const uint16_t v = tab[(size_t)(k * a.nphase + ph) * kp_stride + (dy + R) * n1 + (dx + R)];
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
            nvalid += __popc(__ballot_sync(0xffffffffu, on && s >= 0 && wl.cfirst[j]));
        }
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) key = min(key, __shfl_xor_sync(0xffffffffu, key, d));
        cmp_total += (long long)(nvalid * tri);
        const int best_p = (key == 0xffffffffu) ? 0 : (int)(key & 255u) % 6;
        best_sad = (key == 0xffffffffu) ? 0x7fffffff : (int)(key >> 8);
        const int mvx = best_p == 0 ? 0 : mvpx + (best_p == 3) - (best_p == 5);
        const int mvy = best_p == 0 ? 0 : mvpy + (best_p == 4) - (best_p == 2);
        const bool stop = (best_p <= 1) || abs(mvx) >= 16 || abs(mvy) >= 16;
        mvpx = mvx;
        mvpy = mvy;
        if (stop) break;
    }
}

__global__ void __launch_bounds__(32) fastme_walk_kernel(MeArgs a, const uint8_t* ref_base, size_t ref_plane_bytes, int ref_pitch,
                                                       long long* cmp_out) {
    extern __shared__ __align__(16) uint16_t s_map[];   // [2][nref][nphase][map_stride]
    const int lane = threadIdx.x, fl = blockIdx.x;
    const MeLane& L = a.lanes[fl];
    const uint8_t* curp = a.cur_base + (size_t)L.cur_plane * a.cur_plane_bytes;
    const int nref = L.nref;
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
    WalkLane wl;
    wl.init(lane, 6 * nref);
    long long cmp_total = 0;
    int mvpx = 0, mvpy = 0;   // mv_field = {(0,0): [0,0]}  (PFrame.py:34); carried from block to block
    prefetch(0, 0);
    for (int b = 0; b < a.nblk; b++) {
        cp_async_wait_all();
        __syncwarp();
        if (b + 1 < a.nblk) prefetch(b + 1, (b + 1) & 1);
        const uint16_t* tab = s_map + (size_t)(b & 1) * buf_elems;
        int best_sad = 0;
        walk_block(a, L, wl, curp, ref_base, ref_plane_bytes, ref_pitch, tab, (size_t)chunk, b, lane, mvpx, mvpy, best_sad, cmp_total);
        if (lane == 0) a.out[(size_t)fl * a.nblk + b] = make_int4(mvpx, mvpy, 0, best_sad);
        __syncwarp();   // everyone is done with this block's table before the prefetch after next overwrites it
    }
    if (lane == 0 && cmp_out) cmp_out[fl] = cmp_total;
}

// ---- transfer-table path ---------------------------------------------------------------------------------------
// Predictors and vectors with components within +-16 are numbered q = (y + 16) * 33 + (x + 16).  tables[lane][block][q]
// (uint16) = the q of the vector the block's walk returns when it starts from predictor q; 0xFFFF for the border predictors
// (a component of +-16), whose first level would look outside the SAD map: those are walked on the spot, as are predictors
// that have drifted further out.
constexpr int FT_S = 15;                      // tabulated predictors: components within +-FT_S
constexpr int FT_P = 2 * (FT_S + 1) + 1;      // 33 positions per axis
constexpr int FT_Q = FT_P * FT_P;             // 1089
constexpr int FT_STRIDE = 1096;               // entries per block, 2192 B = 137 x 16 B
constexpr int FT_CENTER = (FT_S + 1) * FT_P + (FT_S + 1);
constexpr int FT_THREADS = 256;
constexpr uint16_t FT_STOP = 1u << 12, FT_NONE = 0xffffu;
__device__ __forceinline__ int ft_q(int x, int y) { return (y + FT_S + 1) * FT_P + (x + FT_S + 1); }
__device__ __forceinline__ int ft_x(int q) { return q % FT_P - (FT_S + 1); }
__device__ __forceinline__ int ft_y(int q) { return q / FT_P - (FT_S + 1); }
// per-block predictor record written by the chain kernel: q, or bit 31 | (x + 16384) | (y + 16384) << 15 outside the grid
__device__ __forceinline__ uint32_t ft_far(int x, int y) { return 0x80000000u | (uint32_t)(x + 16384) | ((uint32_t)(y + 16384) << 15); }
__device__ __forceinline__ void ft_unpack(uint32_t e, int& x, int& y) {
    if (e & 0x80000000u) { x = (int)(e & 32767u) - 16384; y = (int)((e >> 15) & 32767u) - 16384; }
    else { x = ft_x((int)e); y = ft_y((int)e); }
}

// grid = (nblk, lanes).  Phase 1: for every MV position within +-16 the best (SAD, reference) over the references -- the
// first strict minimum over (reference ascending, key order) is the lexicographic minimum of (SAD, reference, key), so the
// references can be reduced per position first.  Phase 2: one level from every predictor (origin + 5 positions).
// Phase 3: follow the step pointers to the fixed point.
__global__ void __launch_bounds__(FT_THREADS) fastme_table_kernel(MeArgs a, uint16_t* tables) {
    __shared__ uint32_t s_best[FT_STRIDE];     // (SAD << 8 | reference) of every position, 0xffffffff = leaves the plane
    __shared__ uint16_t s_step[FT_Q];          // q of the next predictor | FT_STOP
    const int b = blockIdx.x, fl = blockIdx.y, tid = threadIdx.x;
    const MeLane& L = a.lanes[fl];
    const int nref = L.nref, R = a.R, n1 = 2 * R + 1, chunk = a.map_stride, nphase = a.nphase;
    const size_t kp_stride = (size_t)a.nblk * chunk;
    const uint16_t* src = a.sad_map + (size_t)fl * a.max_refs * nphase * kp_stride + (size_t)b * chunk;
    if (a.sc == 1 && chunk == FT_STRIDE) {
        // integer-pel: the map of a reference *is* the position grid (n1 = 33); 8 positions per 16-byte load, all
        // references requested before the first one is used
        for (int v = tid; v < FT_STRIDE / 8; v += FT_THREADS) {
            uint4 m[BVC_MAX_REFS];
#pragma unroll
            for (int k = 0; k < BVC_MAX_REFS; k++)
                if (k < nref) m[k] = __ldg(reinterpret_cast<const uint4*>(src + (size_t)k * kp_stride) + v);
#pragma unroll
            for (int j = 0; j < 8; j++) {
                uint32_t best = 0xffffffffu;
#pragma unroll
                for (int k = 0; k < BVC_MAX_REFS; k++) {
                    if (k < nref) {
                        const uint32_t w = j < 2 ? m[k].x : j < 4 ? m[k].y : j < 6 ? m[k].z : m[k].w;
                        const uint32_t val = (j & 1) ? w >> 16 : w & 0xffffu;
                        if (val != 0xffffu) best = min(best, (val << 8) | (uint32_t)k);
                    }
                }
                s_best[8 * v + j] = best;
            }
        }
    } else {
#pragma unroll 1
        for (int q = tid; q < FT_Q; q += FT_THREADS) {
            const int cy = ft_y(q), cx = ft_x(q);
            int ph = 0, dx = cx, dy = cy;
            if (a.sc == 2) { ph = (cx & 1) | ((cy & 1) << 1); dx = cx >> 1; dy = cy >> 1; }
            const uint16_t* pos = src + (size_t)ph * kp_stride + (dy + R) * n1 + (dx + R);   // always inside the map
            uint32_t val[BVC_MAX_REFS];
#pragma unroll
            for (int k = 0; k < BVC_MAX_REFS; k++)
                if (k < nref) val[k] = __ldg(pos + (size_t)k * nphase * kp_stride);
            uint32_t best = 0xffffffffu;
#pragma unroll
            for (int k = 0; k < BVC_MAX_REFS; k++)
                if (k < nref && val[k] != 0xffffu) best = min(best, (val[k] << 8) | (uint32_t)k);
            s_best[q] = best;
        }
    }
    __syncthreads();
    const uint32_t korg = s_best[FT_CENTER];
    for (int q = tid; q < FT_Q; q += FT_THREADS) {
        const int py0 = ft_y(q), px0 = ft_x(q);
        if (abs(px0) > FT_S || abs(py0) > FT_S) { s_step[q] = FT_NONE; continue; }
        // key = (SAD, reference, key index); an invalid position stays above every valid one
        auto mk = [](uint32_t v, uint32_t p) { return v == 0xffffffffu ? v : (v << 3) | p; };   // v < 2^24
        uint32_t key = mk(korg, 0);
        key = min(key, mk(s_best[q], 1));
        key = min(key, mk(s_best[q - FT_P], 2));
        key = min(key, mk(s_best[q + 1], 3));
        key = min(key, mk(s_best[q + FT_P], 4));
        key = min(key, mk(s_best[q - 1], 5));
        const int best_p = (key == 0xffffffffu) ? 0 : (int)(key & 7u);
        const int mvx = best_p == 0 ? 0 : px0 + (best_p == 3) - (best_p == 5);
        const int mvy = best_p == 0 ? 0 : py0 + (best_p == 4) - (best_p == 2);
        const bool stop = (best_p <= 1) || abs(mvx) >= 16 || abs(mvy) >= 16;
        s_step[q] = (uint16_t)(ft_q(mvx, mvy) | (stop ? FT_STOP : 0));
    }
    __syncthreads();
    // follow the step pointers to the fixed point (every move lowers (SAD, reference), so there are no cycles; a move
    // that does not stop lands on a tabulated predictor)
    uint16_t* out = tables + ((size_t)fl * a.nblk + b) * FT_STRIDE;
    for (int q = tid; q < FT_Q; q += FT_THREADS) {
        uint16_t e = s_step[q];
        if (e != FT_NONE) {
            for (int guard = 0; !(e & FT_STOP) && guard < FT_Q; guard++) e = s_step[e];
            e &= FT_STOP - 1;
        }
        out[q] = e;
    }
}

// grid = lanes, FC_THREADS threads: mvp_in[lane][b] = predictor of block b (q, or ft_far()).
// Warps 1.. stage the tables of FC_BATCH blocks at a time (cp.async, FC_NBUF buffers); warp 0 walks the batches: one
// dependent shared-memory look-up per block.  Warp 0 issues no copies: behind a full load/store queue its own cp.async
// instructions would hold the walk back until the batch has been requested (profiles/microbench/chain_probe.cu:
// fetch 15.6 us + walk 18.8 us per CIF frame add up to 32 us when warp 0 copies too).
constexpr int FC_THREADS = 160, FC_BATCH = 16, FC_NBUF = 4;
__global__ void __launch_bounds__(FC_THREADS) fastme_chain_kernel(MeArgs a, const uint8_t* ref_base, size_t ref_plane_bytes,
                                                                int ref_pitch, const uint16_t* tables, uint32_t* mvp_in) {
    extern __shared__ __align__(16) uint16_t s_tab[];   // [FC_NBUF][FC_BATCH][FT_STRIDE]
    const int tid = threadIdx.x, lane = tid & 31, fl = blockIdx.x;
    const MeLane& L = a.lanes[fl];
    const uint8_t* curp = a.cur_base + (size_t)L.cur_plane * a.cur_plane_bytes;
    const uint16_t* tab_lane = tables + (size_t)fl * a.nblk * FT_STRIDE;
    const size_t kp_stride = (size_t)a.nblk * a.map_stride;
    const uint16_t* map_lane = a.sad_map + (size_t)fl * a.max_refs * a.nphase * kp_stride;
    const int nbatch = (a.nblk + FC_BATCH - 1) / FC_BATCH;
    auto prefetch = [&](int bt) {
        if (bt < nbatch) {
            const int nb = min(FC_BATCH, a.nblk - bt * FC_BATCH);
            const uint16_t* src = tab_lane + (size_t)bt * FC_BATCH * FT_STRIDE;
            uint16_t* dst = s_tab + (size_t)(bt % FC_NBUF) * FC_BATCH * FT_STRIDE;
            for (int v = tid - 32; v < nb * (FT_STRIDE / 8); v += FC_THREADS - 32) cp_async16(dst + v * 8, src + v * 8);
        }
        cp_async_commit();   // one group per batch, empty past the end, so the wait count below stays constant
    };
    WalkLane wl;
    wl.init(lane, 6 * L.nref);
    uint32_t* out = mvp_in + (size_t)fl * a.nblk;
    // mv_field = {(0,0): [0,0]}  (PFrame.py:34); the running predictor lives in warp 0: q inside the grid, else (mvpx, mvpy)
    int q = FT_CENTER, mvpx = 0, mvpy = 0;
    bool far = false;
    if (tid >= 32)
        for (int bt = 0; bt < FC_NBUF - 1; bt++) prefetch(bt);
    for (int bt = 0; bt < nbatch; bt++) {
        if (tid >= 32) cp_async_wait<FC_NBUF - 2>();   // this thread's part of batch bt has landed
        __syncthreads();                               // ... and everybody else's; batch bt-1 is consumed
        if (tid >= 32) prefetch(bt + FC_NBUF - 1);     // into the buffer batch bt-1 used
        if (tid < 32) {
            const uint16_t* tb = s_tab + (size_t)(bt % FC_NBUF) * FC_BATCH * FT_STRIDE;
            const int b0 = bt * FC_BATCH, nb = min(FC_BATCH, a.nblk - b0);
            for (int i = 0; i < nb; i++, tb += FT_STRIDE) {
                if (!far) {
                    if (lane == 0) out[b0 + i] = (uint32_t)q;
                    const uint16_t e = tb[q];
                    if (e != FT_NONE) { q = e; continue; }
                    mvpx = ft_x(q);
                    mvpy = ft_y(q);
                } else if (lane == 0) {
                    out[b0 + i] = ft_far(mvpx, mvpy);
                }
                // border predictor or outside the grid: walk this block on the spot
                int sad;
                long long cmp = 0;
                walk_block(a, L, wl, curp, ref_base, ref_plane_bytes, ref_pitch, map_lane + (size_t)(b0 + i) * a.map_stride,
                           kp_stride, b0 + i, lane, mvpx, mvpy, sad, cmp);
                far = abs(mvpx) > FT_S + 1 || abs(mvpy) > FT_S + 1;
                if (!far) q = ft_q(mvpx, mvpy);
            }
        }
    }
}

// one warp per (lane, block): the block's walk from its known predictor
__global__ void __launch_bounds__(256) fastme_finish_kernel(MeArgs a, int lanes, const uint8_t* ref_base, size_t ref_plane_bytes,
                                                          int ref_pitch, const uint32_t* mvp_in, unsigned long long* cmp_out) {
    const int lane = threadIdx.x & 31;
    const long long w = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (w >= (long long)lanes * a.nblk) return;
    const int fl = (int)(w / a.nblk), b = (int)(w - (long long)fl * a.nblk);
    const MeLane& L = a.lanes[fl];
    const uint8_t* curp = a.cur_base + (size_t)L.cur_plane * a.cur_plane_bytes;
    const size_t kp_stride = (size_t)a.nblk * a.map_stride;
    const uint16_t* map_lane = a.sad_map + (size_t)fl * a.max_refs * a.nphase * kp_stride;
    WalkLane wl;
    wl.init(lane, 6 * L.nref);
    int mvpx, mvpy, sad = 0;
    ft_unpack(mvp_in[w], mvpx, mvpy);
    long long cmp = 0;
    walk_block(a, L, wl, curp, ref_base, ref_plane_bytes, ref_pitch, map_lane + (size_t)b * a.map_stride, kp_stride, b, lane, mvpx, mvpy,
               sad, cmp);
    if (lane == 0) {
        a.out[w] = make_int4(mvpx, mvpy, 0, sad);
        if (cmp_out) atomicAdd(cmp_out + fl, (unsigned long long)cmp);
    }
}

// ---- window walk (no SAD map) -------------------------------------------------------------------------------------------
// The SAD map costs as much as a full search of +-16 although a walk touches a few dozen positions, and the first direct
// kernel above pays an L2 round trip per candidate and a CTA-wide barrier pair per level.  Here the serial chain stays,
// but everything a level needs is in shared memory when the walk reaches the block: NS-1 blocks ahead the TMA unit is asked
// for the block's reference windows (every reference and phase plane, 16 pixels around the block -- where the candidates
// of a predictor within +-16 lie; rows and columns outside the plane are zero-filled and never read by a valid
// candidate), so the fetch costs a handful of instructions instead of a CTA's worth of address arithmetic.  The 6 x nRef
// candidates of a level are dealt to warps (one warp per candidate SAD, redux.sync for the sum); after one barrier every
// warp reduces the keys with redux.sync.  Candidates outside the window (predictor drifted past 16) read global memory.
// One CTA per frame; lanes run in parallel.  Two kernels: fastme_window_kernel for block sizes 4 / 8 / 32 (simple, byte-wise
// SADs) and fastme_window16_kernel, the tuned one (history in profiles/r1_experiments.md, section 7).
constexpr int FW_PITCH = 64, FW16_PITCH = 48, FW_MARGIN = 16, FW_WARPS = 12;
__device__ __forceinline__ void cp_async4(void* dst, const void* src) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(smem_u32(dst)), "l"(src) : "memory");
}

// candidate outside the staged window (predictor drifted past 16): 16x16 SAD straight from the plane.  Rare: kept out of line.
__device__ __noinline__ int fw_sad16_global(const uint8_t* ref_base, size_t ref_plane_bytes, int ref_pitch, int plane, int y, int x,
                                            uint32_t c0, uint32_t c1) {
    const uint8_t* pp = ref_base + (size_t)plane * ref_plane_bytes + (size_t)y * ref_pitch + x;
    const uint32_t sh = (uint32_t)(reinterpret_cast<uintptr_t>(pp) & 3u) * 8u;
    const uint32_t* q = reinterpret_cast<const uint32_t*>(pp - (sh >> 3));
    const uint32_t w0 = __ldg(q), w1 = __ldg(q + 1), w2 = sh ? __ldg(q + 2) : 0u;   // q[2] only when it holds pixels of the block
    return (int)sad4(__funnelshift_r(w1, w2, sh), c1, sad4(__funnelshift_r(w0, w1, sh), c0, 0u));
}

// Generic block sizes (4, 8, 32); 16x16 has its own kernel below.
template <int NS>   // NS: blocks in flight (ring of window slots): the fetch latency is several block walks long
__global__ void __launch_bounds__(32 * FW_WARPS) fastme_window_kernel(const __grid_constant__ CUtensorMap win_map, MeArgs a,
                                                                         const uint8_t* ref_base, size_t ref_plane_bytes, int ref_pitch,
                                                                         long long* cmp_out) {
    extern __shared__ __align__(128) uint8_t s_win[];   // [NS][max_refs * nphase][rows][FW_PITCH] | 16 | cur [NS][bs * bs]
    __shared__ uint64_t bars[NS];
    __shared__ int s_sad[2][FM_MAXC];
    __shared__ int s_plane[BVC_MAX_REFS];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, fl = blockIdx.x;
    const int bs = a.bs, rows = bs + 2 * FW_MARGIN, nphase = a.nphase;
    const int nref = a.lanes[fl].nref, ncand = 6 * nref, tri = nref * (nref + 1) / 2, nkp = nref * nphase;
    const int win_bytes = rows * FW_PITCH, slot_bytes = a.max_refs * nphase * win_bytes;
    uint8_t* s_cur = s_win + NS * slot_bytes + 16;
    const uint8_t* curp = a.cur_base + (size_t)a.lanes[fl].cur_plane * a.cur_plane_bytes;
    if (tid < BVC_MAX_REFS) s_plane[tid] = a.lanes[fl].ref_plane[tid];
    if (tid == 0) {
        for (int i = 0; i < NS; i++) mbar_init(&bars[i], 1);
        fence_mbar_init();
    }
    __syncthreads();
    int pox = 0, poy = 0, pbx = 0, pb = 0;   // the block the next prefetch is for (warp 0)
    auto prefetch = [&]() {
        if (warp == 0) {
            if (pb < a.nblk) {
                const int slot = pb % NS;
                if (lane == 0) {
                    uint8_t* dst = s_win + slot * slot_bytes;
                    mbar_arrive_expect_tx(&bars[slot], (uint32_t)(nkp * win_bytes));
                    const int x0 = (pox - FW_MARGIN) & ~15, y0 = poy - FW_MARGIN;
                    for (int kp = 0; kp < nkp; kp++) {
                        const int k = nphase == 4 ? kp >> 2 : kp, ph = nphase == 4 ? kp & 3 : 0;
                        tma_load_3d(dst + kp * win_bytes, &win_map, &bars[slot], x0, y0, s_plane[k] + ph);
                    }
                }
                uint8_t* dc = s_cur + slot * bs * bs;
                const uint8_t* sc = curp + (size_t)poy * a.cur_pitch + pox;
                for (int i = lane; i < bs * bs / 4; i += 32) {
                    const int y = (4 * i) / bs, xx = 4 * i - y * bs;
                    cp_async4(dc + 4 * i, sc + (size_t)y * a.cur_pitch + xx);
                }
                pb++;
                pox += bs;
                if (++pbx == a.bw) { pbx = 0; pox = 0; poy += bs; }
            }
            cp_async_commit();   // one group per block, empty past the end, so the wait count below stays constant
        }
    };
    // this warp's candidates c = warp + nwarps * j: reference k, key p -> offsets from the predictor
    constexpr int MAXJ = (FM_MAXC + FW_WARPS - 1) / FW_WARPS;   // candidates per warp (one reference: 6 candidates, 6 warps)
    const int nwarps = blockDim.x >> 5;
    int cwin[MAXJ], cdx[MAXJ], cdy[MAXJ], cpl[MAXJ];
    bool corg[MAXJ];
#pragma unroll
    for (int j = 0; j < MAXJ; j++) {
        const int c = warp + nwarps * j, k = c / 6, p = c - 6 * k;
        cwin[j] = k * nphase * win_bytes;
        corg[j] = p == 0;
        cdx[j] = (p == 3) - (p == 5);
        cdy[j] = (p == 4) - (p == 2);
        cpl[j] = c < ncand ? s_plane[k] : 0;
    }
    const bool frac = a.sc == 2;
    const int wmax = a.W - bs, hmax = a.H - bs;
    long long cmp_total = 0;
    int mvpx = 0, mvpy = 0;   // mv_field = {(0,0): [0,0]}  (PFrame.py:34); every thread carries the same values
    int par = 0;
    for (int b = 0; b < NS - 1; b++) prefetch();
    int ox = 0, oy = 0, bx = 0;
    for (int b = 0; b < a.nblk; b++) {
        if (warp == 0) cp_async_wait<NS - 2>();   // block b's current pixels (copied by this warp)
        __syncthreads();                          // ... visible to everyone; everybody is done with block b-1's slot
        prefetch();                               // block b + NS - 1, into that slot
        mbar_wait(&bars[b % NS], (uint32_t)(b / NS) & 1u);   // block b's windows have landed
        const int wx = ox - ((ox - FW_MARGIN) & ~15);   // window column of the block's own position
        const uint8_t* win = s_win + (b % NS) * slot_bytes;
        const uint8_t* cur = s_cur + (b % NS) * bs * bs;
        int best_sad = 0;
        for (;;) {
#pragma unroll
            for (int j = 0; j < MAXJ; j++) {
                const int c = warp + nwarps * j;
                if (c >= ncand) break;
                const int cx = corg[j] ? 0 : mvpx + cdx[j];
                const int cy = corg[j] ? 0 : mvpy + cdy[j];
                int ph = 0, phx = 0, phy = 0, dx = cx, dy = cy;
                if (frac) { phx = cx & 1; phy = cy & 1; ph = phx | (phy << 1); dx = cx >> 1; dy = cy >> 1; }
                // is_out_of_range (block_predictor.py:116-143), expressed on the phase plane
                const bool ok = (ox + dx >= 0) && (oy + dy >= 0) && (ox + dx <= wmax - phx) && (oy + dy <= hmax - phy);
                const bool inwin = (unsigned)(dx + FW_MARGIN) <= 2u * FW_MARGIN && (unsigned)(dy + FW_MARGIN) <= 2u * FW_MARGIN;
                int s = -1;
                if (ok) {
                    const uint8_t* rp = inwin ? win + cwin[j] + ph * win_bytes + (dy + FW_MARGIN) * FW_PITCH + wx + dx
                                              : ref_base + (size_t)(cpl[j] + ph) * ref_plane_bytes + (size_t)(oy + dy) * ref_pitch + (ox + dx);
                    const int rpitch = inwin ? FW_PITCH : ref_pitch;
                    int t = 0;
                    for (int i = lane; i < bs * bs; i += 32) {
                        const int y = i / bs, x = i - y * bs;
                        t += abs((int)cur[i] - (int)rp[(size_t)y * rpitch + x]);
                    }
                    s = (int)__reduce_add_sync(0xffffffffu, (unsigned)t);
                }
                if (lane == 0) s_sad[par][c] = s;
            }
            __syncthreads();
            // first strict minimum in (reference ascending, key order): lane c holds candidate c (and c + 32)
            uint32_t key = 0xffffffffu;
            bool v0 = false;
            if (lane < ncand) {
                const int s = s_sad[par][lane];
                if (s >= 0) { key = ((uint32_t)s << 8) | (uint32_t)lane; v0 = lane < 6; }
            }
            if (lane + 32 < ncand) {
                const int s = s_sad[par][lane + 32];
                if (s >= 0) key = min(key, ((uint32_t)s << 8) | (uint32_t)(lane + 32));
            }
            key = __reduce_min_sync(0xffffffffu, key);
            const int nvalid = __popc(__ballot_sync(0xffffffffu, v0));
            par ^= 1;
            cmp_total += (long long)(nvalid * tri);
            const int best_p = (key == 0xffffffffu) ? 0 : (int)(key & 255u) % 6;
            best_sad = (key == 0xffffffffu) ? 0x7fffffff : (int)(key >> 8);
            const int mvx = best_p == 0 ? 0 : mvpx + (best_p == 3) - (best_p == 5);
            const int mvy = best_p == 0 ? 0 : mvpy + (best_p == 4) - (best_p == 2);
            const bool stop = (best_p <= 1) || abs(mvx) >= 16 || abs(mvy) >= 16;
            mvpx = mvx;
            mvpy = mvy;
            if (stop) break;
        }
        if (tid == 0) a.out[(size_t)fl * a.nblk + b] = make_int4(mvpx, mvpy, 0, best_sad);
        ox += bs;
        if (++bx == a.bw) { bx = 0; ox = 0; oy += bs; }
    }
    if (tid == 0 && cmp_out) cmp_out[fl] = cmp_total;
}

// The 16x16 specialisation of the window walk, written for a short per-warp instruction stream: the level is a dependent
// chain (predictor -> addresses -> LDS -> 2 x VABSDIFF4 -> redux -> barrier -> LDS -> redux -> decode), and with few warps
// per scheduler every instruction of a warp costs several cycles whether it is on that chain or not.  So: everything that
// does not depend on the predictor is precomputed per warp or per block, the key-to-vector decode goes through two nibble
// tables, arithmetic is 32-bit throughout, the global-memory path for candidates outside the window is out of line, and
// the fetch is warp-specialised: the last warp only feeds the TMA unit (it joins the per-block barrier, which tells it that
// a slot is free), the consumer warps request their own 8 current pixels of the next block one block ahead and meet at a
// named barrier inside a level.
template <bool FRAC, int NJ, int NS>
__global__ void __launch_bounds__(32 * 25) fastme_window16_kernel(const __grid_constant__ CUtensorMap win_map, MeArgs a,
                                                                const uint8_t* ref_base, size_t ref_plane_bytes, int ref_pitch,
                                                                long long* cmp_out) {
    // window rows of 48 bytes (12 banks: rows r and r + 8 share banks, a 2-way conflict; the generic kernel's 64-byte rows
    // alias every second row, 8-way)
    constexpr int BS = 16, ROWS = BS + 2 * FW_MARGIN, PITCH = FW16_PITCH, WIN = ROWS * PITCH, NPH = FRAC ? 4 : 1;
    extern __shared__ __align__(128) uint8_t s_win[];   // [NS][max_refs * NPH][ROWS][PITCH] | 16
    __shared__ uint64_t bars[NS];
    __shared__ int s_sad[2][FM_MAXC];
    __shared__ int s_plane[BVC_MAX_REFS];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, fl = blockIdx.x;
    const int ncw = (blockDim.x >> 5) - 1;   // consumer warps; the last warp only feeds the TMA unit
    const int nref = a.lanes[fl].nref, ncand = 6 * nref, tri = nref * (nref + 1) / 2;
    const int slot_bytes = a.max_refs * NPH * WIN;
    if (tid < BVC_MAX_REFS) s_plane[tid] = a.lanes[fl].ref_plane[tid];
    if (tid == 0) {
        for (int i = 0; i < NS; i++) mbar_init(&bars[i], 1);
        fence_mbar_init();
    }
    __syncthreads();
    if (warp == ncw) {
        // ---- producer: block b's windows are requested NS-1 blocks ahead, into the slot block b-NS has left ----
        int pox = 0, poy = 0, pbx = 0, pb = 0;
        auto request = [&]() {
            if (lane == 0 && pb < a.nblk) {
                const int slot = pb & (NS - 1);
                uint8_t* dst = s_win + slot * slot_bytes;
                mbar_arrive_expect_tx(&bars[slot], (uint32_t)(nref * NPH * WIN));
                for (int r = 0; r < nref; r++)   // the box is NPH planes deep: a reference's phase planes are consecutive
                    tma_load_3d(dst + r * NPH * WIN, &win_map, &bars[slot], pox - FW_MARGIN, poy - FW_MARGIN, s_plane[r]);
                pb++;
                pox += BS;
                if (++pbx == a.bw) { pbx = 0; pox = 0; poy += BS; }
            }
        };
        for (int b = 0; b < NS - 1; b++) request();
        for (int b = 0; b < a.nblk; b++) {
            __syncthreads();   // the consumers are done with block b-1: its slot is free
            request();
        }
        return;
    }
    // ---- consumers ----
    // this warp's candidates c = warp + ncw * j: reference k, key p; everything that does not depend on the predictor
    int c_[NJ], keep[NJ], sx[NJ], sy[NJ], lanebase[NJ], plane[NJ];
#pragma unroll
    for (int j = 0; j < NJ; j++) {
        const int c = warp + ncw * j, k = c / 6, p = c - 6 * k;
        c_[j] = c;
        keep[j] = p == 0 ? 0 : -1;                 // origin ignores the predictor
        sx[j] = (p == 3) - (p == 5);
        sy[j] = (p == 4) - (p == 2);
        // byte offset, inside a slot, of this lane's 8 pixels (row lane/2, half lane%2) for a zero displacement
        lanebase[j] = k * NPH * WIN + (FW_MARGIN + (lane >> 1)) * PITCH + FW_MARGIN + 8 * (lane & 1);
        plane[j] = c < ncand ? s_plane[k] : 0;
    }
    const int wmax = a.W - BS, hmax = a.H - BS, ncthreads = 32 * ncw;
    const uint32_t win0 = smem_u32(s_win);
    // this lane's 8 current pixels of the next block, requested one block ahead
    const uint8_t* curl = a.cur_base + (size_t)a.lanes[fl].cur_plane * a.cur_plane_bytes + (size_t)(lane >> 1) * a.cur_pitch + 8 * (lane & 1);
    const size_t cur_row_step = (size_t)BS * a.cur_pitch;
    uint2 cw_next = *reinterpret_cast<const uint2*>(curl);
    long long cmp_total = 0;
    int mvpx = 0, mvpy = 0;   // mv_field = {(0,0): [0,0]}  (PFrame.py:34); every thread carries the same values
    int par = 0;
    int ox = 0, oy = 0, bx = 0;
    int4* out = a.out + (size_t)fl * a.nblk;
    for (int b = 0; b < a.nblk; b++) {
        const int slot = b & (NS - 1);
        __syncthreads();                          // everybody is done with block b-1 (the producer refills its slot)
        const uint32_t cw0 = cw_next.x, cw1 = cw_next.y;
        if (b + 1 < a.nblk) cw_next = *reinterpret_cast<const uint2*>(bx + 1 == a.bw ? curl + cur_row_step : curl + ox + BS);
        mbar_wait(&bars[slot], (uint32_t)(b / NS) & 1u);   // block b's windows have landed
        const uint32_t wslot = win0 + slot * slot_bytes;
        int best_sad = 0, cmp_blk = 0;
        for (;;) {
            // the candidates of a warp in phases (all loads, then all sums, then all reductions) so that their latencies overlap
            uint32_t w0[NJ], w1[NJ], w2[NJ], sh[NJ];
            int kind[NJ];   // 0 = leaves the plane or no candidate, 1 = in the window, 2 = outside the window
#pragma unroll
            for (int j = 0; j < NJ; j++) {
                const int cx = (mvpx & keep[j]) + sx[j], cy = (mvpy & keep[j]) + sy[j];
                int phx = 0, phy = 0, dx = cx, dy = cy;
                if (FRAC) { phx = cx & 1; phy = cy & 1; dx = cx >> 1; dy = cy >> 1; }
                const int px = ox + dx, py = oy + dy;
                // is_out_of_range (block_predictor.py:116-143), expressed on the phase plane
                const bool ok = c_[j] < ncand && px >= 0 && py >= 0 && px <= wmax - phx && py <= hmax - phy;
                const bool inwin = (unsigned)(dx + FW_MARGIN) <= 2u * FW_MARGIN && (unsigned)(dy + FW_MARGIN) <= 2u * FW_MARGIN;
                kind[j] = ok ? (inwin ? 1 : 2) : 0;
                w0[j] = w1[j] = w2[j] = sh[j] = 0u;
                if (kind[j] == 1) {
                    const uint32_t off = wslot + lanebase[j] + (FRAC ? (phx + 2 * phy) * WIN : 0) + dy * PITCH + dx;
                    sh[j] = (off & 3u) * 8u;
                    // w2 may be the next row / the pad: shifted out when sh == 0
                    asm volatile("ld.shared.u32 %0, [%3];\n\tld.shared.u32 %1, [%3+4];\n\tld.shared.u32 %2, [%3+8];"
                                 : "=r"(w0[j]), "=r"(w1[j]), "=r"(w2[j]) : "r"(off & ~3u));
                } else if (kind[j] == 2) {
                    w0[j] = (uint32_t)fw_sad16_global(ref_base, ref_plane_bytes, ref_pitch, plane[j] + (FRAC ? phx + 2 * phy : 0), py + (lane >> 1),
                                                      px + 8 * (lane & 1), cw0, cw1);
                }
            }
            uint32_t t[NJ];
#pragma unroll
            for (int j = 0; j < NJ; j++)
                t[j] = kind[j] == 1 ? sad4(__funnelshift_r(w1[j], w2[j], sh[j]), cw1, sad4(__funnelshift_r(w0[j], w1[j], sh[j]), cw0, 0u)) : w0[j];
            // a lane's partial sum is below 2^11 and a block's SAD below 2^16: two candidates share one redux.sync
#pragma unroll
            for (int j = 0; j + 1 < NJ; j += 2) {
                const uint32_t both = __reduce_add_sync(0xffffffffu, t[j] | (t[j + 1] << 16));
                t[j] = both & 0xffffu;
                t[j + 1] = both >> 16;
            }
            if (NJ & 1) t[NJ - 1] = __reduce_add_sync(0xffffffffu, t[NJ - 1]);
#pragma unroll
            for (int j = 0; j < NJ; j++)
                if (lane == 0 && c_[j] < ncand) s_sad[par][c_[j]] = kind[j] ? (int)t[j] : -1;
            asm volatile("bar.sync 1, %0;" ::"r"(ncthreads) : "memory");   // consumers only
            // first strict minimum in (reference ascending, key order): lane c holds candidate c (and c + 32)
            uint32_t key = 0xffffffffu;
            bool v0 = false;
            if (lane < ncand) {
                const int s = s_sad[par][lane];
                if (s >= 0) { key = ((uint32_t)s << 8) | (uint32_t)lane; v0 = lane < 6; }
            }
            if (NJ > 1 && lane + 32 < ncand) {   // more than 32 candidates: more than 5 references
                const int s = s_sad[par][lane + 32];
                if (s >= 0) key = min(key, ((uint32_t)s << 8) | (uint32_t)(lane + 32));
            }
            key = __reduce_min_sync(0xffffffffu, key);
            if (warp == 0) cmp_blk += __popc(__ballot_sync(0xffffffffu, v0)) * tri;   // a statistic: only thread 0 reports it
            par ^= 1;
            // key -> vector: p = candidate % 6; offsets from two nibble tables (origin, pmv, top, right, bottom, left)
            const uint32_t c = key & 255u, p = key == 0xffffffffu ? 0u : c - 6u * ((c * 43u) >> 8);
            best_sad = key == 0xffffffffu ? 0x7fffffff : (int)(key >> 8);
            const int ddx = (int)((0x012111u >> (4u * p)) & 15u) - 1, ddy = (int)((0x121011u >> (4u * p)) & 15u) - 1;
            mvpx = (p ? mvpx : 0) + ddx;
            mvpy = (p ? mvpy : 0) + ddy;
            if (p <= 1u || abs(mvpx) >= 16 || abs(mvpy) >= 16) break;
        }
        cmp_total += cmp_blk;
        if (tid == 0) out[b] = make_int4(mvpx, mvpy, 0, best_sad);
        ox += BS;
        if (++bx == a.bw) { bx = 0; ox = 0; oy += BS; curl += cur_row_step; }
    }
    if (tid == 0 && cmp_out) cmp_out[fl] = cmp_total;
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

namespace bvc {

size_t fastme_table_bytes(int lanes, int nblk) {
    // a multiple of 16 per lane, so a lane group's slice (offset = the bytes of the lanes before it) stays cp.async-aligned
    return (size_t)lanes * ((size_t)nblk * FT_STRIDE * sizeof(uint16_t) + ((size_t)nblk * sizeof(uint32_t) + 15) / 16 * 16);
}

cudaError_t launch_fastme_table(const MeArgs& a, int lanes, const uint8_t* ref_base, size_t ref_plane_bytes, int ref_pitch,
                                void* scratch, long long* cmp_out, cudaStream_t st) {
    if (!a.sad_map || a.bs > 32 || a.map_stride % 8 || !scratch) return cudaErrorInvalidValue;
    uint16_t* tables = static_cast<uint16_t*>(scratch);
    uint32_t* mvp_in = reinterpret_cast<uint32_t*>(tables + (size_t)lanes * a.nblk * FT_STRIDE);   // FT_STRIDE even: 4-byte aligned
    const size_t smem = (size_t)FC_NBUF * FC_BATCH * FT_STRIDE * sizeof(uint16_t);
    static bool configured_dev[BVC_MAX_DEVICES] = {};
    bool& configured = configured_dev[current_device_slot()];
    if (!configured) {
        cudaError_t e = cudaFuncSetAttribute(fastme_chain_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        configured = true;
    }
    fastme_table_kernel<<<dim3(a.nblk, lanes), FT_THREADS, 0, st>>>(a, tables);
    fastme_chain_kernel<<<lanes, FC_THREADS, smem, st>>>(a, ref_base, ref_plane_bytes, ref_pitch, tables, mvp_in);
    cudaError_t e = cmp_out ? cudaMemsetAsync(cmp_out, 0, (size_t)lanes * sizeof(long long), st) : cudaSuccess;
    if (e != cudaSuccess) return e;
    const long long warps = (long long)lanes * a.nblk;
    fastme_finish_kernel<<<(unsigned)((warps + 7) / 8), 256, 0, st>>>(a, lanes, ref_base, ref_plane_bytes, ref_pitch, mvp_in,
                                                                       reinterpret_cast<unsigned long long*>(cmp_out));
    return cudaGetLastError();
}

static size_t fw_smem(const MeArgs& a, int max_refs, int ns) {
    if (a.bs == 16) return (size_t)ns * max_refs * a.nphase * (16 + 2 * FW_MARGIN) * FW16_PITCH + 16;   // fastme_window16_kernel
    const size_t slot = (size_t)max_refs * a.nphase * (a.bs + 2 * FW_MARGIN) * FW_PITCH;
    return ns * (slot + (size_t)a.bs * a.bs) + 16;
}
size_t fastme_window_smem(const MeArgs& a, int max_refs) { return fw_smem(a, max_refs, 2); }
void fastme_window_box(int bs, int nphase, int* box_w, int* box_h, int* box_d) {
    *box_w = bs == 16 ? FW16_PITCH : FW_PITCH;
    *box_h = bs + 2 * FW_MARGIN;
    *box_d = bs == 16 ? nphase : 1;   // the 16x16 kernel fetches the phase planes of a reference with one request
}

template <int NS>
static cudaError_t launch_fw(const CUtensorMap& map, const MeArgs& a, int lanes, size_t smem, const uint8_t* ref_base, size_t ref_plane_bytes,
                             int ref_pitch, long long* cmp_out, cudaStream_t st) {
    static size_t configured_dev[BVC_MAX_DEVICES] = {};
    size_t& configured = configured_dev[current_device_slot()];
    if (smem > configured) {
        cudaError_t e = cudaFuncSetAttribute(fastme_window_kernel<NS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        configured = smem;
    }
    const int warps = a.max_refs >= 2 ? FW_WARPS : 6;   // 6 x nRef candidates per level
    fastme_window_kernel<NS><<<lanes, 32 * warps, smem, st>>>(map, a, ref_base, ref_plane_bytes, ref_pitch, cmp_out);
    return cudaGetLastError();
}

template <bool FRAC, int NJ, int NS>
static cudaError_t launch_fw16(const CUtensorMap& map, const MeArgs& a, int lanes, size_t smem, const uint8_t* ref_base, size_t ref_plane_bytes,
                               int ref_pitch, long long* cmp_out, cudaStream_t st) {
    static size_t configured_dev[BVC_MAX_DEVICES] = {};
    size_t& configured = configured_dev[current_device_slot()];
    if (smem > configured) {
        cudaError_t e = cudaFuncSetAttribute(fastme_window16_kernel<FRAC, NJ, NS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        configured = smem;
    }
    const int warps = (6 * a.max_refs + NJ - 1) / NJ + 1;   // consumer warps + the producer warp
    fastme_window16_kernel<FRAC, NJ, NS><<<lanes, 32 * warps, smem, st>>>(map, a, ref_base, ref_plane_bytes, ref_pitch, cmp_out);
    return cudaGetLastError();
}

cudaError_t launch_fastme_window(const CUtensorMap* win_map, const MeArgs& a, int lanes, int max_refs, const uint8_t* ref_base,
                                 size_t ref_plane_bytes, int ref_pitch, long long* cmp_out, cudaStream_t st) {
    const size_t limit = 200 * 1024;
    if (!win_map || a.bs > 32 || a.bs % 4 || max_refs < 1 || max_refs > BVC_MAX_REFS || fw_smem(a, max_refs, 2) > limit)
        return cudaErrorInvalidValue;
    MeArgs aa = a;
    aa.max_refs = max_refs;
    const int ns = fw_smem(a, max_refs, 8) <= limit ? 8 : fw_smem(a, max_refs, 4) <= limit ? 4 : 2;
    const size_t smem = fw_smem(a, max_refs, ns);
    if (a.bs == 16) {
        if ((a.nphase == 4) != (a.sc == 2)) return cudaErrorInvalidValue;
#define BVC_FW16(F, J, N) launch_fw16<F, J, N>(*win_map, aa, lanes, smem, ref_base, ref_plane_bytes, ref_pitch, cmp_out, st)
#define BVC_FW16N(F, J) (ns == 8 ? BVC_FW16(F, J, 8) : ns == 4 ? BVC_FW16(F, J, 4) : BVC_FW16(F, J, 2))
        // one candidate per warp up to 4 references (24 consumer warps: 0.96 us per block against 1.16 with 12 warps x 2)
        if (a.sc == 2) return max_refs <= 4 ? BVC_FW16N(true, 1) : BVC_FW16N(true, 2);
        return max_refs <= 4 ? BVC_FW16N(false, 1) : BVC_FW16N(false, 2);
#undef BVC_FW16N
#undef BVC_FW16
    }
#define BVC_FW(N) launch_fw<N>(*win_map, aa, lanes, smem, ref_base, ref_plane_bytes, ref_pitch, cmp_out, st)
    return ns == 8 ? BVC_FW(8) : ns == 4 ? BVC_FW(4) : BVC_FW(2);
#undef BVC_FW
}

}  // namespace bvc
