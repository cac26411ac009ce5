// tq.cu -- K5 (P-frame residual/transform/reconstruct), K6 (I-frame intra wavefront) and the per-block
// half of K7 (entropy coding of the quantised levels), fused so a block's residual, coefficients and
// levels never round-trip HBM before they are coded.
//
// Replaces PFrame.process_block / generate_residual_block / find_mv_predicted_block
// (reference encoder/PFrame.py:99-125,230-249), IFrame.process_block + intra predictors
// (encoder/IFrame.py:184-231), apply_dct_and_quantization / reconstruct_block (encoder/Frame.py:190-202)
// and the per-block part of entropy_encode_dct_coffs_row (encoder/Frame.py:61-75).
#include "bvc_kernels.h"
#include "tq_device.cuh"
#include <algorithm>

namespace bvc {
namespace {

constexpr int TQ_WARPS = 4;

// q = (mulhi(x, magic) + x) >> shift == x / d for every x < 2^31 (fast_div in tq_device.cuh)
static void fast_div_constants(uint32_t d, uint32_t& magic, uint32_t& shift) {
    shift = 0;
    while ((1u << shift) < d) shift++;
    magic = (uint32_t)((((unsigned long long)1 << 32) * (((unsigned long long)1 << shift) - d)) / d + 1);
}

template <int BS>
struct TqCtaSmem {
    WarpTile<BS> w[TQ_WARPS];
    uint8_t zz[BS * BS];
};

// ---------------------------------------------------------------------------------------------
// P frames: every block independent.  A warp task = NBW consecutive blocks of one lane; the grid is persistent (at most
// the resident CTAs, optionally fewer: a.cta_cap) and its warps draw tasks from a ticket counter.  The cap was an experiment
// on the clip path: this kernel runs on a high-priority stream while the other lane group's search is on the GPU, and the
// block scheduler hands a high-priority kernel every slot the search frees -- capped at one search CTA's registers per SM the
// transform would run *beside* a search CTA; measured slower at every cap (profiles/r2_experiments.md section 4), so the
// default is no cap.
// Warp task wt of a launch: lane group fl = wt / tasks_per_lane, blocks (wt % tasks_per_lane) * NBW + q of the launch's rows.
template <int BS>
__device__ __forceinline__ PTask pframe_task(const TqArgs& a, int wt, int tasks_per_lane, int q) {
    constexpr int NBW = 32 / BS;
    PTask k;
    k.fl = (int)fast_div((uint32_t)wt, a.ux_magic, a.ux_shift);
    const int rem = wt - k.fl * tasks_per_lane;
    const int blk_end = (a.row_begin + a.row_count) * a.bw;
    const int b = a.row_begin * a.bw + rem * NBW + q;
    k.valid = b < blk_end;
    k.b = k.valid ? b : blk_end - 1;
    const int by = (int)fast_div((uint32_t)k.b, a.bw_magic, a.bw_shift);
    k.oy = by * BS;
    k.ox = (k.b - by * a.bw) * BS;
    return k;
}

// Persistent grid (at most the resident CTAs of the GPU): every warp walks warp tasks on its own -- no CTA barrier after
// the zig-zag table -- the first one from its position in the grid, the following ones from a ticket counter (a.ticket, left
// at zero again by the warp that draws the last ticket), so the launch balances itself whatever the other kernels on the
// GPU do.  The ticket is drawn before the transform and read after it; the next task's motion vector is requested before
// the event pass of the entropy coder and its pixel rows before the coding pass, so no global round trip is exposed.
// Register budget: 16x16 blocks are compiled for four CTAs per SM -- ptxas then takes 92 registers (five CTAs still fit) and
// keeps the transform's cosines in uniform registers instead of re-materialising them for every pass (61 instead of 163 UMOV,
// 144 instructions fewer per block pair); the smaller block sizes need fewer registers anyway and keep six CTAs.
template <int BS, bool DBG>
__global__ void __launch_bounds__(TQ_WARPS * 32, BS == 16 ? 4 : 6) tq_pframe_kernel(TqArgs a, int tasks_per_lane, int ntasks) {
    static_assert(sizeof(EntScratch<BS>) <= sizeof(WarpTile<BS>::buf), "entropy scratch must fit the fp64 exchange buffer");
    extern __shared__ __align__(16) uint8_t smraw[];
    TqCtaSmem<BS>& sm = *reinterpret_cast<TqCtaSmem<BS>*>(smraw);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    build_zigzag<BS>(sm.zz, threadIdx.x, blockDim.x);
    __syncthreads();
    WarpTile<BS>& t = sm.w[warp];
    EntScratch<BS>& es = *reinterpret_cast<EntScratch<BS>*>(&t.buf[0][0][0]);
    const int q = lane / BS, x = lane % BS;
    const int nwarps = (int)gridDim.x * TQ_WARPS;
    int wt = (int)blockIdx.x * TQ_WARPS + warp;
    if (wt >= ntasks) return;
    PTask k = pframe_task<BS>(a, wt, tasks_per_lane, q);
    uint32_t sh = pframe_request_rows<BS>(a, t, k, q, x, a.mv[(size_t)k.fl * a.nblk + k.b]);
#pragma unroll 1
    for (;;) {
        int tk = 0;
        if (lane == 0) tk = ticket_draw(a.ticket);
        pframe_stage_rows<BS>(t, q, x, sh);
        if (DBG && a.resid_nomc && k.valid) {
            // PFrame.py:40,64,103,116: int16(cur) - int16(refs[0]) stored into an int8 plane
            const FrameLane& L = a.lanes[k.fl];
            const uint8_t* r0 = a.ref_base + (size_t)L.ref_plane[0] * a.ref_plane_bytes + (size_t)(k.oy + x) * a.ref_pitch + k.ox;
            int8_t* d = a.resid_nomc + ((size_t)k.fl * a.H + k.oy + x) * a.W + k.ox;
#pragma unroll
            for (int i = 0; i < BS; i++) d[i] = (int8_t)((int)t.cur[q][x][i] - (int)r0[i]);
        }
        __syncwarp();
        {
            TqOut o;
            o.levels = a.levels ? a.levels + ((size_t)k.fl * a.H + k.oy) * a.W + k.ox : nullptr;
            o.lev_pitch = a.W;
            o.recon = a.ref_base + (size_t)a.lanes[k.fl].out_plane * a.ref_plane_bytes + (size_t)k.oy * a.ref_pitch + k.ox;
            o.rec_pitch = a.ref_pitch;
            o.resid_mc = a.resid_mc ? a.resid_mc + ((size_t)k.fl * a.H + k.oy) * a.W + k.ox : nullptr;
            o.resid_pitch = a.W;
            o.idct_out = nullptr;
            o.coef_out = nullptr;
            const int qp = a.qp_rows[(size_t)k.fl * a.bh + k.oy / BS];
            tq_warp<BS, DBG>(t, lane, k.valid, qp, o, nullptr, nullptr, false);
        }
        // tickets 0 .. ntasks-1 are drawn in all (ntasks - nwarps good ones, one bad one per warp): the last puts the counter back
        if (lane == 0 && tk == ntasks - 1) atomicExch(a.ticket, 0);
        const int wt2 = nwarps + __shfl_sync(0xffffffffu, tk, 0);
        const bool more = wt2 < ntasks;
        PTask k2 = k;
        int4 mv2 = make_int4(0, 0, 0, 0);
        if (more) {
            k2 = pframe_task<BS>(a, wt2, tasks_per_lane, q);
            mv2 = ldg_mv_keep(a.mv + (size_t)k2.fl * a.nblk + k2.b);
        }
        const uint32_t vmask = __ballot_sync(0xffffffffu, k.valid);
        const int E = entropy_tile_events<BS>(t, es, sm.zz, lane, vmask);
        if (more) sh = pframe_request_rows<BS>(a, t, k2, q, x, mv2);
        entropy_tile_code<BS>(es, E, lane);
        const size_t bi = (size_t)k.fl * a.nblk + k.b;
        entropy_tile_store<BS>(es, lane, k.valid, a.blk_bits + bi * a.blk_words, a.blk_nbits + bi);
        if (!more) break;
        k = k2;
        wt = wt2;
    }
}

// Mode decision of the intra wavefront, IFrame.py:184-195, from registers: lane (q, r) holds row r of the current block
// (cw), the block's left column as a row (lw, the same for every lane of the block) and its own top pixel tv.
//   mode 0 ("horizontal"): pred[r][c] = left[c]  -> this lane's share of the SAD is sum_c (cur[r][c] - left[c]) mod 256
//   mode 1 ("vertical"):   pred[r][c] = top[r]   ->                              sum_c (cur[r][c] - top[r]) mod 256
// In-frame predictors are uint8, so cur - pred wraps mod 256 (:189-190); border predictors are int64 128: a true absolute
// difference.  Per-byte SIMD: __vsub4 wraps, __vsadu4(v, 0) sums four bytes.  The sums over the block's lanes (integer,
// any order) follow with shuffles; same totals as the reference's per-pixel loop.
template <int BS>
__device__ __forceinline__ void intra_mode_sads(const uint32_t (&cw)[BS / 4], const uint32_t (&lw)[BS / 4], int tv, bool has_left, bool has_top,
                                                int& sh, int& sv) {
    const uint32_t tw = (uint32_t)tv * 0x01010101u;
    uint32_t h = 0, v = 0;
#pragma unroll
    for (int j = 0; j < BS / 4; j++) {
        h += has_left ? __vsadu4(__vsub4(cw[j], lw[j]), 0u) : __vsadu4(cw[j], 0x80808080u);
        v += has_top ? __vsadu4(__vsub4(cw[j], tw), 0u) : __vsadu4(cw[j], 0x80808080u);
    }
    sh = (int)h;
    sv = (int)v;
#pragma unroll
    for (int d = 1; d < BS; d <<= 1) {
        sh += __shfl_xor_sync(0xffffffffu, sh, d);
        sv += __shfl_xor_sync(0xffffffffu, sv, d);
    }
}

// ---------------------------------------------------------------------------------------------
// I frames: block (bx,by) needs the reconstructed right column of (bx-1,by) and bottom row of
// (bx,by-1) (IFrame.py:184-213) => anti-diagonal wavefront.  One warp (= one CTA) walks one block row
// of NBW *different frames* (lanes) left to right, row r trailing row r-1 by one block.  The hand-over between rows is
// the data itself: a block posts its bottom row into a mailbox of 32-bit words, pixel | epoch << 8, and the block below
// polls those words until they carry this frame's epoch.  One L2 round trip per block on the consumer side and plain
// stores on the producer side -- no progress counter, no __threadfence in either warp (the counter + two fences + a
// separate top-row load cost about a quarter of the 5 us a block took, profiles/r2_experiments.md).
// grid = (bh * ceil(lanes/NBW)).  A CTA does not derive its row from blockIdx: it takes a ticket when it starts, and
// tickets are handed out row-major.  The CTA of the row above therefore holds a smaller ticket, i.e. it is already
// resident and running whenever this CTA waits for it -- forward progress does not depend on the order in which the
// hardware dispatches CTAs, nor on the whole grid being co-resident (grids beyond resident capacity, MPS, sanitizers).
template <int BS>
__global__ void __launch_bounds__(32) tq_iframe_kernel(TqArgs a, int lanes) {
    constexpr int NBW = 32 / BS;
    extern __shared__ __align__(16) uint8_t smraw[];
    struct ISmem {
        WarpTile<BS> t;
        uint8_t zz[BS * BS];
        __align__(16) uint8_t left[NBW][BS];
    };
    ISmem& sm = *reinterpret_cast<ISmem*>(smraw);
    const int lane = threadIdx.x;
    build_zigzag<BS>(sm.zz, lane, 32);
    __syncwarp();
    const int ngrp = (lanes + NBW - 1) / NBW;
    const int tk = wavefront_ticket(a.ticket, lane);
    const int by = a.row_begin + tk / ngrp, grp = tk % ngrp;
    const int q = lane / BS, x = lane % BS;
    const int fl_raw = grp * NBW + q;
    const bool valid = fl_raw < lanes;
    const int fl = valid ? fl_raw : lanes - 1;
    const FrameLane& L = a.lanes[fl];
    const int oy = by * BS;
    WarpTile<BS>& t = sm.t;
    uint8_t* recon_plane = a.ref_base + (size_t)L.out_plane * a.ref_plane_bytes;
    const uint8_t* cur_plane = a.cur_base + (size_t)L.cur_plane * a.cur_plane_bytes;
    const int qp = a.qp_rows[(size_t)fl * a.bh + by];
    // mailbox row `by` = bottom rows of block row by - 1 (written by the warp above), row by + 1 = ours for the warp below
    const volatile uint32_t* mail_up = a.top_mail + (((size_t)fl * a.bh + by) * a.bw) * BS + x;
    uint32_t* mail_dn = (by + 1 < a.bh && valid) ? a.top_mail + (((size_t)fl * a.bh + by + 1) * a.bw) * BS : nullptr;
    const uint32_t tag = a.epoch << 8;
    uint32_t cw_next[BS / 4];     // the next block's current pixels are requested one block ahead
    load_row_aligned<BS>(cur_plane + (size_t)(oy + x) * a.cur_pitch, cw_next);

    for (int bx = 0; bx < a.bw; bx++) {
        const int ox = bx * BS;
        uint32_t cw[BS / 4];
#pragma unroll
        for (int i = 0; i < BS / 4; i++) cw[i] = cw_next[i];
        if (bx + 1 < a.bw) load_row_aligned<BS>(cur_plane + (size_t)(oy + x) * a.cur_pitch + ox + BS, cw_next);
        // top neighbour: poll this lane's mailbox word until the block above has posted it for this frame; the left column
        // was left in shared memory by this warp's previous block
        int tv = 128;
        if (oy > 0 && valid) {
            uint32_t v = mail_up[bx * BS];
            while ((v & 0xffffff00u) != tag) v = mail_up[bx * BS];   // pure spin: __nanosleep(20) costs about a microsecond here (1.24 against 0.91 ms per I step)
            tv = (int)(v & 255u);
        }
        __syncwarp();
        if (ox == 0) sm.left[q][x] = 128;
        __syncwarp();
        uint32_t lw[BS / 4];
#pragma unroll
        for (int i = 0; i < BS / 4; i++) lw[i] = reinterpret_cast<const uint32_t*>(&sm.left[q][0])[i];
        int sh, sv;
        intra_mode_sads<BS>(cw, lw, tv, ox > 0, oy > 0, sh, sv);
        const int mode = (sh < sv) ? 0 : 1;  // tie -> vertical, IFrame.py:192-195
        {
            // mode 0: row x of pred = left[0..BS-1]; mode 1: row x = top[x] replicated
            uint32_t pw[BS / 4];
#pragma unroll
            for (int i = 0; i < BS / 4; i++) pw[i] = mode == 0 ? lw[i] : (uint32_t)tv * 0x01010101u;
            stage_row<BS>(t, q, x, cw, pw);
        }
        if (valid && x == 0) {
            a.modes[(size_t)fl * a.nblk + by * a.bw + bx] = mode;
            a.isad[(size_t)fl * a.nblk + by * a.bw + bx] = mode == 0 ? sh : sv;
        }
        __syncwarp();

        TqOut o;
        o.levels = a.levels ? a.levels + ((size_t)fl * a.H + oy) * a.W + ox : nullptr;
        o.lev_pitch = a.W;
        o.recon = recon_plane + (size_t)oy * a.ref_pitch + ox;
        o.rec_pitch = a.ref_pitch;
        o.resid_mc = a.resid_mc ? a.resid_mc + ((size_t)fl * a.H + oy) * a.W + ox : nullptr;
        o.resid_pitch = a.W;
        o.idct_out = nullptr;
        o.coef_out = nullptr;
        tq_warp<BS>(t, lane, valid, qp, o, nullptr, nullptr, true, &sm.left[0][0], mail_dn ? mail_dn + bx * BS : nullptr, tag);
    }
}

// The same wavefront with a block pair spread over the four schedulers of an SM (BS = 8, 16): one CTA of four warps per
// (block row, NBW lanes).  Warp 0 runs the serial part of a block -- poll the top row, decide the mode, form the
// residual -- then each of the four passes of the transform is split four ways (tq_device.cuh, quad_*), with a CTA barrier
// between passes; afterwards warp 0 posts the bottom row and the right column (what the next blocks wait for) while
// warps 1 and 2 write the reconstruction and the levels to the planes.  Bit-identical to tq_iframe_kernel: every output is
// the same fma chain.  Per 16x16 block pair ~3000 instead of ~9000 cycles (profiles/r2_experiments.md).
template <int BS>
__global__ void __launch_bounds__(128) tq_iframe_quad_kernel(TqArgs a, int lanes) {
    constexpr int NBW = 32 / BS;
    struct QSmem {
        QuadTile<BS> t;
        __align__(16) uint8_t left[NBW][BS];
        int ticket;
    };
    extern __shared__ __align__(16) uint8_t smraw[];
    QSmem& sm = *reinterpret_cast<QSmem*>(smraw);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        const int tk = atomicAdd(a.ticket, 1);
        if (tk == (int)gridDim.x - 1) atomicExch(a.ticket, 0);
        sm.ticket = tk;
    }
    __syncthreads();
    const int ngrp = (lanes + NBW - 1) / NBW;
    const int tk = sm.ticket;
    const int by = a.row_begin + tk / ngrp, grp = tk % ngrp;
    const int q = lane / BS, x = lane % BS;
    const int fl_raw = grp * NBW + q;
    const bool valid = fl_raw < lanes;
    const int fl = valid ? fl_raw : lanes - 1;
    const FrameLane& L = a.lanes[fl];
    const int oy = by * BS;
    QuadTile<BS>& t = sm.t;
    uint8_t* recon_plane = a.ref_base + (size_t)L.out_plane * a.ref_plane_bytes;
    const uint8_t* cur_plane = a.cur_base + (size_t)L.cur_plane * a.cur_plane_bytes;
    const int qp = a.qp_rows[(size_t)fl * a.bh + by];
    const volatile uint32_t* mail_up = a.top_mail + (((size_t)fl * a.bh + by) * a.bw) * BS + x;
    uint32_t* mail_dn = (by + 1 < a.bh && valid) ? a.top_mail + (((size_t)fl * a.bh + by + 1) * a.bw) * BS : nullptr;
    const uint32_t tag = a.epoch << 8;
    uint32_t cw_next[BS / 4];
    if (warp == 0) load_row_aligned<BS>(cur_plane + (size_t)(oy + x) * a.cur_pitch, cw_next);

    for (int bx = 0; bx < a.bw; bx++) {
        const int ox = bx * BS;
        if (warp == 0) {
            uint32_t cw[BS / 4];
#pragma unroll
            for (int i = 0; i < BS / 4; i++) cw[i] = cw_next[i];
            if (bx + 1 < a.bw) load_row_aligned<BS>(cur_plane + (size_t)(oy + x) * a.cur_pitch + ox + BS, cw_next);
            int tv = 128;
            if (oy > 0 && valid) {
                uint32_t v = mail_up[bx * BS];
                while ((v & 0xffffff00u) != tag) v = mail_up[bx * BS];
                tv = (int)(v & 255u);
            }
            __syncwarp();
            if (ox == 0) sm.left[q][x] = 128;   // else written by this warp at the end of the previous block
            __syncwarp();
            uint32_t lw[BS / 4];
#pragma unroll
            for (int i = 0; i < BS / 4; i++) lw[i] = reinterpret_cast<const uint32_t*>(&sm.left[q][0])[i];
            // mode decision, IFrame.py:184-195 (see intra_mode_sads)
            int sh, sv;
            intra_mode_sads<BS>(cw, lw, tv, ox > 0, oy > 0, sh, sv);
            const int mode = (sh < sv) ? 0 : 1;
            uint32_t pw[BS / 4];
#pragma unroll
            for (int i = 0; i < BS / 4; i++) pw[i] = mode == 0 ? lw[i] : (uint32_t)tv * 0x01010101u;
            stage_row<BS>(t, q, x, cw, pw);
            if (valid && x == 0) {
                a.modes[(size_t)fl * a.nblk + by * a.bw + bx] = mode;
                a.isad[(size_t)fl * a.nblk + by * a.bw + bx] = mode == 0 ? sh : sv;
            }
            if (a.resid_mc) __syncwarp();   // warp-uniform
            if (a.resid_mc && valid) {   // debug plane of the frame-level calls: the int16 residual as uint8 (IFrame.py:30,57-58)
                int8_t* d = a.resid_mc + ((size_t)fl * a.H + oy) * a.W + ox;
#pragma unroll
                for (int y = 0; y < BS; y++) d[(size_t)y * a.W + x] = (int8_t)(uint8_t)t.res[q][y][x];
            }
        }
        __syncthreads();
        BVC_QUAD_DISPATCH(quad_f1, warp, t, q, x)
        __syncthreads();
        BVC_QUAD_DISPATCH(quad_f2, warp, t, q, x, qp)
        __syncthreads();
        BVC_QUAD_DISPATCH(quad_i1, warp, t, q, x)
        __syncthreads();
        BVC_QUAD_DISPATCH(quad_i2, warp, t, q, x)
        __syncthreads();
        if (warp == 0) {
            // what the neighbours wait for: the bottom row for the block below (one tagged word per pixel, lane x -> pixel x),
            // the right column for this warp's next block
            if (mail_dn) __stcg(mail_dn + bx * BS + x, (uint32_t)t.rec[q][BS - 1][x] | tag);
            sm.left[q][x] = t.rec[q][x][BS - 1];
        } else if (warp == 1) {
            if (valid) {
                uint32_t ow[BS / 4];
                load_row_aligned<BS>(&t.rec[q][x][0], ow);
                store_row_words<BS>(recon_plane + (size_t)(oy + x) * a.ref_pitch + ox, ow);
            }
        } else if (warp == 2) {
            if (valid && a.levels) {
                const uint32_t* ls = reinterpret_cast<const uint32_t*>(&t.lev[q][x][0]);
                uint32_t* lg = reinterpret_cast<uint32_t*>(a.levels + ((size_t)fl * a.H + oy + x) * a.W + ox);
                if constexpr (BS == 16) {
                    *reinterpret_cast<uint4*>(lg) = *reinterpret_cast<const uint4*>(ls);
                    *reinterpret_cast<uint4*>(lg + 4) = *reinterpret_cast<const uint4*>(ls + 4);
                } else {
                    *reinterpret_cast<uint4*>(lg) = *reinterpret_cast<const uint4*>(ls);
                }
            }
        }
    }
}

// Entropy coding of I-frame blocks from the level plane the wavefront kernel wrote (frame layout, a.levels): one warp
// per block, every block independent -- kept out of the wavefront so that the serial chain through a frame is only
// predict -> transform -> reconstruct.  grid = (ceil(blocks / TQ_WARPS), lanes)
template <int BS>
__global__ void __launch_bounds__(TQ_WARPS * 32) entropy_levels_kernel(TqArgs a) {
    struct ESmem {
        __align__(16) int16_t lev[TQ_WARPS][BS * BS];
        uint32_t bits[TQ_WARPS][blk_words_for<BS>() + 4];
        uint8_t zz[BS * BS];
    };
    __shared__ ESmem sm;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    build_zigzag<BS>(sm.zz, threadIdx.x, blockDim.x);
    __syncthreads();
    const int fl = blockIdx.y;
    const int blk_begin = a.row_begin * a.bw, blk_end = (a.row_begin + a.row_count) * a.bw;
    const int b = blk_begin + blockIdx.x * TQ_WARPS + warp;
    if (b >= blk_end) return;
    const int ox = (b % a.bw) * BS, oy = (b / a.bw) * BS;
    const int16_t* src = a.levels + ((size_t)fl * a.H + oy) * a.W + ox;
    if constexpr (BS == 16) {   // 16 rows x 32 bytes: lane -> (row, half)
        const uint4 v = *reinterpret_cast<const uint4*>(src + (size_t)(lane >> 1) * a.W + (lane & 1) * 8);
        *reinterpret_cast<uint4*>(&sm.lev[warp][lane * 8]) = v;
    } else {
        for (int e = lane; e < BS * BS; e += 32) sm.lev[warp][e] = src[(size_t)(e / BS) * a.W + (e % BS)];
    }
    __syncwarp();
    uint32_t* gout = a.blk_bits + ((size_t)fl * a.nblk + b) * a.blk_words;
    const int nb = entropy_block_warp<BS>(sm.lev[warp], sm.zz, sm.bits[warp], lane, gout);
    if (lane == 0) a.blk_nbits[(size_t)fl * a.nblk + b] = nb;
}

// ---------------------------------------------------------------------------------------------
// Block-level test hook: dense int16 residual / pred blocks.
template <int BS>
__global__ void __launch_bounds__(TQ_WARPS * 32) tq_blocks_kernel(const int16_t* res, const int16_t* pred, int nblocks, int qp,
                                                                 int16_t* level, uint8_t* recon, double* idct, double* coef) {
    constexpr int NBW = 32 / BS;
    extern __shared__ __align__(16) uint8_t smraw[];
    TqCtaSmem<BS>& sm = *reinterpret_cast<TqCtaSmem<BS>*>(smraw);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    WarpTile<BS>& t = sm.w[warp];
    const int q = lane / BS;
    const int b = (blockIdx.x * TQ_WARPS + warp) * NBW + q;
    const bool valid = b < nblocks;
    const int bb = valid ? b : nblocks - 1;
    TqOut o;
    o.levels = level + (size_t)bb * BS * BS;
    o.lev_pitch = BS;
    o.recon = recon + (size_t)bb * BS * BS;
    o.rec_pitch = BS;
    o.resid_mc = nullptr;
    o.resid_pitch = 0;
    o.idct_out = idct ? idct + (size_t)bb * BS * BS : nullptr;
    o.coef_out = coef ? coef + (size_t)bb * BS * BS : nullptr;
    tq_warp<BS>(t, lane, valid, qp, o, res + (size_t)bb * BS * BS, pred + (size_t)bb * BS * BS, false);
}

template <int BS, bool DBG>
cudaError_t launch_pd(const TqArgs& a, int lanes, cudaStream_t st) {
    constexpr int NBW = 32 / BS;
    const size_t smem = sizeof(TqCtaSmem<BS>);
    static bool once_dev[BVC_MAX_DEVICES] = {};
    bool& once = once_dev[current_device_slot()];
    if (!once) {
        cudaError_t e = cudaFuncSetAttribute(tq_pframe_kernel<BS, DBG>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        e = cudaFuncSetAttribute(tq_pframe_kernel<BS, DBG>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
        if (e != cudaSuccess) return e;
        once = true;
    }
    const int nb = a.row_count * a.bw;
    const int units_x = (nb + TQ_WARPS * NBW - 1) / (TQ_WARPS * NBW), units = units_x * lanes;
    // persistent grid: the CTAs that can be resident at once (6 per SM), each warp drawing tasks until none is left
    static int slots_dev[BVC_MAX_DEVICES] = {};
    int& slots = slots_dev[current_device_slot()];
    if (slots == 0) {
        int per_sm = 0, dev = 0, sms = 0;
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, tq_pframe_kernel<BS, DBG>, TQ_WARPS * 32, smem) != cudaSuccess || per_sm < 1) per_sm = 1;
        cudaGetDevice(&dev);
        if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || sms < 1) sms = 148;
        slots = per_sm * sms;
    }
    if (!a.ticket) return cudaErrorInvalidValue;   // the task counter (zero between launches)
    int grid = std::min(units, slots);
    if (a.cta_cap > 0) grid = std::min(grid, a.cta_cap);
    TqArgs b = a;
    const int tasks_per_lane = units_x * TQ_WARPS;
    fast_div_constants((uint32_t)tasks_per_lane, b.ux_magic, b.ux_shift);
    fast_div_constants((uint32_t)a.bw, b.bw_magic, b.bw_shift);
    tq_pframe_kernel<BS, DBG><<<grid, TQ_WARPS * 32, smem, st>>>(b, tasks_per_lane, units * TQ_WARPS);
    return cudaGetLastError();
}
template <int BS>
cudaError_t launch_p(const TqArgs& a, int lanes, cudaStream_t st) {
    // the debug planes (residuals with / without motion compensation) exist only for the frame-level calls
    if (a.resid_mc || a.resid_nomc) return launch_pd<BS, true>(a, lanes, st);
    return launch_pd<BS, false>(a, lanes, st);
}

template <int BS>
cudaError_t launch_ie(const TqArgs& a, int lanes, cudaStream_t st) {
    if (!a.levels) return cudaErrorInvalidValue;
    dim3 grid((a.row_count * a.bw + TQ_WARPS - 1) / TQ_WARPS, lanes);
    entropy_levels_kernel<BS><<<grid, TQ_WARPS * 32, 0, st>>>(a);
    return cudaGetLastError();
}

template <int BS>
cudaError_t launch_i(const TqArgs& a, int lanes, cudaStream_t st, bool with_entropy) {
    constexpr int NBW = 32 / BS;
    const size_t smem = sizeof(WarpTile<BS>) + BS * BS + 2 * NBW * BS + 64;
    const int ngrp = (lanes + NBW - 1) / NBW;
    static bool once_dev[BVC_MAX_DEVICES] = {};
    bool& once = once_dev[current_device_slot()];
    if (!once) {
        cudaError_t e = cudaFuncSetAttribute(tq_iframe_kernel<BS>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
        if (e != cudaSuccess) return e;
        once = true;
    }
    if (!a.levels) return cudaErrorInvalidValue;   // the level plane feeds the entropy kernel
    if constexpr (BS >= 8) {
        // Four warps per block pair shorten the dependent chain (one 16x16 block step 4.8 -> 2.9 us), but the rows of a frame
        // run in lock step, so every CTA of an SM is in the same pass at the same time: beyond about two CTAs per SM the passes
        // queue for the SM's fp64 pipe and shared memory and the gain is gone (20 lanes of 1080p: 0.94 against 0.91 ms per I
        // step; 3 lanes: 0.54 against 1.11, profiles/r2_iframe_quad.jsonl).  8x8 blocks (80 registers, less work per pass)
        // gain at any size: 20 lanes 1.04 against 1.19-1.35 ms.
        static int sms_dev[BVC_MAX_DEVICES] = {};
        int& sms = sms_dev[current_device_slot()];
        if (sms == 0) {
            int dev = 0;
            cudaGetDevice(&dev);
            if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || sms < 1) sms = 148;
        }
        if (a.quad == 2 || (a.quad && (BS == 8 || a.row_count * ngrp <= 2 * sms))) tq_iframe_quad_kernel<BS><<<a.row_count * ngrp, 128, sizeof(QuadTile<BS>) + NBW * BS + 32, st>>>(a, lanes);
        else tq_iframe_kernel<BS><<<a.row_count * ngrp, 32, smem, st>>>(a, lanes);
    } else {
        tq_iframe_kernel<BS><<<a.row_count * ngrp, 32, smem, st>>>(a, lanes);
    }
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess || !with_entropy) return e;
    return launch_ie<BS>(a, lanes, st);
}

template <int BS>
cudaError_t launch_b(const int16_t* res, const int16_t* pred, int nblocks, int qp, int16_t* level, uint8_t* recon,
                     double* idct, double* coef, cudaStream_t st) {
    constexpr int NBW = 32 / BS;
    const size_t smem = sizeof(TqCtaSmem<BS>);
    static bool once_dev[BVC_MAX_DEVICES] = {};
    bool& once = once_dev[current_device_slot()];
    if (!once) {
        cudaError_t e = cudaFuncSetAttribute(tq_blocks_kernel<BS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        once = true;
    }
    const int grid = (nblocks + TQ_WARPS * NBW - 1) / (TQ_WARPS * NBW);
    tq_blocks_kernel<BS><<<grid, TQ_WARPS * 32, smem, st>>>(res, pred, nblocks, qp, level, recon, idct, coef);
    return cudaGetLastError();
}

}  // namespace

int tq_blk_words(int bs) { return bs == 16 ? blk_words_for<16>() : bs == 8 ? blk_words_for<8>() : blk_words_for<4>(); }

cudaError_t launch_tq_pframe(const TqArgs& a, int lanes, cudaStream_t st) {
    switch (a.bs) {
        case 16: return launch_p<16>(a, lanes, st);
        case 8: return launch_p<8>(a, lanes, st);
        case 4: return launch_p<4>(a, lanes, st);
    }
    return cudaErrorInvalidValue;
}
cudaError_t launch_tq_iframe(const TqArgs& a, int lanes, cudaStream_t st, bool with_entropy) {
    switch (a.bs) {
        case 16: return launch_i<16>(a, lanes, st, with_entropy);
        case 8: return launch_i<8>(a, lanes, st, with_entropy);
        case 4: return launch_i<4>(a, lanes, st, with_entropy);
    }
    return cudaErrorInvalidValue;
}
cudaError_t launch_tq_ientropy(const TqArgs& a, int lanes, cudaStream_t st) {
    switch (a.bs) {
        case 16: return launch_ie<16>(a, lanes, st);
        case 8: return launch_ie<8>(a, lanes, st);
        case 4: return launch_ie<4>(a, lanes, st);
    }
    return cudaErrorInvalidValue;
}
cudaError_t launch_tq_blocks(const int16_t* res, const int16_t* pred, int nblocks, int bs, int qp, int16_t* level,
                             uint8_t* recon, double* idct, double* coef, cudaStream_t st) {
    switch (bs) {
        case 16: return launch_b<16>(res, pred, nblocks, qp, level, recon, idct, coef, st);
        case 8: return launch_b<8>(res, pred, nblocks, qp, level, recon, idct, coef, st);
        case 4: return launch_b<4>(res, pred, nblocks, qp, level, recon, idct, coef, st);
    }
    return cudaErrorInvalidValue;
}

}  // namespace bvc
