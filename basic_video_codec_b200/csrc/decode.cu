// decode.cu -- the decoder: decode_video of the reference (decoder.py:26-87) on the GPU.
//
//   D1 eg_spec_kernel    exp-Golomb tokenizer, speculative pass: the bit streams of ALL frames of a clip are cut
//   D2 eg_chain_kernel   into 512-bit chunks; a warp walks one chunk from every possible entry offset (0..31 --
//   D3 eg_emit_kernel    a code is at most 31 bits here) and records exit offset / symbol count / EOB count; one
//                        thread per stream then chains the chunks (tables staged in shared memory) to get each
//                        chunk's true entry offset and symbol base; the emit pass re-walks every chunk from its
//                        true entry and writes the symbols and, for coefficient streams, the first symbol of
//                        every block (blocks end at the EOB marker 8190).
//                        Replaces exp_golomb_decode (encoder/entropy_encoder.py:32-62), the symbol loop of
//                        Frame.entropy_decode_dct_coffs (encoder/Frame.py:81-98) and entropy_decode in decoder.py:18-23.
//   D4 pred_decode_kernel  row QPs and differential MVs (prefix sum in raster order) or intra modes
//                        (encoder/PFrame.py:166-228, encoder/IFrame.py:132-166).
//   D5 dec_pframe_kernel RLE -> inverse zig-zag -> rescale -> IDCT -> + motion-compensated prediction -> clip
//                        (rle_decode / inverse_zigzag_order entropy_encoder.py:91-112,138-160;
//                        construct_frame_from_dct_and_mv encoder/PFrame.py:252-317).
//   D6 dec_iframe_kernel the same with intra prediction from the frame being rebuilt: anti-diagonal wavefront
//                        (IFrame.decode_mc_q_dct encoder/IFrame.py:85-114).
// Tokenizing does not depend on other frames, so D1-D4 run once for the whole clip; only D5/D6 follow the
// frame order inside a GOP (GOP lanes, like the encoder).
#include "bvc_kernels.h"
#include "tq_device.cuh"

namespace bvc {
namespace {

constexpr int EG_CHUNK_BITS = 512;
constexpr int EG_CHUNK_WORDS = EG_CHUNK_BITS / 32;
constexpr int EG_STAGE_WORDS = EG_CHUNK_WORDS + 3;   // a 64-bit window at the chunk's last bit; odd stride = no bank conflicts
constexpr int EG_MAX_Z = 15;                         // longest accepted prefix (31-bit code); the format needs 13

// Stage the chunk starting at byte `cb` of the container into big-endian words.  Aligned 32-bit loads + funnel
// shift; bytes past the stream's end are whatever follows in the buffer (the walk never trusts them, see eg_step).
__device__ __forceinline__ uint32_t stage_word(const uint8_t* data, long long cb, int w) {
    const uintptr_t ad = reinterpret_cast<uintptr_t>(data) + (uintptr_t)cb + 4u * (uintptr_t)w;
    const uint32_t* base = reinterpret_cast<const uint32_t*>(ad & ~(uintptr_t)3);
    const uint32_t sh = (uint32_t)(ad & 3) * 8;
    const uint32_t lo = base[0], hi = sh ? base[1] : 0u;
    return __byte_perm(__funnelshift_r(lo, hi, sh), 0, 0x0123);   // memory order -> MSB-first word
}
__device__ __forceinline__ unsigned long long window64(const uint32_t* sm, int p) {
    const int wi = p >> 5, sh = p & 31;
    const uint32_t a = sm[wi], b = sm[wi + 1], c = sm[wi + 2];
    return ((unsigned long long)__funnelshift_l(b, a, sh) << 32) | __funnelshift_l(c, b, sh);
}
// One exp_golomb_decode at bit `p` of the staged chunk with `rem` (> 0) bits left in the stream.
//  return 1: symbol (value, len)   0: end of stream (fewer than 8 zero padding bits, entropy_encoder.py:40-44)
//        -1: malformed ("Not enough bits", :45-46, or a code running past the end)
__device__ __forceinline__ int eg_step(const uint32_t* sm, int p, long long rem, int& value, int& len) {
    const unsigned long long w = window64(sm, p);
    const int z = __clzll((long long)w);
    if ((long long)z >= rem) return rem < 8 ? 0 : -1;
    if (z > EG_MAX_Z) return -1;
    len = 2 * z + 1;
    if ((long long)len > rem) return -1;
    const uint32_t v = (uint32_t)(w >> (64 - len)) - 1u;
    value = (v & 1u) ? (int)((v + 1u) >> 1) : -(int)(v >> 1);   // entropy_encoder.py:57
    return 1;
}

// The 32 walks of a chunk visit the same bit positions over and over (prefix codes resynchronise after a few symbols), and
// what a walk does at a position does not depend on where it came from.  So every position is decoded once -- 16 per
// lane -- into a table (length, EOB flag, or how the walk ends there), and the 32 walks just follow the table.
__global__ void __launch_bounds__(128) eg_spec_kernel(const uint8_t* data, const EgStream* streams, const int* chunk_stream,
                                                      long long chunk_begin, long long nchunks, uint8_t* exit_tab, uint16_t* nsym_tab,
                                                      uint8_t* neob_tab) {
    __shared__ uint32_t sm[4][EG_STAGE_WORDS + 1];
    // per position: bits 0-4 code length (1..31, always odd); bit 5 EOB marker; 0x40 = end of stream, 0x80 = malformed
    __shared__ uint8_t tab[4][EG_CHUNK_BITS];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const long long gc = chunk_begin + (long long)blockIdx.x * 4 + warp;   // chunks [chunk_begin, chunk_begin + nchunks): one step's streams
    if (gc >= chunk_begin + nchunks) return;
    const EgStream st = streams[chunk_stream[gc]];
    const long long bit0 = (gc - st.chunk0) * EG_CHUNK_BITS;
    if (lane < EG_STAGE_WORDS) sm[warp][lane] = stage_word(data, st.byte0 + (bit0 >> 3), lane);
    __syncwarp();
    const long long left = st.nbits - bit0;
    const int end = (int)min((long long)EG_CHUNK_BITS, left);
    // lane l decodes positions l, l + 32, ...: neighbouring lanes read neighbouring windows (no bank conflicts on `sm`)
    for (int p = lane; p < end; p += 32) {
        int v, len;
        const int k = eg_step(sm[warp], p, left - p, v, len);
        tab[warp][p] = k > 0 ? (uint8_t)(len | (v == BVC_EOB_MARKER ? 0x20 : 0)) : (k < 0 ? 0x80 : 0x40);
    }
    __syncwarp();
    int pos = lane, nsym = 0, neob = 0;
    bool err = false;
    while (pos < end) {
        const uint32_t t = tab[warp][pos];
        if (t & 0xc0u) { err = (t & 0x80u) != 0; pos = EG_CHUNK_BITS; break; }
        nsym++;
        neob += (int)(t >> 5);
        pos += (int)(t & 31u);
    }
    exit_tab[gc * 32 + lane] = err ? 255 : (uint8_t)max(pos - EG_CHUNK_BITS, 0);
    nsym_tab[gc * 32 + lane] = (uint16_t)nsym;
    neob_tab[gc * 32 + lane] = (uint8_t)neob;
}

// chunk -> stream map, built on the device from the streams' chunk ranges (1.2 M entries for the headline clip: as a host
// array it cost a fill and a 5 MB pageable upload in front of everything else).  One CTA per stream.
__global__ void __launch_bounds__(256) eg_chunk_map_kernel(const EgStream* streams, int* chunk_stream) {
    const EgStream st = streams[blockIdx.x];
    const long long nch = (st.nbits + EG_CHUNK_BITS - 1) / EG_CHUNK_BITS;
    for (long long c = threadIdx.x; c < nch; c += blockDim.x) chunk_stream[st.chunk0 + c] = (int)blockIdx.x;
}

constexpr int CHAIN_TILE = 128;
__global__ void __launch_bounds__(128) eg_chain_kernel(EgStream* streams, const int* stream_list, const uint8_t* exit_tab,
                                                       const uint16_t* nsym_tab, const uint8_t* neob_tab, uint8_t* entry_tab, int* symbase,
                                                       int* eobbase, int* err_flag) {
    __shared__ uint32_t s_exit[CHAIN_TILE * 8];
    __shared__ uint32_t s_nsym[CHAIN_TILE * 16];
    __shared__ uint32_t s_neob[CHAIN_TILE * 8];
    EgStream& st = streams[stream_list[blockIdx.x]];
    const long long nch = (st.nbits + EG_CHUNK_BITS - 1) / EG_CHUNK_BITS;
    int e = 0, S = 0, E = 0;
    bool bad = false;
    for (long long c0 = 0; c0 < nch; c0 += CHAIN_TILE) {
        const int n = (int)min((long long)CHAIN_TILE, nch - c0);
        const long long g0 = st.chunk0 + c0;
        const uint32_t* ge = reinterpret_cast<const uint32_t*>(exit_tab + g0 * 32);     // tables are 32-byte rows: word aligned
        const uint32_t* gn = reinterpret_cast<const uint32_t*>(nsym_tab + g0 * 32);
        const uint32_t* gb = reinterpret_cast<const uint32_t*>(neob_tab + g0 * 32);
        for (int i = threadIdx.x; i < n * 8; i += blockDim.x) { s_exit[i] = ge[i]; s_neob[i] = gb[i]; }
        for (int i = threadIdx.x; i < n * 16; i += blockDim.x) s_nsym[i] = gn[i];
        __syncthreads();
        if (threadIdx.x == 0 && !bad) {
            const uint8_t* xe = reinterpret_cast<const uint8_t*>(s_exit);
            const uint16_t* xn = reinterpret_cast<const uint16_t*>(s_nsym);
            const uint8_t* xb = reinterpret_cast<const uint8_t*>(s_neob);
            for (int c = 0; c < n; c++) {
                entry_tab[g0 + c] = (uint8_t)e;
                symbase[g0 + c] = S;
                eobbase[g0 + c] = E;
                S += xn[c * 32 + e];
                E += xb[c * 32 + e];
                e = xe[c * 32 + e];
                if (e > 31) { bad = true; for (int k = c + 1; k < n; k++) entry_tab[g0 + k] = 255; break; }
            }
        } else if (threadIdx.x == 0) {
            for (int c = 0; c < n; c++) entry_tab[g0 + c] = 255;
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        st.nsym = S;
        st.neob = E;
        if (bad) atomicExch(err_flag, 1);
    }
}

// Where the symbols of one step's streams go: consecutive runs in the step's slab of the symbol ring, in list order.  The
// counts come from the chain pass, so the host never has to see them (no round trip between tokenizing and decoding).  A
// coefficient stream must hold exactly one EOB-terminated run per block (the reference's decoder would run out of symbols);
// frames that do not are marked so that the rebuild kernels leave them alone.
__global__ void eg_offsets_kernel(EgStream* streams, const int* stream_list, int n, long long slab_base, long long* coef_sym0,
                                  uint8_t* frame_ok, int nblk, int pred_only, int* err_flag) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    long long off = slab_base;
    for (int i = 0; i < n; i++) {
        EgStream& s = streams[stream_list[i]];
        s.sym0 = off;
        off += s.nsym;
        if (s.kind == 1) {
            coef_sym0[s.frame] = s.sym0;
            const bool ok = s.neob == nblk;
            frame_ok[s.frame] = ok ? 1 : 0;
            if (!ok && !pred_only) atomicExch(err_flag, 1);
        }
    }
}

__global__ void __launch_bounds__(128) eg_emit_kernel(const uint8_t* data, const EgStream* streams, const int* chunk_stream,
                                                      long long chunk_begin, long long nchunks, const uint8_t* entry_tab,
                                                      const int* symbase, const int* eobbase, int16_t* syms, int* blk_start, int nblk) {
    __shared__ uint32_t sm[128 * EG_STAGE_WORDS];
    const long long gc = chunk_begin + (long long)blockIdx.x * 128 + threadIdx.x;
    if (gc >= chunk_begin + nchunks) return;
    const EgStream st = streams[chunk_stream[gc]];
    const long long bit0 = (gc - st.chunk0) * EG_CHUNK_BITS;
    uint32_t* my = sm + threadIdx.x * EG_STAGE_WORDS;
#pragma unroll
    for (int w = 0; w < EG_STAGE_WORDS; w++) my[w] = stage_word(data, st.byte0 + (bit0 >> 3), w);
    int* bs = (st.kind == 1) ? blk_start + (size_t)st.frame * (nblk + 1) : nullptr;
    if (bs && gc == st.chunk0) bs[0] = 0;
    int pos = entry_tab[gc];
    if (pos > 31) return;
    const long long left = st.nbits - bit0;
    const int end = (int)min((long long)EG_CHUNK_BITS, left);
    int si = symbase[gc], eb = eobbase[gc];
    int16_t* out = syms + st.sym0;
    while (pos < end) {
        int v, len;
        if (eg_step(my, pos, left - pos, v, len) <= 0) break;
        out[si++] = (int16_t)v;
        if (bs && v == BVC_EOB_MARKER) { eb++; if (eb <= nblk) bs[eb] = si; }
        pos += len;
    }
}

// ---------------------------------------------------------------------------------------------
// D4: prediction data of one frame.  Symbol layout per block row: EG(qp - base), then per block spb symbols.
__global__ void __launch_bounds__(256) pred_decode_kernel(const EgStream* streams, const int16_t* syms, const uint8_t* intra_flags,
                                                          const int* frame_list, int4* mv_all, int32_t* modes_all, int32_t* qp_all,
                                                          int bw, int bh, int base_qp, int with_ref, int* err_flag) {
    __shared__ int wsum[3][9];
    const int f = frame_list[blockIdx.x], tid = threadIdx.x, nblk = bw * bh;
    const EgStream st = streams[2 * f];             // stream 2f = prediction data, 2f+1 = coefficients
    const int16_t* s = syms + st.sym0;
    const bool intra = intra_flags[f] != 0;
    const int spb = intra ? 1 : (with_ref ? 3 : 2);
    const int per_row = 1 + bw * spb;
    if (st.nsym < bh * per_row) {                   // the reference runs out of bits: ValueError / TypeError
        if (tid == 0) atomicExch(err_flag, 1);
        return;
    }
    for (int r = tid; r < bh; r += blockDim.x) qp_all[(size_t)f * bh + r] = base_qp + (int)s[r * per_row];
    if (intra) {
        for (int b = tid; b < nblk; b += blockDim.x) {
            const int m = s[(b / bw) * per_row + 1 + (b % bw)];
            if (m != 0 && m != 1) atomicExch(err_flag, 1);   // find_intra_predict_block raises ValueError (IFrame.py:175-182)
            modes_all[(size_t)f * nblk + b] = m;
        }
        return;
    }
    // mv[b] = sum of the differences of blocks 0..b (PFrame.py:204, prev_mv chained in raster order)
    const int chunk = (nblk + blockDim.x - 1) / blockDim.x;
    const int b0 = min(tid * chunk, nblk), b1 = min(b0 + chunk, nblk);
    int acc[3] = {0, 0, 0};
    for (int b = b0; b < b1; b++) {
        const int16_t* p = s + (b / bw) * per_row + 1 + (b % bw) * spb;
        acc[0] += p[0];
        acc[1] += p[1];
        if (with_ref) acc[2] += p[2];
    }
    int pre[3];
    const int lane = tid & 31, warp = tid >> 5;
#pragma unroll
    for (int k = 0; k < 3; k++) {
        int incl = acc[k];
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const int o = __shfl_up_sync(0xffffffffu, incl, d);
            if (lane >= d) incl += o;
        }
        if (lane == 31) wsum[k][warp] = incl;
        pre[k] = incl - acc[k];
    }
    __syncthreads();
    if (tid < 3) {
        int run = 0;
        for (int w = 0; w < 8; w++) { const int t = wsum[tid][w]; wsum[tid][w] = run; run += t; }
    }
    __syncthreads();
    int cur[3] = {pre[0] + wsum[0][warp], pre[1] + wsum[1][warp], pre[2] + wsum[2][warp]};
    for (int b = b0; b < b1; b++) {
        const int16_t* p = s + (b / bw) * per_row + 1 + (b % bw) * spb;
        cur[0] += p[0];
        cur[1] += p[1];
        if (with_ref) cur[2] += p[2];
        mv_all[(size_t)f * nblk + b] = make_int4(cur[0], cur[1], cur[2], 0);
    }
}

// ---------------------------------------------------------------------------------------------
// Shared block-level pieces of D5 / D6.

// rle_decode + pad_with_zeros + inverse_zigzag_order for block `b` of frame `f`, by one lane, into the zeroed tile.
template <int BS>
__device__ __forceinline__ void rle_to_tile(const DecArgs& a, int f, int b, const uint8_t* zz, int16_t* tile) {
    const int* bst = a.blk_start + (size_t)f * (a.nblk + 1);
    const int16_t* s = a.syms + a.coef_sym0[f];
    int i = bst[b];
    const int e = bst[b + 1] - 1;   // the block's symbols end before its EOB marker
    int pos = 0;
    while (i < e) {
        const int c = s[i];
        if (c == 0) break;                       // "rest is zero"
        if (c > 0) { pos += c; i++; continue; }  // run of zeros
        const int k = -c;
        for (int j = 0; j < k && i + 1 + j < e; j++, pos++)
            if (pos < BS * BS) tile[zz[pos]] = s[i + 1 + j];
        i += k + 1;
    }
}

// rescale_block + apply_idct_2d + reconstruction (dct.py:15-18,40-42; PFrame.py:305-308 / IFrame.py:107-108) of the
// NBW blocks whose levels are in t.lev and predictions in t.pred.  Same arithmetic, in the same order, as the
// second half of tq_warp (the encoder's reconstruction), so decode(encode(x)) == the encoder's reconstruction bit for bit.
template <int BS>
__device__ __forceinline__ void dequant_idct_recon_warp(WarpTile<BS>& t, int lane, bool valid, int qp, uint8_t* recon, int rec_pitch,
                                                        int16_t* levels, int lev_pitch, uint8_t* last_col = nullptr,
                                                        uint32_t* mail_row = nullptr, uint32_t mail_tag = 0) {
    const int q = lane / BS, x = lane % BS, u = x;
    double a[BS], r[BS];
    const bool su = (u == 0) || (2 * u == BS);
    const double w_sp = su ? DctC<BS>::w(0) : DctC<BS>::w(1);
    const double w_nm = su ? DctC<BS>::w(1) : DctC<BS>::w(2);
#pragma unroll
    for (int v = 0; v < BS; v++) {
        const bool sv = (v == 0) || (2 * v == BS);
        const double w = sv ? w_sp : w_nm;
        const int s = qp + __vimin_s32_relu(u + v - (BS - 2), 2);
        const short lv = t.lev[q][u][v];
        if (levels && valid) levels[(size_t)u * lev_pitch + v] = lv;
        const double wq = __hiloint2double(__double2hiint(w) + (s << 20), __double2loint(w));
        a[v] = __dmul_rn((double)(int)lv, wq);
    }
#pragma unroll
    for (int v = 0; v < BS; v++) t.buf[q][u][v] = a[v];
    __syncwarp();
#pragma unroll
    for (int i = 0; i < BS; i++) a[i] = t.buf[q][i][x];
    fold_inv<BS>(a, r);
#pragma unroll
    for (int y = 0; y < BS; y++) t.buf[q][y][x] = r[y];
    __syncwarp();
    const int y = x;
#pragma unroll
    for (int i = 0; i < BS; i++) a[i] = t.buf[q][y][i];
    fold_inv<BS>(a, r);
    if (valid) {
        uint32_t pw[BS / 4], ow[BS / 4];
        load_row_aligned<BS>(&t.pred[q][y][0], pw);
#pragma unroll
        for (int i = 0; i < BS; i++) {
            const int pb = (int)((pw[i >> 2] >> (8 * (i & 3))) & 255u);
            const int v = (int)(short)(int)rint(__dadd_rn(r[i], (double)pb));
            const uint32_t c8 = (uint32_t)__vimin_s32_relu(v, 255);
            if ((i & 3) == 0) ow[i >> 2] = c8; else ow[i >> 2] |= c8 << (8 * (i & 3));
        }
        store_row_words<BS>(recon + (size_t)y * rec_pitch, ow);
        if (last_col) last_col[q * BS + y] = (uint8_t)(ow[BS / 4 - 1] >> 24);   // right column for the next block of the row
        // bottom row for the block below: every pixel as a tagged word the consumer polls (see tq_iframe_kernel)
        if (mail_row && y == BS - 1) {
#pragma unroll
            for (int i = 0; i < BS; i += 4) {
                uint4 v;
                v.x = ((ow[i >> 2]) & 255u) | mail_tag;
                v.y = ((ow[i >> 2] >> 8) & 255u) | mail_tag;
                v.z = ((ow[i >> 2] >> 16) & 255u) | mail_tag;
                v.w = (ow[i >> 2] >> 24) | mail_tag;
                __stcg(reinterpret_cast<uint4*>(mail_row + i), v);
            }
        }
    }
    __syncwarp();
}

template <int BS>
__device__ __forceinline__ void zero_lev(WarpTile<BS>& t, int lane) {
    constexpr int NBW = 32 / BS;
    uint4* p = reinterpret_cast<uint4*>(&t.lev[0][0][0]);
    constexpr int NV = NBW * BS * BS * 2 / 16;
    for (int i = lane; i < NV; i += 32) p[i] = make_uint4(0, 0, 0, 0);
}

constexpr int DEC_WARPS = 4;
template <int BS>
struct DecCtaSmem {
    WarpTile<BS> w[DEC_WARPS];
    uint8_t zz[BS * BS];
};

// D5: P frames, blocks independent.  grid = (ceil(nblk / (DEC_WARPS*NBW)), lanes)
template <int BS>
__global__ void __launch_bounds__(DEC_WARPS * 32, 4) dec_pframe_kernel(DecArgs a) {
    constexpr int NBW = 32 / BS;
    extern __shared__ __align__(16) uint8_t smraw[];
    DecCtaSmem<BS>& sm = *reinterpret_cast<DecCtaSmem<BS>*>(smraw);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    build_zigzag<BS>(sm.zz, threadIdx.x, blockDim.x);
    __syncthreads();
    WarpTile<BS>& t = sm.w[warp];
    const FrameLane& L = a.lanes[blockIdx.y];
    const int f = L.slot;
    if (!a.frame_ok[f]) return;   // malformed coefficient stream (the call fails): its block table cannot be trusted
    const int q = lane / BS, x = lane % BS;
    const int b = (blockIdx.x * DEC_WARPS + warp) * NBW + q;
    const bool valid = b < a.nblk;
    const int bb = valid ? b : a.nblk - 1;
    const int bx = bb % a.bw, by = bb / a.bw;
    const int ox = bx * BS, oy = by * BS;
    zero_lev<BS>(t, lane);
    __syncwarp();
    if (valid && x == 0) rle_to_tile<BS>(a, f, bb, sm.zz, &t.lev[q][0][0]);
    // find_mv_predicted_block PFrame.py:230-244 (refs[mv[2]] only with more than one reference in the window)
    const int4 mv = a.mv_all[(size_t)f * a.nblk + bb];
    const int k = (L.nref > 1) ? mv.z : 0;
    int dx = mv.x, dy = mv.y, ph = 0;
    if (a.frac) { ph = (mv.x & 1) | ((mv.y & 1) << 1); dx = mv.x >> 1; dy = mv.y >> 1; }
    // is_out_of_range block_predictor.py:116-143 raises ValueError; here the frame is flagged and the block predicted from 0
    bool inr = k >= 0 && k < L.nref && ox + dx >= 0 && oy + dy >= 0;
    if (a.frac) inr = inr && (2 * ox + mv.x + 2 * BS <= 2 * a.W) && (2 * oy + mv.y + 2 * BS <= 2 * a.H);
    else inr = inr && (ox + dx + BS <= a.W) && (oy + dy + BS <= a.H);
    uint32_t pw[BS / 4];
#pragma unroll
    for (int i = 0; i < BS / 4; i++) pw[i] = 0;
    if (inr) {
        const uint8_t* pr = a.ref_base + (size_t)(L.ref_plane[k] + ph) * a.ref_plane_bytes + (size_t)(oy + dy + x) * a.ref_pitch + (ox + dx);
        load_row_unaligned<BS>(pr, pw);
    } else if (valid && x == 0) {
        atomicExch(a.err_flag, 1);
    }
    store_row_words<BS>(&t.pred[q][x][0], pw);
    __syncwarp();
    const int qp = a.qp_all[(size_t)f * a.bh + by];
    uint8_t* recon = a.ref_base + (size_t)L.out_plane * a.ref_plane_bytes + (size_t)oy * a.ref_pitch + ox;
    int16_t* lev = a.levels_out ? a.levels_out + ((size_t)f * a.H + oy) * a.W + ox : nullptr;
    dequant_idct_recon_warp<BS>(t, lane, valid, qp, recon, a.ref_pitch, lev, a.W);
}

// D6: I frames.  Same wavefront as tq_iframe_kernel (one warp per block row of NBW different frames; a block's bottom row
// is handed to the row below through epoch-tagged mailbox words, the right column through shared memory), with the mode
// read from the stream instead of decided.
template <int BS>
__global__ void __launch_bounds__(32) dec_iframe_kernel(DecArgs a, int lanes) {
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
    const int tk = wavefront_ticket(a.ticket, lane);   // start order, not blockIdx: see tq_iframe_kernel
    const int by = tk / ngrp, grp = tk % ngrp;
    const int q = lane / BS, x = lane % BS;
    const int fl_raw = grp * NBW + q;
    const bool in_range = fl_raw < lanes;
    const int fl = in_range ? fl_raw : lanes - 1;
    const FrameLane& L = a.lanes[fl];
    const int f = L.slot;
    const bool valid = in_range && a.frame_ok[f];   // a malformed frame (the call fails) is left alone, all its rows alike
    const int oy = by * BS;
    WarpTile<BS>& t = sm.t;
    uint8_t* recon_plane = a.ref_base + (size_t)L.out_plane * a.ref_plane_bytes;
    const int qp = a.qp_all[(size_t)f * a.bh + by];
    const volatile uint32_t* mail_up = a.top_mail + (((size_t)fl * a.bh + by) * a.bw) * BS + x;
    uint32_t* mail_dn = (by + 1 < a.bh && valid) ? a.top_mail + (((size_t)fl * a.bh + by + 1) * a.bw) * BS : nullptr;
    const uint32_t tag = a.epoch << 8;
    for (int bx = 0; bx < a.bw; bx++) {
        const int ox = bx * BS, b = by * a.bw + bx;
        zero_lev<BS>(t, lane);
        __syncwarp();
        if (valid && x == 0) rle_to_tile<BS>(a, f, b, sm.zz, &t.lev[q][0][0]);
        const int mode = a.modes_all[(size_t)f * a.nblk + b];
        // find_intra_predict_block IFrame.py:175-213: mode 0 -> pred[r][c] = recon[oy+c][ox-1] (the right column this warp's
        // previous block left in shared memory); mode 1 -> recon[oy-1][ox+r] (this lane's mailbox word, posted by the row above)
        int tv = 128;
        if (oy > 0 && valid) {
            uint32_t v = mail_up[bx * BS];
            while ((v & 0xffffff00u) != tag) v = mail_up[bx * BS];   // pure spin: __nanosleep(20) sleeps for about a microsecond
            tv = (int)(v & 255u);
        }
        if (ox == 0) sm.left[q][x] = 128;
        __syncwarp();
        uint32_t pw[BS / 4];
#pragma unroll
        for (int i = 0; i < BS / 4; i++) pw[i] = mode == 0 ? reinterpret_cast<const uint32_t*>(&sm.left[q][0])[i] : (uint32_t)tv * 0x01010101u;
        store_row_words<BS>(&t.pred[q][x][0], pw);
        __syncwarp();
        int16_t* lev = a.levels_out ? a.levels_out + ((size_t)f * a.H + oy) * a.W + ox : nullptr;
        dequant_idct_recon_warp<BS>(t, lane, valid, qp, recon_plane + (size_t)oy * a.ref_pitch + ox, a.ref_pitch, lev, a.W, &sm.left[0][0],
                                    mail_dn ? mail_dn + bx * BS : nullptr, tag);
    }
}

// fill planes with a constant (the decoder's initial 128 reference, decoder.py:35)
__global__ void fill_plane_kernel(uint8_t* p, size_t n, uint8_t v) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) p[i] = v;
}

template <int BS>
cudaError_t launch_dec_p(const DecArgs& a, int lanes, cudaStream_t st) {
    constexpr int NBW = 32 / BS;
    const size_t smem = sizeof(DecCtaSmem<BS>);
    static bool once_dev[BVC_MAX_DEVICES] = {};
    bool& once = once_dev[current_device_slot()];
    if (!once) {
        cudaError_t e = cudaFuncSetAttribute(dec_pframe_kernel<BS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        once = true;
    }
    dim3 grid((a.nblk + DEC_WARPS * NBW - 1) / (DEC_WARPS * NBW), lanes);
    dec_pframe_kernel<BS><<<grid, DEC_WARPS * 32, smem, st>>>(a);
    return cudaGetLastError();
}
template <int BS>
cudaError_t launch_dec_i(const DecArgs& a, int lanes, cudaStream_t st) {
    constexpr int NBW = 32 / BS;
    const size_t smem = sizeof(WarpTile<BS>) + BS * BS + NBW * BS + 64;
    const int ngrp = (lanes + NBW - 1) / NBW;
    dec_iframe_kernel<BS><<<a.bh * ngrp, 32, smem, st>>>(a, lanes);
    return cudaGetLastError();
}

}  // namespace

int eg_chunk_bits() { return EG_CHUNK_BITS; }

cudaError_t launch_eg_tokenize_spec(const uint8_t* data, EgStream* streams, const int* stream_list, int nstreams, const int* chunk_stream,
                                    long long chunk_begin, long long nchunks, uint8_t* exit_tab, uint16_t* nsym_tab, uint8_t* neob_tab,
                                    uint8_t* entry_tab, int* symbase, int* eobbase, int* err_flag, cudaStream_t st) {
    if (nchunks > 0) eg_spec_kernel<<<(unsigned)((nchunks + 3) / 4), 128, 0, st>>>(data, streams, chunk_stream, chunk_begin, nchunks, exit_tab, nsym_tab, neob_tab);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return e;
    if (nstreams > 0) eg_chain_kernel<<<nstreams, 128, 0, st>>>(streams, stream_list, exit_tab, nsym_tab, neob_tab, entry_tab, symbase, eobbase, err_flag);
    return cudaGetLastError();
}
cudaError_t launch_eg_chunk_map(const EgStream* streams, int nstreams, int* chunk_stream, cudaStream_t st) {
    if (nstreams > 0) eg_chunk_map_kernel<<<nstreams, 256, 0, st>>>(streams, chunk_stream);
    return cudaGetLastError();
}
cudaError_t launch_eg_offsets(EgStream* streams, const int* stream_list, int nstreams, long long slab_base, long long* coef_sym0,
                              uint8_t* frame_ok, int nblk, int pred_only, int* err_flag, cudaStream_t st) {
    eg_offsets_kernel<<<1, 32, 0, st>>>(streams, stream_list, nstreams, slab_base, coef_sym0, frame_ok, nblk, pred_only, err_flag);
    return cudaGetLastError();
}
cudaError_t launch_eg_tokenize_emit(const uint8_t* data, const EgStream* streams, const int* chunk_stream, long long chunk_begin,
                                    long long nchunks, const uint8_t* entry_tab, const int* symbase, const int* eobbase, int16_t* syms,
                                    int* blk_start, int nblk, cudaStream_t st) {
    if (nchunks > 0) eg_emit_kernel<<<(unsigned)((nchunks + 127) / 128), 128, 0, st>>>(data, streams, chunk_stream, chunk_begin, nchunks, entry_tab, symbase, eobbase, syms, blk_start, nblk);
    return cudaGetLastError();
}
cudaError_t launch_pred_decode(const EgStream* streams, const int16_t* syms, const uint8_t* intra_flags, const int* frame_list, int nframes,
                               int4* mv_all, int32_t* modes_all, int32_t* qp_all, int bw, int bh, int base_qp, int with_ref, int* err_flag,
                               cudaStream_t st) {
    if (nframes > 0) pred_decode_kernel<<<nframes, 256, 0, st>>>(streams, syms, intra_flags, frame_list, mv_all, modes_all, qp_all, bw, bh, base_qp, with_ref, err_flag);
    return cudaGetLastError();
}
cudaError_t launch_dec_pframe(const DecArgs& a, int lanes, cudaStream_t st) {
    switch (a.bs) {
        case 16: return launch_dec_p<16>(a, lanes, st);
        case 8: return launch_dec_p<8>(a, lanes, st);
        case 4: return launch_dec_p<4>(a, lanes, st);
    }
    return cudaErrorInvalidValue;
}
cudaError_t launch_dec_iframe(const DecArgs& a, int lanes, cudaStream_t st) {
    switch (a.bs) {
        case 16: return launch_dec_i<16>(a, lanes, st);
        case 8: return launch_dec_i<8>(a, lanes, st);
        case 4: return launch_dec_i<4>(a, lanes, st);
    }
    return cudaErrorInvalidValue;
}
cudaError_t launch_fill_plane(uint8_t* p, size_t n, uint8_t v, cudaStream_t st) {
    fill_plane_kernel<<<296, 256, 0, st>>>(p, n, v);
    return cudaGetLastError();
}

}  // namespace bvc
