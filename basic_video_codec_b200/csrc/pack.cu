// pack.cu -- K7b: assemble the two per-frame bit streams in raster order.
//
//   prediction stream: per block row EG(row_qp - base_qp), then per block the differential
//       motion vector EG(dmvx) EG(dmvy) [EG(dref)] chained in raster order from (0,0,0)
//       (reference encoder/PFrame.py:136-163) or EG(mode) (encoder/IFrame.py:116-130);
//   coefficient stream: the per-block strings produced by the transform kernels, concatenated
//       (encoder/Frame.py:61-75).
// Streams are MSB-first bit strings, zero padded to whole bytes (bitarray.tobytes()).
//
// Three launches per step, every block of every frame in parallel (a frame used to be scanned by ONE CTA: 32 us per
// launch at 1080p / 16x16, four times that with 8x8 blocks -- 29 % of the kernel time of 1080p i=8 r=2):
//   pack_sums_kernel: tiles of 1024 blocks; per block the bits of its coefficient string and of its prediction symbols,
//     tile-local exclusive scans, tile totals.
//   pack_base_kernel: one CTA per frame; scan of the tile totals (a few dozen values), frame totals, overflow check,
//     zeroing of the used part of both streams.
//   pack_emit_kernel: one thread per block; funnel-shifts the block's words to its bit offset and ORs them into the
//     coefficient stream, ORs its prediction symbols into the prediction stream, and at every row start the row's bit
//     total (bits_per_row, PFrame.py:76-83).
#include "bvc_kernels.h"

namespace bvc {
namespace {

constexpr int PACK_TILE = 1024;   // blocks per tile of the first pass

__device__ __forceinline__ void put_bits_global_be(uint32_t* stream, long long off, unsigned long long code, int len) {
    const unsigned long long V = code << (64 - len);
    const int sh = (int)(off & 31);
    const long long wi = off >> 5;
    const uint32_t hi = (uint32_t)(V >> 32), lo = (uint32_t)V;
    const uint32_t w0 = hi >> sh;
    const uint32_t w1 = sh ? ((hi << (32 - sh)) | (lo >> sh)) : lo;
    const uint32_t w2 = sh ? (lo << (32 - sh)) : 0u;
    // store byte-swapped so that memory order is the MSB-first byte stream
    if (w0) atomicOr(&stream[wi], __byte_perm(w0, 0, 0x0123));
    if (w1) atomicOr(&stream[wi + 1], __byte_perm(w1, 0, 0x0123));
    if (w2) atomicOr(&stream[wi + 2], __byte_perm(w2, 0, 0x0123));
}

// block-wide exclusive scan of one 64-bit value per thread; returns the exclusive prefix, total in *tot
__device__ __forceinline__ long long block_exscan(long long v, long long* warp_sums, long long* tot) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    long long incl = v;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        long long o = __shfl_up_sync(0xffffffffu, incl, d);
        if (lane >= d) incl += o;
    }
    if (lane == 31) warp_sums[warp] = incl;
    __syncthreads();
    if (warp == 0) {
        long long w = (lane < (int)(blockDim.x >> 5)) ? warp_sums[lane] : 0;
        long long wi = w;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            long long o = __shfl_up_sync(0xffffffffu, wi, d);
            if (lane >= d) wi += o;
        }
        warp_sums[lane] = wi - w;  // exclusive
        if (lane == 31) warp_sums[32] = wi;
    }
    __syncthreads();
    const long long res = warp_sums[warp] + incl - v;
    *tot = warp_sums[32];
    __syncthreads();
    return res;
}

struct PredCode {
    unsigned long long code;  // up to 57 bits of motion-vector / mode symbols
    int len;
    uint32_t qcode;           // row-start qp symbol
    int qlen;
};

__device__ __forceinline__ PredCode pred_code(const PackArgs& a, int fl, int b) {
    PredCode pc;
    pc.qcode = 0;
    pc.qlen = 0;
    const int bx = b % a.bw, by = b / a.bw;
    if (bx == 0) {  // EG(rc_qp - base qp) opens every block row (PFrame.py:146-147, IFrame.py:120-121)
        pc.qcode = eg_code(a.qp_rows[(size_t)fl * a.bh + by] - a.base_qp);
        pc.qlen = eg_len_of_code(pc.qcode);
    }
    if (a.intra) {
        const uint32_t e = eg_code(a.modes[(size_t)fl * a.nblk + b]);
        pc.code = e;
        pc.len = eg_len_of_code(e);
    } else {
        const int4 m = a.mv[(size_t)fl * a.nblk + b];
        int4 p = make_int4(0, 0, 0, 0);
        if (b > 0) p = a.mv[(size_t)fl * a.nblk + b - 1];  // previous block in raster order (PFrame.py:140-144,163)
        uint32_t e = eg_code(m.x - p.x);
        int l = eg_len_of_code(e);
        pc.code = e;
        pc.len = l;
        e = eg_code(m.y - p.y);
        l = eg_len_of_code(e);
        pc.code = (pc.code << l) | e;
        pc.len += l;
        if (a.with_ref) {
            e = eg_code(m.z - p.z);
            l = eg_len_of_code(e);
            pc.code = (pc.code << l) | e;
            pc.len += l;
        }
    }
    return pc;
}

// Frame totals of both streams are known: close the tile table, publish the sizes, check the slot, and zero the used
// part of both streams (+2 words of slack for the 3-word OR window).  Called by every thread of the CTA.
__device__ __forceinline__ void pack_frame_totals(const PackArgs& a, int fl, long long ctot, long long ptot) {
    const int tid = threadIdx.x;
    long long* tb = a.tile_base + (size_t)fl * (a.tiles + 1) * 2;
    const int slot = a.lanes[fl].slot;   // per-frame slot in the stream arena
    if (tid == 0) {
        if (a.tiles == 1) { tb[0] = 0; tb[1] = 0; }
        tb[2 * a.tiles] = ctot; tb[2 * a.tiles + 1] = ptot;   // frame totals close the table
        a.frame_bits[2 * slot] = ptot; a.frame_bits[2 * slot + 1] = ctot;
        // the slots are sized for 6 bits per pixel by default, not for the worst case: tell the host instead of overrunning
        if (a.slot_overflow && (((ctot + 31) >> 5) + 2 > (long long)a.coef_cap_words || ((ptot + 31) >> 5) + 2 > (long long)a.pred_cap_words))
            atomicExch(a.slot_overflow, 1);
    }
    uint32_t* cstream = a.coef_stream + (size_t)slot * a.coef_cap_words;
    uint32_t* pstream = a.pred_stream + (size_t)slot * a.pred_cap_words;
    const long long cw = min((long long)a.coef_cap_words, ((ctot + 31) >> 5) + 2);
    const long long pw = min((long long)a.pred_cap_words, ((ptot + 31) >> 5) + 2);
    for (long long w = tid; w < cw; w += blockDim.x) cstream[w] = 0;
    for (long long w = tid; w < pw; w += blockDim.x) pstream[w] = 0;
}

__global__ void __launch_bounds__(PACK_TILE) pack_sums_kernel(PackArgs a) {
    __shared__ long long warp_sums[33];
    const int fl = blockIdx.y, tile = blockIdx.x, tid = threadIdx.x;
    const int b = tile * PACK_TILE + tid;
    const bool valid = b < a.nblk;
    long long cb = 0, pb = 0;
    if (valid) {
        cb = a.blk_nbits[(size_t)fl * a.nblk + b];
        const PredCode pc = pred_code(a, fl, b);
        pb = pc.len + pc.qlen;
    }
    long long ctot, ptot;
    const long long cpre = block_exscan(cb, warp_sums, &ctot);
    const long long ppre = block_exscan(pb, warp_sums, &ptot);
    if (valid) {
        a.coef_off[(size_t)fl * (a.nblk + 1) + b] = cpre;     // tile-local; pack_emit adds the tile's base
        a.pred_off[(size_t)fl * a.nblk + b] = (int32_t)ppre;
    }
    if (tid == 0) {
        long long* tt = a.tile_tot + ((size_t)fl * (a.tiles + 1) + tile) * 2;
        tt[0] = ctot; tt[1] = ptot;
    }
    if (a.tiles == 1) pack_frame_totals(a, fl, ctot, ptot);   // small frames: the only tile is the frame, no second pass
}

__global__ void __launch_bounds__(1024) pack_base_kernel(PackArgs a) {
    __shared__ long long warp_sums[33];
    const int fl = blockIdx.x, tid = threadIdx.x;
    const long long* tt = a.tile_tot + (size_t)fl * (a.tiles + 1) * 2;
    long long* tb = a.tile_base + (size_t)fl * (a.tiles + 1) * 2;
    const int chunk = (a.tiles + (int)blockDim.x - 1) / (int)blockDim.x;
    const int t0 = min(a.tiles, tid * chunk), t1 = min(a.tiles, t0 + chunk);
    long long cs = 0, ps = 0;
    for (int t = t0; t < t1; t++) { cs += tt[2 * t]; ps += tt[2 * t + 1]; }
    long long ctot, ptot;
    long long cpre = block_exscan(cs, warp_sums, &ctot);
    long long ppre = block_exscan(ps, warp_sums, &ptot);
    for (int t = t0; t < t1; t++) {
        tb[2 * t] = cpre; tb[2 * t + 1] = ppre;
        cpre += tt[2 * t]; ppre += tt[2 * t + 1];
    }
    pack_frame_totals(a, fl, ctot, ptot);
}

__device__ __forceinline__ int rc_pick_qp(const RcArgs& rc, double budget) {
    for (int i = 0; i < rc.n; i++)
        if ((double)rc.bits[i] <= budget) return rc.qp[i];   // find_rc_qp_for_row, RateControl.py:34-43
    return rc.qp[rc.n - 1];
}

// bits of one block row: coefficient strings of its blocks + its prediction symbols; with rate control the row is
// charged to the frame's budget and the next row's QP is chosen (Frame.py:168-188, PFrame.py:53-83, IFrame.py:38-70)
__global__ void __launch_bounds__(256) pack_row_bits_kernel(PackArgs a, RcArgs rc, int row, long long* out) {
    __shared__ long long warp_sums[33];
    const int fl = blockIdx.x;
    long long s = 0;
    for (int bx = threadIdx.x; bx < a.bw; bx += blockDim.x) {
        const int b = row * a.bw + bx;
        const PredCode pc = pred_code(a, fl, b);
        s += a.blk_nbits[(size_t)fl * a.nblk + b] + pc.len + pc.qlen;
    }
    long long tot;
    block_exscan(s, warp_sums, &tot);
    if (threadIdx.x == 0) {
        if (out) out[fl] = tot;
        if (rc.n > 0) {
            const double rem = rc.remaining[fl] - (double)tot;
            rc.remaining[fl] = rem;
            if (row + 1 < a.bh) rc.qp_rows[(size_t)fl * a.bh + row + 1] = rc_pick_qp(rc, rem / (double)(a.bh - (row + 1)));
        }
    }
}

__global__ void rc_begin_kernel(RcArgs rc, int lanes, int bh) {
    const int fl = blockIdx.x * blockDim.x + threadIdx.x;
    if (fl >= lanes) return;
    rc.remaining[fl] = rc.frame_budget;
    rc.qp_rows[(size_t)fl * bh] = rc_pick_qp(rc, rc.frame_budget / (double)bh);
}

// One thread per block: after quantisation a block's string is a handful of words (about 125 bits at the headline QP), so
// a warp per block would leave most lanes idle; dense blocks (up to 203 words) just loop longer.
constexpr int EMIT_THREADS = 256;
__global__ void __launch_bounds__(EMIT_THREADS) pack_emit_kernel(PackArgs a) {
    const int fl = blockIdx.y;
    const int b = blockIdx.x * EMIT_THREADS + threadIdx.x;
    if (b >= a.nblk) return;
    const long long* tb = a.tile_base + (size_t)fl * (a.tiles + 1) * 2;
    const long long* cl = a.coef_off + (size_t)fl * (a.nblk + 1);
    const int32_t* pl = a.pred_off + (size_t)fl * a.nblk;
    const int tile = b / PACK_TILE;
    const int slot = a.lanes[fl].slot;
    // ---- coefficient string ----
    const int nbits = a.blk_nbits[(size_t)fl * a.nblk + b];
    const long long D = tb[2 * tile] + cl[b];
    const uint32_t* src = a.blk_bits + ((size_t)fl * a.nblk + b) * a.blk_words;
    uint32_t* dst = a.coef_stream + (size_t)slot * a.coef_cap_words;
    const int n = (nbits + 31) >> 5;
    const int sh = (int)(D & 31);
    const long long w0 = D >> 5;
    uint32_t hi = 0;
    for (int j = 0; j <= n; j++) {
        const uint32_t lo = (j < n) ? src[j] : 0u;
        const uint32_t v = sh ? ((hi << (32 - sh)) | (lo >> sh)) : lo;
        if (v && w0 + j < (long long)a.coef_cap_words) atomicOr(&dst[w0 + j], __byte_perm(v, 0, 0x0123));
        hi = lo;
    }
    // ---- prediction symbols ----
    const PredCode pc = pred_code(a, fl, b);
    const long long off = tb[2 * tile + 1] + pl[b];
    uint32_t* pstream = a.pred_stream + (size_t)slot * a.pred_cap_words;
    if (((off + pc.qlen + pc.len + 31) >> 5) + 2 <= (long long)a.pred_cap_words) {
        if (pc.qlen) put_bits_global_be(pstream, off, pc.qcode, pc.qlen);
        put_bits_global_be(pstream, off + pc.qlen, pc.code, pc.len);
    }
    // ---- bits_per_row (PFrame.py:76-83): growth of both streams while the row was coded ----
    if (b % a.bw == 0) {
        const int r = b / a.bw, b2 = b + a.bw;
        long long end;
        if (b2 < a.nblk) {
            const int t2 = b2 / PACK_TILE;
            end = tb[2 * t2] + cl[b2] + tb[2 * t2 + 1] + pl[b2];
        } else {
            end = tb[2 * a.tiles] + tb[2 * a.tiles + 1];
        }
        a.row_bits[(size_t)fl * a.bh + r] = end - (D + off);
    }
}

// ---- container assembly (encoder.py:104-121) on the device ---------------------------------------
// Per frame: mode byte | 2-byte BE prediction length | prediction bytes | 3-byte BE coefficient length |
// coefficient bytes.  container_scan_kernel turns the per-frame bit counts into byte offsets (one CTA);
// container_copy_kernel writes headers and copies payloads, so the host receives the finished
// encoded.bin image with a single download.
__global__ void __launch_bounds__(1024) container_scan_kernel(ContainerArgs a) {
    __shared__ long long warp_sums[33];
    const int tid = threadIdx.x;
    const int chunk = (a.nframes + 1023) / 1024;
    const int f0 = tid * chunk, f1 = min(a.nframes, f0 + chunk);
    long long sum = 0;
    int bad = 0;
    for (int f = f0; f < f1; f++) {
        const long long pb = (a.frame_bits[2 * f] + 7) >> 3, cb = (a.frame_bits[2 * f + 1] + 7) >> 3;
        if (pb > 0xFFFF || cb > 0xFFFFFF) bad = 1;   // OverflowError in the reference (int.to_bytes)
        sum += 6 + pb + cb;
    }
    long long tot;
    long long pre = block_exscan(sum, warp_sums, &tot);
    for (int f = f0; f < f1; f++) {
        a.frame_off[f] = pre;
        pre += 6 + ((a.frame_bits[2 * f] + 7) >> 3) + ((a.frame_bits[2 * f + 1] + 7) >> 3);
    }
    if (tid == 0) a.frame_off[a.nframes] = tot;
    if (bad) atomicExch(a.overflow, 1);
}

__global__ void __launch_bounds__(256) container_copy_kernel(ContainerArgs a) {
  for (int f = blockIdx.y; f < a.nframes; f += gridDim.y) {
    const long long pb = (a.frame_bits[2 * f] + 7) >> 3, cb = (a.frame_bits[2 * f + 1] + 7) >> 3;
    const long long off = a.frame_off[f];
    if (off + 6 + pb + cb > a.out_cap) continue;     // host reports BVC_ERR_NOMEM from frame_off[nframes]
    if (pb > 4 * (long long)a.pred_cap_words || cb > 4 * (long long)a.coef_cap_words) continue;   // slot overflow, reported by pack_scan
    uint8_t* o = a.out + off;
    const uint8_t* ps = reinterpret_cast<const uint8_t*>(a.pred_stream + (size_t)f * a.pred_cap_words);
    const uint8_t* cs = reinterpret_cast<const uint8_t*>(a.coef_stream + (size_t)f * a.coef_cap_words);
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        o[0] = (uint8_t)((f % a.i_period == 0) ? 1 : 0);   // PredictionMode: INTRA_FRAME = 1, INTER_FRAME = 0
        o[1] = (uint8_t)(pb >> 8); o[2] = (uint8_t)pb;
        o[3 + pb] = (uint8_t)(cb >> 16); o[4 + pb] = (uint8_t)(cb >> 8); o[5 + pb] = (uint8_t)cb;
    }
    const long long stride = (long long)gridDim.x * blockDim.x;
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    for (long long i = t; i < pb; i += stride) o[3 + i] = ps[i];
    uint8_t* oc = o + 6 + pb;
    // coefficient payload: 16-byte vector copies once the destination is aligned
    const long long head = min(cb, (long long)((16 - ((uintptr_t)oc & 15)) & 15));
    for (long long i = t; i < head; i += stride) oc[i] = cs[i];
    const long long nvec = (cb - head) >> 4;
    for (long long v = t; v < nvec; v += stride) {
        const uint8_t* src = cs + head + (v << 4);
        uint4 val;
        if ((((uintptr_t)src) & 3) == 0) {
            const uint32_t* s4 = reinterpret_cast<const uint32_t*>(src);
            val = make_uint4(s4[0], s4[1], s4[2], s4[3]);
        } else {
            uint32_t w[4];
#pragma unroll
            for (int k = 0; k < 4; k++) w[k] = src[4 * k] | (src[4 * k + 1] << 8) | (src[4 * k + 2] << 16) | ((uint32_t)src[4 * k + 3] << 24);
            val = make_uint4(w[0], w[1], w[2], w[3]);
        }
        *reinterpret_cast<uint4*>(oc + head + (v << 4)) = val;
    }
    for (long long i = head + (nvec << 4) + t; i < cb; i += stride) oc[i] = cs[i];
  }
}

}  // namespace

int pack_tiles(int nblk) { return (nblk + PACK_TILE - 1) / PACK_TILE; }

cudaError_t launch_pack(const PackArgs& a, int lanes, cudaStream_t st) {
    if (a.tiles != pack_tiles(a.nblk) || !a.pred_off || !a.tile_tot || !a.tile_base) return cudaErrorInvalidValue;
    pack_sums_kernel<<<dim3(a.tiles, lanes), PACK_TILE, 0, st>>>(a);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return e;
    if (a.tiles > 1) {
        pack_base_kernel<<<lanes, 1024, 0, st>>>(a);
        e = cudaGetLastError();
        if (e != cudaSuccess) return e;
    }
    dim3 grid((a.nblk + EMIT_THREADS - 1) / EMIT_THREADS, lanes);
    pack_emit_kernel<<<grid, EMIT_THREADS, 0, st>>>(a);
    return cudaGetLastError();
}

}  // namespace bvc

namespace bvc {
cudaError_t launch_row_bits(const PackArgs& a, int lanes, int row, long long* out, cudaStream_t st) {
    RcArgs off{};
    pack_row_bits_kernel<<<lanes, 256, 0, st>>>(a, off, row, out);
    return cudaGetLastError();
}
cudaError_t launch_row_bits_rc(const PackArgs& a, const RcArgs& rc, int lanes, int row, long long* out, cudaStream_t st) {
    pack_row_bits_kernel<<<lanes, 256, 0, st>>>(a, rc, row, out);
    return cudaGetLastError();
}
cudaError_t launch_rc_begin(const RcArgs& rc, int lanes, int bh, cudaStream_t st) {
    rc_begin_kernel<<<(lanes + 63) / 64, 64, 0, st>>>(rc, lanes, bh);
    return cudaGetLastError();
}
cudaError_t launch_container(const ContainerArgs& a, cudaStream_t st) {
    container_scan_kernel<<<1, 1024, 0, st>>>(a);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return e;
    dim3 grid(32, a.nframes < 65535 ? a.nframes : 65535);
    container_copy_kernel<<<grid, 256, 0, st>>>(a);
    return cudaGetLastError();
}
}  // namespace bvc
