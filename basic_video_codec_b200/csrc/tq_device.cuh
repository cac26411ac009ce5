// tq_device.cuh -- warp-level device code shared by the P-frame and I-frame kernels:
//   K5  residual -> defined fp64 DCT -> quantise -> rescale -> IDCT -> reconstruct
//       (reference encoder/Frame.py:190-202, encoder/dct.py:9-42; arithmetic defined in DESIGN.md §DCT
//        and restated on the CPU in oracle/bvc_oracle.c -- the two must agree bit for bit)
//   K7a zig-zag -> RLE -> signed exp-Golomb of one block's levels + end-of-block marker
//       (reference encoder/Frame.py:61-75, encoder/entropy_encoder.py:8-29,65-88,115-135)
//
// One warp owns NBW = 32/BS blocks side by side: lane = (q = block in warp, x = column/row index).
// Every 1-D pass gives each lane one line of BS values and all BS outputs of it, BS/2 DFMA each with
// compile-time table indices (operands come from the constant bank, no shared-memory traffic for the
// cosines).  Lines are exchanged through a padded fp64 shared-memory tile between passes.
#pragma once
#include "bvc_common.cuh"
#include "bvc_dct_tables.h"
#include "bvc_kernels.h"

namespace bvc {

// const-qualified device tables: every index is a compile-time constant after unrolling, so the
// compiler folds the loads and the cosines become immediate / constant-bank operands of the DFMAs.
__device__ static const double c_ct4[16] = BVC_CT4_INIT;
__device__ static const double c_ct8[64] = BVC_CT8_INIT;
__device__ static const double c_ct16[256] = BVC_CT16_INIT;
__device__ static const double c_w4[3] = BVC_W4_INIT;
__device__ static const double c_w8[3] = BVC_W8_INIT;
__device__ static const double c_w16[3] = BVC_W16_INIT;

template <int BS> struct DctC;
template <> struct DctC<4>  { static __device__ __forceinline__ double ct(int i) { return c_ct4[i]; }  static __device__ __forceinline__ double w(int i) { return c_w4[i]; } };
template <> struct DctC<8>  { static __device__ __forceinline__ double ct(int i) { return c_ct8[i]; }  static __device__ __forceinline__ double w(int i) { return c_w8[i]; } };
template <> struct DctC<16> { static __device__ __forceinline__ double ct(int i) { return c_ct16[i]; } static __device__ __forceinline__ double w(int i) { return c_w16[i]; } };

template <int BS>
__host__ __device__ constexpr int blk_words_for() {
    // worst case: every coefficient non-zero with the largest magnitude 255*BS
    return BS == 16 ? 208 : BS == 8 ? 48 : 12;
}

// Shared-memory working set of one warp.
template <int BS>
struct WarpTile {
    static constexpr int NBW = 32 / BS;
    double buf[NBW][BS][BS + 1];
    __align__(16) int16_t lev[NBW][BS][BS];
    __align__(16) int16_t res[NBW][BS][BS];   // residual cur - pred (PFrame.py:248 / IFrame.py:222)
    __align__(16) uint8_t cur[NBW][BS][BS];
    __align__(16) uint8_t pred[NBW][BS][BS];
    uint32_t bits[blk_words_for<BS>() + 4];
};

// out[u] = sum_{x<BS/2} ct[u][x] * (u even ? a[x]+a[BS-1-x] : a[x]-a[BS-1-x])
template <int BS>
__device__ __forceinline__ void fold_fwd(const double (&a)[BS], double (&out)[BS]) {
    constexpr int Hh = BS / 2;
    double s[Hh], d[Hh];
#pragma unroll
    for (int x = 0; x < Hh; x++) { s[x] = __dadd_rn(a[x], a[BS - 1 - x]); d[x] = __dsub_rn(a[x], a[BS - 1 - x]); }
#pragma unroll
    for (int u = 0; u < BS; u++) {
        double acc = 0.0;
#pragma unroll
        for (int x = 0; x < Hh; x++) acc = __fma_rn(DctC<BS>::ct(u * BS + x), (u & 1) ? d[x] : s[x], acc);
        out[u] = acc;
    }
}
// out[y] = E+O, out[BS-1-y] = E-O with E/O the even/odd-u fma chains
template <int BS>
__device__ __forceinline__ void fold_inv(const double (&v)[BS], double (&out)[BS]) {
    constexpr int Hh = BS / 2;
#pragma unroll
    for (int y = 0; y < Hh; y++) {
        double e = 0.0, o = 0.0;
#pragma unroll
        for (int u = 0; u < BS; u += 2) e = __fma_rn(DctC<BS>::ct(u * BS + y), v[u], e);
#pragma unroll
        for (int u = 1; u < BS; u += 2) o = __fma_rn(DctC<BS>::ct(u * BS + y), v[u], o);
        out[y] = __dadd_rn(e, o);
        out[BS - 1 - y] = __dsub_rn(e, o);
    }
}

// Integer <-> double without the conversion unit (XU pipe: I2F / F2I / FRND.F64 issue at a quarter of the fp64 rate).
// 1.5 * 2^52 + t rounds t to the nearest integer, ties to even (the ulp there is 1), exactly like rint() for |t| < 2^51;
// the low mantissa word of that sum is the integer in two's complement.  2^52 + 2^31 + n has n + 2^31 in its low word.
__device__ __forceinline__ double rint_magic(double t, int& as_int) {
    const double m = __dadd_rn(t, 6755399441055744.0);
    as_int = __double2loint(m);
    return __dsub_rn(m, 6755399441055744.0);
}
__device__ __forceinline__ double int_to_double(int n) {   // exact for any int32
    return __dsub_rn(__hiloint2double(0x43300000, (int)(0x80000000u ^ (unsigned)n)), 4503601774854144.0);
}
__device__ __forceinline__ double pow2_neg(int s) { return __hiloint2double((1023 - s) << 20, 0); }
// Quantise and rescale in one go: np.round(coef / 2^s) (dct.py:35-37, half to even; the division is exact) and level * 2^s
// (dct.py:40-42).  1.5 * 2^(52+s) + coef rounds coef to the nearest multiple of 2^s, ties to the even multiple -- the same
// rounding as rint(coef * 2^-s), because scaling by a power of two commutes with round-to-nearest -- the low mantissa word
// of the sum is the level in two's complement, and subtracting the constant again leaves level * 2^s exactly.
// magic_hi = 0x43380000 + (s << 20).  Two fp64 additions instead of multiply + add + subtract (+ a multiply to rescale).
__device__ __forceinline__ double quant_rescale(double coef, int magic_hi, int& level) {
    const double M = __hiloint2double(magic_hi, 0);
    const double m = __dadd_rn(coef, M);
    level = __double2loint(m);
    return __dsub_rn(m, M);
}
__device__ __forceinline__ double pow2_pos(int s) { return __hiloint2double((1023 + s) << 20, 0); }

struct TqOut {
    int16_t* levels;   // global, frame layout, pitch W (elements) -- may be null
    int lev_pitch;
    uint8_t* recon;    // global plane base for this block row/col, pitch
    int rec_pitch;
    int8_t* resid_mc;  // may be null; pitch W
    int resid_pitch;
    double* idct_out;  // block hook only (dense bs*bs), may be null
    double* coef_out;  // block hook only
};

// vector helpers for one block row of BS bytes / BS int16
template <int BS> struct RowVec;
template <> struct RowVec<16> { using B = uint4; static constexpr int NW = 4; };
template <> struct RowVec<8>  { using B = uint2; static constexpr int NW = 2; };
template <> struct RowVec<4>  { using B = uint32_t; static constexpr int NW = 1; };

template <int BS>
__device__ __forceinline__ void load_row_aligned(const uint8_t* p, uint32_t (&w)[BS / 4]) {
    typename RowVec<BS>::B v = *reinterpret_cast<const typename RowVec<BS>::B*>(p);
    if constexpr (BS == 16) { w[0] = v.x; w[1] = v.y; w[2] = v.z; w[3] = v.w; }
    else if constexpr (BS == 8) { w[0] = v.x; w[1] = v.y; }
    else { w[0] = v; }
}
// BS bytes from an arbitrary byte address: aligned 32-bit loads + funnel shifts
template <int BS>
__device__ __forceinline__ void load_row_unaligned(const uint8_t* p, uint32_t (&w)[BS / 4]) {
    const uintptr_t ad = reinterpret_cast<uintptr_t>(p);
    const uint32_t* base = reinterpret_cast<const uint32_t*>(ad & ~(uintptr_t)3);
    const uint32_t sh = (uint32_t)(ad & 3) * 8;
    uint32_t raw[BS / 4 + 1];
#pragma unroll
    for (int i = 0; i <= BS / 4; i++) raw[i] = base[i];
#pragma unroll
    for (int i = 0; i < BS / 4; i++) w[i] = __funnelshift_r(raw[i], raw[i + 1], sh);
}
template <int BS>
__device__ __forceinline__ void store_row_words(void* dst, const uint32_t (&w)[BS / 4]) {
    if constexpr (BS == 16) *reinterpret_cast<uint4*>(dst) = make_uint4(w[0], w[1], w[2], w[3]);
    else if constexpr (BS == 8) *reinterpret_cast<uint2*>(dst) = make_uint2(w[0], w[1]);
    else *reinterpret_cast<uint32_t*>(dst) = w[0];
}

// Stage row r of block q: cur and pred bytes (as words) -> t.cur / t.pred / t.res (int16 residual).
template <int BS, typename Tile>
__device__ __forceinline__ void stage_row(Tile& t, int q, int r, const uint32_t (&cw)[BS / 4], const uint32_t (&pw)[BS / 4]) {
    store_row_words<BS>(&t.cur[q][r][0], cw);
    store_row_words<BS>(&t.pred[q][r][0], pw);
    uint32_t rw[BS / 2];
#pragma unroll
    for (int i = 0; i < BS / 4; i++) {
        // bytes -> two packed int16 differences per word pair
        const uint32_t c_lo = __byte_perm(cw[i], 0, 0x4140), c_hi = __byte_perm(cw[i], 0, 0x4342);
        const uint32_t p_lo = __byte_perm(pw[i], 0, 0x4140), p_hi = __byte_perm(pw[i], 0, 0x4342);
        rw[2 * i] = __vsub2(c_lo, p_lo);
        rw[2 * i + 1] = __vsub2(c_hi, p_hi);
    }
    uint32_t* d = reinterpret_cast<uint32_t*>(&t.res[q][r][0]);
    if constexpr (BS == 16) {
        *reinterpret_cast<uint4*>(d) = make_uint4(rw[0], rw[1], rw[2], rw[3]);
        *reinterpret_cast<uint4*>(d + 4) = make_uint4(rw[4], rw[5], rw[6], rw[7]);
    } else if constexpr (BS == 8) {
        *reinterpret_cast<uint4*>(d) = make_uint4(rw[0], rw[1], rw[2], rw[3]);
    } else {
        *reinterpret_cast<uint2*>(d) = make_uint2(rw[0], rw[1]);
    }
}

// Transform + quantise + reconstruct the NBW blocks held in `t` (t.res and t.pred staged, or dense
// int16 residual / pred given by the overrides for the block hook).  lane -> (q, x).  `valid` masks a
// warp's trailing blocks.  qp: quantisation parameter of this block row.  Writes the lev tile (smem)
// and the outputs in `o`.  intra_u8_resid: I frames store the raw int16 residual as uint8 in the debug
// plane (IFrame.py:30,57-58).
// DBG = false compiles the debug outputs (residual planes, idct / coefficient dumps) out: the clip path never asks for them.
template <int BS, bool DBG = true>
__device__ __forceinline__ void tq_warp(WarpTile<BS>& t, int lane, bool valid, int qp, const TqOut& o,
                                        const int16_t* res_override, const int16_t* pred_override, bool intra_u8_resid,
                                        uint8_t* last_col = nullptr, uint32_t* mail_row = nullptr, uint32_t mail_tag = 0) {
    const int q = lane / BS, x = lane % BS;
    double a[BS], r[BS];
    // ---- forward pass 1: columns (apply_dct_2d transforms columns first, dct.py:12) ----
#pragma unroll
    for (int y = 0; y < BS; y++) a[y] = int_to_double(res_override ? (int)res_override[y * BS + x] : (int)t.res[q][y][x]);
    if (DBG && intra_u8_resid && o.resid_mc && valid) {
#pragma unroll
        for (int y = 0; y < BS; y++) o.resid_mc[(size_t)y * o.resid_pitch + x] = (int8_t)(uint8_t)(res_override ? (int)res_override[y * BS + x] : (int)t.res[q][y][x]);
    }
    fold_fwd<BS>(a, r);
#pragma unroll
    for (int u = 0; u < BS; u++) t.buf[q][u][x] = r[u];
    __syncwarp();
    // ---- forward pass 2: rows; this lane owns row u = x ----
    const int u = x;
#pragma unroll
    for (int i = 0; i < BS; i++) a[i] = t.buf[q][u][i];
    fold_fwd<BS>(a, r);
    const bool su = (u == 0) || (2 * u == BS);
    const double w_sp = su ? DctC<BS>::w(0) : DctC<BS>::w(1);  // for v in {0, BS/2}
    const double w_nm = su ? DctC<BS>::w(1) : DctC<BS>::w(2);
    short lv[BS];
#pragma unroll
    for (int v = 0; v < BS; v++) {
        const bool sv = (v == 0) || (2 * v == BS);
        const double w = sv ? w_sp : w_nm;
        const double coef = __dmul_rn(r[v], w);
        if (DBG && o.coef_out && valid) o.coef_out[u * BS + v] = coef;
        // generate_quantization_matrix dct.py:21-32: shift s = qp + {0,1,2} for u+v <,=,> BS-1;
        // quantize_block :35-37 = round half to even of coef * 2^-s (exact scaling)
        const int s = qp + __vimin_s32_relu(u + v - (BS - 2), 2);
        int li;
        const double resc = quant_rescale(coef, 0x43380000 + (s << 20), li);
        lv[v] = (short)li;
        a[v] = __dmul_rn(resc, w);   // rescale_block dct.py:40-42 is exact: (level * 2^s) * w, one rounding
    }
    {
        uint32_t pk[BS / 2];
#pragma unroll
        for (int v = 0; v < BS / 2; v++) pk[v] = (uint32_t)(uint16_t)lv[2 * v] | ((uint32_t)(uint16_t)lv[2 * v + 1] << 16);
        uint32_t* ls = reinterpret_cast<uint32_t*>(&t.lev[q][u][0]);
        uint32_t* lg = (o.levels && valid) ? reinterpret_cast<uint32_t*>(o.levels + (size_t)u * o.lev_pitch) : nullptr;
        if constexpr (BS == 16) {
            *reinterpret_cast<uint4*>(ls) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
            *reinterpret_cast<uint4*>(ls + 4) = make_uint4(pk[4], pk[5], pk[6], pk[7]);
            if (lg) { *reinterpret_cast<uint4*>(lg) = make_uint4(pk[0], pk[1], pk[2], pk[3]); *reinterpret_cast<uint4*>(lg + 4) = make_uint4(pk[4], pk[5], pk[6], pk[7]); }
        } else if constexpr (BS == 8) {
            *reinterpret_cast<uint4*>(ls) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
            if (lg) *reinterpret_cast<uint4*>(lg) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
        } else {
            *reinterpret_cast<uint2*>(ls) = make_uint2(pk[0], pk[1]);
            if (lg) *reinterpret_cast<uint2*>(lg) = make_uint2(pk[0], pk[1]);
        }
    }
#pragma unroll
    for (int v = 0; v < BS; v++) t.buf[q][u][v] = a[v];
    __syncwarp();
    // ---- inverse pass 1: over u for column v = x ----
#pragma unroll
    for (int i = 0; i < BS; i++) a[i] = t.buf[q][i][x];
    fold_inv<BS>(a, r);
#pragma unroll
    for (int y = 0; y < BS; y++) t.buf[q][y][x] = r[y];
    __syncwarp();
    // ---- inverse pass 2: over v for row y = x ----
    const int y = x;
#pragma unroll
    for (int i = 0; i < BS; i++) a[i] = t.buf[q][y][i];
    fold_inv<BS>(a, r);
    if (valid) {
        uint32_t pw[BS / 4];
        if (!pred_override) load_row_aligned<BS>(&t.pred[q][y][0], pw);
        uint32_t ow[BS / 4];
#pragma unroll
        for (int i = 0; i < BS; i++) {
            const int pb = pred_override ? (int)pred_override[y * BS + i] : (int)((pw[i >> 2] >> (8 * (i & 3))) & 255u);
            // reconstruct_block Frame.py:197-202: round(idct + pred) -> int16 -> clip -> uint8
            int vi;
            rint_magic(__dadd_rn(r[i], int_to_double(pb)), vi);
            const int v = (int)(short)vi;
            const uint32_t c8 = (uint32_t)__vimin_s32_relu(v, 255);
            if ((i & 3) == 0) ow[i >> 2] = c8; else ow[i >> 2] |= c8 << (8 * (i & 3));
            if (DBG && o.idct_out) o.idct_out[y * BS + i] = r[i];
            // PFrame.py:39,63: float64 idct residual stored into an int8 plane (C cast: truncate, wrap)
            if (DBG && !intra_u8_resid && o.resid_mc) o.resid_mc[(size_t)y * o.resid_pitch + i] = (int8_t)(int)r[i];
        }
        store_row_words<BS>(o.recon + (size_t)y * o.rec_pitch, ow);
        // the intra wavefront predicts the next block of the row from this block's right column: hand it over in shared
        // memory instead of reading it back from the plane
        if (last_col) last_col[q * BS + y] = (uint8_t)(ow[BS / 4 - 1] >> 24);
        // ... and the block below predicts from this block's bottom row: post it in the row's mailbox, every pixel tagged
        // with the launch epoch, so that the consumer polls the data itself (no counter, no fence on either side)
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

// ---------------------------------------------------------------------------------------------
// The same transform spread over four warps (I-frame wavefront, tq.cu): the wavefront is a chain of bw + bh dependent
// blocks, each walked by a single warp -- ~9000 cycles per 16x16 block pair, two thirds of it the 512 DFMA + ~220 other
// fp64 operations every lane issues on ONE scheduler while the SM's other three idle.  Here a block pair belongs to a CTA
// of four warps (one per scheduler): lane -> (q, line) as before, but warp WQ computes only BS/4 of each line's outputs
// (forward: u or v in [WQ*BS/4, (WQ+1)*BS/4); inverse: the BS/8 mirrored pairs y, BS-1-y with y in [WQ*BS/8, ...)).
// Every output is the same fma chain as in fold_fwd / fold_inv (same operands, same order), so the results are
// bit-identical to tq_warp.  Lines are exchanged through two fp64 tiles (A: F1 -> F2, I1 -> I2; B: F2 -> I1) with one
// CTA barrier per pass, issued by the caller (all four warps meet at the same barrier instruction).
template <int BS>
struct QuadTile {
    static constexpr int NBW = 32 / BS;
    double A[NBW][BS][BS + 1];
    double B[NBW][BS][BS + 1];
    __align__(16) int16_t lev[NBW][BS][BS];
    __align__(16) int16_t res[NBW][BS][BS];
    __align__(16) uint8_t cur[NBW][BS][BS];
    __align__(16) uint8_t pred[NBW][BS][BS];
    __align__(16) uint8_t rec[NBW][BS][BS];
};

// F1: columns of the residual (dct.py:12), outputs u of warp WQ -> A[q][u][x]
template <int BS, int WQ>
__device__ __forceinline__ void quad_f1(QuadTile<BS>& t, int q, int x) {
    constexpr int Hh = BS / 2, UW = BS / 4;
    double s[Hh], d[Hh];
#pragma unroll
    for (int y = 0; y < Hh; y++) {
        const double lo = int_to_double((int)t.res[q][y][x]), hi = int_to_double((int)t.res[q][BS - 1 - y][x]);
        s[y] = __dadd_rn(lo, hi);
        d[y] = __dsub_rn(lo, hi);
    }
#pragma unroll
    for (int k = 0; k < UW; k++) {
        const int u = WQ * UW + k;
        double acc = 0.0;
#pragma unroll
        for (int x2 = 0; x2 < Hh; x2++) acc = __fma_rn(DctC<BS>::ct(u * BS + x2), (u & 1) ? d[x2] : s[x2], acc);
        t.A[q][u][x] = acc;
    }
}
// F2: rows (this lane owns row u = x), outputs v of warp WQ, quantise / rescale -> lev[q][u][v], B[q][u][v]
template <int BS, int WQ>
__device__ __forceinline__ void quad_f2(QuadTile<BS>& t, int q, int x, int qp) {
    constexpr int Hh = BS / 2, UW = BS / 4;
    const int u = x;
    double s[Hh], d[Hh];
#pragma unroll
    for (int i = 0; i < Hh; i++) {
        const double lo = t.A[q][u][i], hi = t.A[q][u][BS - 1 - i];
        s[i] = __dadd_rn(lo, hi);
        d[i] = __dsub_rn(lo, hi);
    }
    const bool su = (u == 0) || (2 * u == BS);
    const double w_sp = su ? DctC<BS>::w(0) : DctC<BS>::w(1);
    const double w_nm = su ? DctC<BS>::w(1) : DctC<BS>::w(2);
    short lv[UW];
#pragma unroll
    for (int k = 0; k < UW; k++) {
        const int v = WQ * UW + k;
        double acc = 0.0;
#pragma unroll
        for (int x2 = 0; x2 < Hh; x2++) acc = __fma_rn(DctC<BS>::ct(v * BS + x2), (v & 1) ? d[x2] : s[x2], acc);
        const bool sv = (v == 0) || (2 * v == BS);
        const double w = sv ? w_sp : w_nm;
        const double coef = __dmul_rn(acc, w);
        const int sh = qp + __vimin_s32_relu(u + v - (BS - 2), 2);
        int li;
        const double resc = quant_rescale(coef, 0x43380000 + (sh << 20), li);
        lv[k] = (short)li;
        t.B[q][u][v] = __dmul_rn(resc, w);
    }
    uint32_t* ls = reinterpret_cast<uint32_t*>(&t.lev[q][u][WQ * UW]);
#pragma unroll
    for (int k = 0; k < UW / 2; k++) ls[k] = (uint32_t)(uint16_t)lv[2 * k] | ((uint32_t)(uint16_t)lv[2 * k + 1] << 16);
}
// I1: over u for column v = x, mirrored output pairs of warp WQ -> A[q][y][x]
template <int BS, int WQ>
__device__ __forceinline__ void quad_i1(QuadTile<BS>& t, int q, int x) {
    constexpr int PW = BS / 8;
    double a[BS];
#pragma unroll
    for (int i = 0; i < BS; i++) a[i] = t.B[q][i][x];
#pragma unroll
    for (int p = 0; p < PW; p++) {
        const int y = WQ * PW + p;
        double e = 0.0, o = 0.0;
#pragma unroll
        for (int u = 0; u < BS; u += 2) e = __fma_rn(DctC<BS>::ct(u * BS + y), a[u], e);
#pragma unroll
        for (int u = 1; u < BS; u += 2) o = __fma_rn(DctC<BS>::ct(u * BS + y), a[u], o);
        t.A[q][y][x] = __dadd_rn(e, o);
        t.A[q][BS - 1 - y][x] = __dsub_rn(e, o);
    }
}
// I2: over v for row y = x, mirrored pixel pairs of warp WQ; reconstruct_block Frame.py:197-202 -> rec[q][y][i]
template <int BS, int WQ>
__device__ __forceinline__ void quad_i2(QuadTile<BS>& t, int q, int x) {
    constexpr int PW = BS / 8;
    const int y = x;
    double a[BS];
#pragma unroll
    for (int i = 0; i < BS; i++) a[i] = t.A[q][y][i];
#pragma unroll
    for (int p = 0; p < PW; p++) {
        const int i0 = WQ * PW + p;
        double e = 0.0, o = 0.0;
#pragma unroll
        for (int u = 0; u < BS; u += 2) e = __fma_rn(DctC<BS>::ct(u * BS + i0), a[u], e);
#pragma unroll
        for (int u = 1; u < BS; u += 2) o = __fma_rn(DctC<BS>::ct(u * BS + i0), a[u], o);
        const double r0 = __dadd_rn(e, o), r1 = __dsub_rn(e, o);
        int v0, v1;
        rint_magic(__dadd_rn(r0, int_to_double((int)t.pred[q][y][i0])), v0);
        rint_magic(__dadd_rn(r1, int_to_double((int)t.pred[q][y][BS - 1 - i0])), v1);
        t.rec[q][y][i0] = (uint8_t)__vimin_s32_relu((int)(short)v0, 255);
        t.rec[q][y][BS - 1 - i0] = (uint8_t)__vimin_s32_relu((int)(short)v1, 255);
    }
}
#define BVC_QUAD_DISPATCH(fn, warp, ...)                     \
    switch (warp) {                                          \
        case 0: fn<BS, 0>(__VA_ARGS__); break;               \
        case 1: fn<BS, 1>(__VA_ARGS__); break;               \
        case 2: fn<BS, 2>(__VA_ARGS__); break;               \
        default: fn<BS, 3>(__VA_ARGS__); break;              \
    }

// ---------------------------------------------------------------------------------------------
// zig-zag position table: zz[p] = row*BS + col (entropy_encoder.py:115-135):
// even diagonals run (row=i, col=s-i) with i ascending, odd ones (row=s-i, col=i).
template <int BS>
__device__ __forceinline__ void build_zigzag(uint8_t* zz, int tid, int nthreads) {
    for (int e = tid; e < BS * BS; e += nthreads) {
        const int r = e / BS, c = e % BS, s = r + c;
        int before = (s < BS) ? s * (s + 1) / 2 : BS * BS - (2 * BS - 1 - s) * (2 * BS - s) / 2;
        const int i0 = max(0, s - BS + 1);
        const int idx = (s & 1) ? (c - i0) : (r - i0);
        zz[before + idx] = (uint8_t)e;
    }
}

// OR `len` bits of `code` (right aligned) into the big-endian word stream `buf` at bit offset `off`.
__device__ __forceinline__ void put_bits_smem(uint32_t* buf, int off, unsigned long long code, int len) {
    const unsigned long long V = code << (64 - len);
    const int sh = off & 31, wi = off >> 5;
    const uint32_t hi = (uint32_t)(V >> 32), lo = (uint32_t)V;
    const uint32_t w0 = hi >> sh;
    const uint32_t w1 = sh ? ((hi << (32 - sh)) | (lo >> sh)) : lo;
    const uint32_t w2 = sh ? (lo << (32 - sh)) : 0u;
    if (w0) atomicOr(&buf[wi], w0);
    if (w1) atomicOr(&buf[wi + 1], w1);
    if (w2) atomicOr(&buf[wi + 2], w2);
}

// Entropy-code one block (levels in smem tile `lev`, BS x BS) cooperatively by one warp.
// Returns the number of bits; the bits are left in `bits` (big-endian words) and copied to `gout`.
//
// M[i] = ballot of the non-zero levels at zig-zag positions 32*i + lane; a run starts where the non-zero state flips
// (position 0 always starts one) and ends where the next one starts.  "Events" are the positions that emit symbols
// (run starts and non-zero values).  A block with at most 32 events (the usual case after quantisation) is coded in
// one pass with lane j on event j -- the j-th set bit of the 256-bit event mask -- so the work is balanced whatever
// the positions are; denser blocks take the position-parallel loop, 32 positions per iteration.
template <int BS>
__device__ __forceinline__ int entropy_block_warp(const int16_t* lev, const uint8_t* zz, uint32_t* bits, int lane,
                                                  uint32_t* gout) {
    constexpr int N = BS * BS;
    constexpr int NI = N >= 32 ? N / 32 : 1;   // iterations
    constexpr int AL = N >= 32 ? 32 : N;       // positions (= active lanes) per iteration
    constexpr uint32_t ALMASK = AL == 32 ? 0xffffffffu : ((1u << AL) - 1u);
    const bool act = lane < AL;
    int c[NI];
    uint32_t M[NI];
    int nnz = 0;
#pragma unroll
    for (int i = 0; i < NI; i++) {
        c[i] = act ? (int)lev[zz[i * AL + lane]] : 0;
        M[i] = __ballot_sync(0xffffffffu, c[i] != 0);
        nnz += __popc(M[i]);
    }
    // zero what the block can need: <= 31 bits per value, <= 19 per run header, <= 2*nnz+1 runs, 27 for the end marker
    {
        constexpr int WCAP = blk_words_for<BS>() + 4;
        const int wmax = min(WCAP, ((69 * nnz + 46 + 31) >> 5) + 2);
        for (int w = lane; w < wmax; w += 32) bits[w] = 0;
    }
    __syncwarp();
    int base = 0;
    // events = positions that emit symbols (run starts and non-zero values), in scan order
    uint32_t S[NI], EV[NI];
    int P[NI];
    int E = 0;
#pragma unroll
    for (int i = 0; i < NI; i++) {
        const uint32_t carry = (i == 0) ? ((~M[0]) & 1u) : (M[i > 0 ? i - 1 : 0] >> (AL - 1)) & 1u;
        S[i] = (M[i] ^ ((M[i] << 1) | carry)) & ALMASK;
        EV[i] = S[i] | M[i];
        P[i] = E;
        E += __popc(EV[i]);
    }
    if (E <= 32) {
        // fast path (sparse block): lane j codes event j -- one pass, balanced over the lanes
        const bool on = lane < E;
        int wi = 0, pb = 0;
        uint32_t evw = EV[0], mw = M[0], sw = S[0];
#pragma unroll
        for (int i = 1; i < NI; i++)
            if (lane >= P[i]) { wi = i; pb = P[i]; evw = EV[i]; mw = M[i]; sw = S[i]; }
        // position of the (lane - pb)-th set bit of evw
        int n = lane - pb, bit = 0;
        {
            uint32_t w = evw;
            int k = __popc(w & 0xffffu);
            if (n >= k) { n -= k; bit += 16; w >>= 16; }
            k = __popc(w & 0xffu);
            if (n >= k) { n -= k; bit += 8; w >>= 8; }
            k = __popc(w & 0xfu);
            if (n >= k) { n -= k; bit += 4; w >>= 4; }
            k = __popc(w & 0x3u);
            if (n >= k) { n -= k; bit += 2; w >>= 2; }
            if (n >= (int)(w & 1u)) bit += 1;
        }
        const int p = on ? wi * AL + bit : 0;
        const bool isnz = on && ((mw >> bit) & 1u);
        const bool st = on && ((sw >> bit) & 1u);
        // a run ends where the next one starts: the next start among the events (they are in scan order)
        const uint32_t sb = __ballot_sync(0xffffffffu, st) & ~((2u << lane) - 1u);
        const int pn = __shfl_sync(0xffffffffu, p, sb ? __ffs(sb) - 1 : 0);
        const int next = sb ? pn : N;
        unsigned long long cd = 0;
        int ln = 0;
        if (st) {
            const int runlen = next - p;
            const uint32_t e = eg_code(isnz ? -runlen : (next == N ? 0 : runlen));
            ln = eg_len_of_code(e);
            cd = e;
        }
        if (isnz) {
            const uint32_t e = eg_code((int)lev[zz[p]]);
            const int l2 = eg_len_of_code(e);
            cd = (cd << l2) | e;
            ln += l2;
        }
        int incl = ln;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const int o = __shfl_up_sync(0xffffffffu, incl, d);
            if (lane >= d) incl += o;
        }
        if (ln) put_bits_smem(bits, incl - ln, cd, ln);
        base = __shfl_sync(0xffffffffu, incl, 31);
    } else {
        // dense block: position-parallel, 32 positions per iteration
#pragma unroll 1
        for (int i = 0; i < NI; i++) {
            uint32_t Mi = 0, Si = 0;
            int ci = 0;
#pragma unroll
            for (int k = 0; k < NI; k++)
                if (k == i) { Mi = M[k]; Si = S[k]; ci = c[k]; }
            if ((Si | Mi) == 0) continue;
            const bool isnz = (Mi >> lane) & 1u;
            const bool st = act && ((Si >> lane) & 1u);
            const int p = i * AL + lane;
            unsigned long long cd = 0;
            int ln = 0;
            if (st) {
                const uint32_t inv = isnz ? 0xffffffffu : 0u;
                const uint32_t w = ((Mi ^ inv) & ALMASK) & ~((2u << lane) - 1u);
                int next = w ? i * AL + __ffs(w) - 1 : -1;
#pragma unroll
                for (int j = 1; j < NI; j++) {
                    const uint32_t x = (M[j] ^ inv) & ALMASK;
                    if (j > i && next < 0 && x) next = j * AL + __ffs(x) - 1;
                }
                if (next < 0) next = N;
                const int runlen = next - p;
                const uint32_t e = eg_code(isnz ? -runlen : (next == N ? 0 : runlen));
                ln = eg_len_of_code(e);
                cd = e;
            }
            if (isnz) {
                const uint32_t e = eg_code(ci);
                const int l2 = eg_len_of_code(e);
                cd = (cd << l2) | e;
                ln += l2;
            }
            int incl = ln;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const int o = __shfl_up_sync(0xffffffffu, incl, d);
                if (lane >= d) incl += o;
            }
            if (ln) put_bits_smem(bits, base + incl - ln, cd, ln);
            base += __shfl_sync(0xffffffffu, incl, 31);
        }
    }
    const int nbits = base + 27;  // + EG(8190) end-of-block marker (Frame.py:75)
    const int nwords = (nbits + 31) >> 5;
    if (lane == 0) put_bits_smem(bits, base, (unsigned long long)eg_code(BVC_EOB_MARKER), 27);
    __syncwarp();
    for (int w = lane; w < nwords; w += 32) gout[w] = bits[w];
    __syncwarp();
    return nbits;
}

// ---------------------------------------------------------------------------------------------
// Entropy coding of ALL blocks of a warp tile in one go (P-frame path).  entropy_block_warp above walks one block at a
// time: its dependent chain (ballots -> event selection -> scan -> shared-memory atomics -> copy) is paid per block and
// the end-of-block marker is a divergent extra.  Here the emitting positions ("events": run starts and non-zero values,
// scan order) of every block of the tile are first compacted into ONE list -- each block closed by a terminator event
// that carries the end-of-block marker and, being a run start at position N, also ends the block's last run -- and then
// lane j codes event j: a tile of sparse blocks (two 16x16 blocks with <= 30 events in all, the usual case after
// quantisation) takes a single pass, the bit offsets restart at every terminator (segmented scan), and the copy to the
// per-block bit strings runs for all blocks at once (lane group q = block q).
// The scratch lives in the tile's fp64 exchange buffer, which is idle while the levels are coded.
constexpr uint32_t EV_NZ = 1u << 12, EV_ST = 1u << 13, EV_EOB = 1u << 14;   // bits 0-8 position, 9-11 block, 16-31 level
template <int BS>
struct EntScratch {
    static constexpr int NBW = 32 / BS, N = BS * BS, WCAP = blk_words_for<BS>() + 4;
    uint32_t bits[NBW][WCAP];
    uint32_t ev[NBW * (N + 1)];
    int nbits[NBW];
};

// Phase A: event list of the tile's valid blocks (vmask = ballot of the lanes' `valid`), bit strings zeroed.
// Returns the number of events (warp-uniform).
template <int BS>
__device__ __forceinline__ int entropy_tile_events(const WarpTile<BS>& t, EntScratch<BS>& es, const uint8_t* zz, int lane, uint32_t vmask) {
    constexpr int NBW = 32 / BS, N = BS * BS;
    constexpr int NI = N >= 32 ? N / 32 : 1;
    constexpr int AL = N >= 32 ? 32 : N;
    constexpr uint32_t ALMASK = AL == 32 ? 0xffffffffu : ((1u << AL) - 1u);
    const bool act = lane < AL;
    const uint32_t lt = (1u << lane) - 1u;
    int zp[NI];   // this lane's zig-zag positions: the same for every block and every tile
#pragma unroll
    for (int i = 0; i < NI; i++) zp[i] = act ? (int)zz[i * AL + lane] : 0;
    int E = 0;
#pragma unroll 1
    for (int q = 0; q < NBW; q++) {
        if (!((vmask >> (q * BS)) & 1u)) continue;
        const int16_t* lev = &t.lev[q][0][0];
        uint32_t top = 0;
        int nnz = 0;
#pragma unroll
        for (int i = 0; i < NI; i++) {
            const int c = act ? (int)lev[zp[i]] : 0;
            const uint32_t M = __ballot_sync(0xffffffffu, c != 0);
            // a run starts where the non-zero state flips; position 0 always starts one
            const uint32_t carry = (i == 0) ? ((~M) & 1u) : top;
            const uint32_t S = (M ^ ((M << 1) | carry)) & ALMASK;
            const uint32_t EV = S | M;
            if ((EV >> lane) & 1u)
                es.ev[E + __popc(EV & lt)] = ((uint32_t)c << 16) | (uint32_t)(i * AL + lane) | ((uint32_t)q << 9) |
                                             (((M >> lane) & 1u) ? EV_NZ : 0u) | (((S >> lane) & 1u) ? EV_ST : 0u);
            E += __popc(EV);
            nnz += __popc(M);
            top = (M >> (AL - 1)) & 1u;
        }
        if (lane == 0) es.ev[E] = (uint32_t)N | ((uint32_t)q << 9) | EV_ST | EV_EOB;
        E++;
        // zero what the block can need: <= 31 bits per value, <= 19 per run header, <= 2*nnz+1 runs, 27 for the end marker
        const int wmax = min(EntScratch<BS>::WCAP, ((69 * nnz + 46 + 31) >> 5) + 2);
        for (int w = lane; w < wmax; w += 32) es.bits[q][w] = 0;
    }
    __syncwarp();
    return E;
}

// Phase B: lane j codes event j (32 events per pass); Frame.py:61-75, entropy_encoder.py:8-29,65-88.
template <int BS>
__device__ __forceinline__ void entropy_tile_code(EntScratch<BS>& es, int E, int lane) {
    constexpr int N = BS * BS;
    const uint32_t lt = (1u << lane) - 1u;
    int carry_bits = 0;   // bits already written for the block that is open at the start of the pass
#pragma unroll 1
    for (int base = 0; base < E; base += 32) {
        const uint32_t w = (base + lane < E) ? es.ev[base + lane] : 0u;
        const int p = (int)(w & 511u), q = (int)((w >> 9) & 7u);
        const bool isnz = w & EV_NZ, st = w & EV_ST, eob = w & EV_EOB;
        // a run ends where the next one starts (the terminator is a start at position N)
        const uint32_t later = __ballot_sync(0xffffffffu, st) & ~((2u << lane) - 1u);
        int pn = __shfl_sync(0xffffffffu, p, later ? __ffs(later) - 1 : 0);
        if (st && !eob && !later) {   // dense tiles only: the next start lies in a later pass
            int k = base + 32;
            while (!(es.ev[k] & EV_ST)) k++;
            pn = (int)(es.ev[k] & 511u);
        }
        unsigned long long cd = 0;
        int ln = 0;
        if (eob) {
            cd = eg_code(BVC_EOB_MARKER);   // Frame.py:75
            ln = 27;
        } else if (st) {
            const int runlen = pn - p;
            const uint32_t e = eg_code(isnz ? -runlen : (pn == N ? 0 : runlen));
            ln = eg_len_of_code(e);
            cd = e;
        }
        if (isnz) {
            const uint32_t e = eg_code((int)w >> 16);
            const int l2 = eg_len_of_code(e);
            cd = (cd << l2) | e;
            ln += l2;
        }
        int incl = ln;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const int o = __shfl_up_sync(0xffffffffu, incl, d);
            if (lane >= d) incl += o;
        }
        // bit offsets restart after every terminator
        const uint32_t eobb = __ballot_sync(0xffffffffu, eob);
        const uint32_t prior = eobb & lt;
        const int corr = __shfl_sync(0xffffffffu, incl, prior ? 31 - __clz(prior) : 0);
        const int off = prior ? (incl - ln - corr) : (carry_bits + incl - ln);
        if (ln) put_bits_smem(es.bits[q], off, cd, ln);
        if (eob) es.nbits[q] = off + 27;
        const int tot = __shfl_sync(0xffffffffu, incl, 31);
        const int last = __shfl_sync(0xffffffffu, incl, eobb ? 31 - __clz(eobb) : 0);
        carry_bits = eobb ? tot - last : carry_bits + tot;
    }
    __syncwarp();
}

// Phase C: every lane group copies its block's bit string (lane = (q, x)); `gout` / `nbits_out` belong to the lane's block.
template <int BS>
__device__ __forceinline__ void entropy_tile_store(const EntScratch<BS>& es, int lane, bool valid, uint32_t* gout, int32_t* nbits_out) {
    const int q = lane / BS, x = lane % BS;
    const int nb = valid ? es.nbits[q] : 0;
    const int nwords = (nb + 31) >> 5;
    for (int w = x; w < nwords; w += BS) gout[w] = es.bits[q][w];
    if (valid && x == 0) *nbits_out = nb;
    __syncwarp();
}

// ---------------------------------------------------------------------------------------------
// P-frame path, one warp task = NBW blocks (lane -> block `b` of lane group `fl`; lanes of one block hold the same values;
// !valid lanes name any in-range block and produce nothing): PFrame.process_block PFrame.py:99-125,230-249;
// Frame.py:61-75,190-202.  The task's pixel rows are requested one task ahead, behind the entropy coding of the task
// before (pframe_request_rows / pframe_stage_rows), so the two dependent global round trips (motion vector -> predicted
// row) are off the critical path.
// Division by a launch constant (x < 2^31): q = (mulhi(x, magic) + x) >> shift.
__device__ __forceinline__ uint32_t fast_div(uint32_t x, uint32_t magic, uint32_t shift) { return (__umulhi(x, magic) + x) >> shift; }

struct PTask {
    int fl, b, oy, ox;   // lane group, block, block origin
    bool valid;
};

// Asynchronous global -> shared copies (LDGSTS): the rows of the NEXT task travel into tile regions that are idle while
// the current task's levels are coded, without occupying registers or a scoreboard.
template <int BS>
__device__ __forceinline__ void cp_async_row(void* dst, const void* src) {   // BS bytes, BS-byte aligned
    const uint32_t d = (uint32_t)__cvta_generic_to_shared(dst);
    if constexpr (BS == 16) asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"(d), "l"(src) : "memory");
    else if constexpr (BS == 8) asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(d), "l"(src) : "memory");
    else asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(d), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }
// motion vector without its fourth component (the SAD): a destination register nobody reads would be reused at once and
// the write-after-write hazard would stall the warp until the load lands
__device__ __forceinline__ int4 ldg_mv_keep(const int4* p) {
    int4 v;
    asm volatile("ld.global.v2.s32 {%0, %1}, [%2];" : "=r"(v.x), "=r"(v.y) : "l"(p));
    asm volatile("ld.global.s32 %0, [%1+8];" : "=r"(v.z) : "l"(p));
    v.w = 0;
    return v;
}
// one ticket of the task counter; asm so that the compiler neither aggregates it across the warp nor waits for it early
__device__ __forceinline__ int ticket_draw(int* counter) {
    int v;
    asm volatile("atom.global.add.s32 %0, [%1], 1;" : "=r"(v) : "l"(counter) : "memory");
    return v;
}

// Request the pixel rows of task k (lane x = row x of its block): the current row goes straight to t.cur[q][x], the aligned
// words covering the predicted row (find_mv_predicted_block PFrame.py:230-244: refs[mv[2]] only when more than one
// reference is present) to the lane's own slot in t.res, where pframe_stage_rows picks them up.  Returns the predicted
// row's byte offset inside the slot.
template <int BS>
__device__ __forceinline__ uint32_t pframe_request_rows(const TqArgs& a, WarpTile<BS>& t, const PTask& k, int q, int x, int4 mv) {
    const FrameLane& L = a.lanes[k.fl];
    const uint8_t* cur = a.cur_base + (size_t)L.cur_plane * a.cur_plane_bytes + (size_t)(k.oy + x) * a.cur_pitch + k.ox;
    const int kref = (L.nref > 1) ? mv.z : 0;
    int plane = L.ref_plane[kref];
    int dx = mv.x, dy = mv.y;
    if (a.frac) {  // half-pel MV = integer offset on one of the four phase planes
        plane += (mv.x & 1) | ((mv.y & 1) << 1);
        dx = mv.x >> 1;
        dy = mv.y >> 1;
    }
    const uint8_t* pr = a.ref_base + (size_t)plane * a.ref_plane_bytes + (size_t)(k.oy + dy + x) * a.ref_pitch + (k.ox + dx);
    // the BS predicted bytes lie in at most two aligned BS-byte chunks of the reference row: two wide copies into the lane's
    // 2*BS-byte slot instead of BS/4+1 narrow ones (the second chunk is skipped when the row is aligned: it could lie
    // beyond the pool)
    const uintptr_t ad = reinterpret_cast<uintptr_t>(pr);
    const uint8_t* base = reinterpret_cast<const uint8_t*>(ad & ~(uintptr_t)(BS - 1));
    cp_async_row<BS>(&t.cur[q][x][0], cur);
    uint8_t* slot = reinterpret_cast<uint8_t*>(&t.res[q][x][0]);
    cp_async_row<BS>(slot, base);
    if (ad & (BS - 1)) cp_async_row<BS>(slot + BS, base + BS);
    return (uint32_t)(ad & (BS - 1));
}
// ... and turn them into the staged tile rows (cur, pred, residual) once they have arrived
template <int BS>
__device__ __forceinline__ void pframe_stage_rows(WarpTile<BS>& t, int q, int x, uint32_t off) {   // off = byte offset of the row in its slot
    cp_async_wait_all();
    uint32_t cw[BS / 4], raw[BS / 4 + 1], pw[BS / 4];
    load_row_aligned<BS>(&t.cur[q][x][0], cw);
    const uint32_t* rp = reinterpret_cast<const uint32_t*>(&t.res[q][x][0]) + (off >> 2);
    const uint32_t sh = (off & 3u) * 8u;
#pragma unroll
    for (int i = 0; i < BS / 4; i++) raw[i] = rp[i];
    raw[BS / 4] = sh ? rp[BS / 4] : 0u;   // an aligned row ends with its last word (the slot may end there too)
#pragma unroll
    for (int i = 0; i < BS / 4; i++) pw[i] = __funnelshift_r(raw[i], raw[i + 1], sh);
    stage_row<BS>(t, q, x, cw, pw);
}

}  // namespace bvc
