// me_fullsearch.cu -- K1/K3: exhaustive block-matching motion estimation (integer-pel and half-pel),
// multi-reference, with the reference's exact tie-break.
//
// Replaces find_lowest_mae_block + get_ref_block_at_mv + is_out_of_range + common.mae
// (reference encoder/block_predictor.py:61-143, common.py:43-45).
//
// Winner rule (block_predictor.py:76-88): minimum SAD; among equal SAD the smaller |mvx|+|mvy|;
// among those the first in scan order ref ascending, mv_y ascending, mv_x ascending.  Candidates whose
// block leaves the plane are skipped (no clamping / padding).  With fracMeEnabled the search runs
// over +-2r half-pel positions of the 2x interpolated plane, sampled with stride 2
// (block_predictor.py:65-66,103-111): we keep the 2x plane as four W x H *phase planes*
// (even/odd column x even/odd row), so a half-pel candidate (mvx,mvy) is the integer candidate
// (mvx>>1, mvy>>1) on phase plane (mvx&1, mvy&1) and the same kernel serves both modes.
//
// Kernel design (sm_100a, HBM is irrelevant here: ~2000 byte-ops per byte read):
//   * bound by the ALU pipe: VABSDIFF4.U8.ACC issues at 16 lanes/clk/SMSP (measured,
//     profiles/microbench) => 256 pixel-absdiffs/clk/SM;  PRMT/SHF share that pipe, so byte
//     re-alignment must not be in the inner loop.
//   * one CTA = NB horizontally adjacent blocks.  The search window (NB*BS+2R) x (BS+2R) is staged
//     in shared memory by one TMA tile load (out-of-frame bytes are zero-filled by the TMA unit; the
//     box origin must be 16-byte aligned -- a byte-granular origin raises an illegal-instruction fault on
//     sm_100a, profiles/microbench/tma_probe.cu -- so the box starts at the aligned column at or left of
//     x0-R).  The CTA then derives three byte-shifted copies (+1,+2,+3) with funnel shifts, once per
//     window, so every candidate column reads naturally aligned 32-bit words and the inner loop has
//     no PRMT/SHF at all (they would steal VABSDIFF4 issue slots).
//   * one thread = one candidate column mx of one block; it slides down all 2R+1 vertical offsets.
//     The current block lives in registers (BS*BS/4 words); each window row is loaded once (BS/4
//     LDS.32) and feeds BS candidates (BS*BS/4 VABSDIFF4) => LDS:VABSDIFF4 = 1:BS.  BS accumulators
//     rotate through compile-time slots; the row loop is unrolled BS deep in three variants (ramp-up,
//     steady, ramp-down) so no candidate outside [-R,R] is ever evaluated: executed VABSDIFF4 count
//     equals the algorithmic count exactly.
//   * argmin: per thread one IMAD (FMA pipe) + one VIMNMX per candidate on a packed key
//     (SAD | L1 | m); per pass merged with the full 64-bit key (SAD, L1, ref, mvy, mvx); per block a
//     shared-memory 64-bit atomicMin.
//   * the last candidate column (dx = +R) of every block would leave a warp 1/32 full: it is split into vertical segments,
//     and the segments are jobs of 32 that the CTA's warps draw from a counter once their own columns are done, so lane
//     utilisation stays ~98 % without extra warps (256 threads per CTA leave 128 registers for two CTAs per SM).
//   * window words are requested a whole word ahead of their use (software-pipelined bodies); the prologue of a CTA keeps
//     off the ALU pipe, because a CTA that starts beside a searching CTA only gets that pipe's gaps.
#include "bvc_common.cuh"
#include "bvc_kernels.h"

namespace bvc {

namespace {

template <int BS>
struct CurBlock {
    static constexpr int WPR = BS / 4;  // words per row
    uint32_t w[BS][WPR];
};

template <int BS>
__device__ __forceinline__ void load_cur(CurBlock<BS>& c, const uint8_t* p, int pitch) {
#pragma unroll
    for (int r = 0; r < BS; r++) {
        const uint8_t* row = p + (size_t)r * pitch;
        if constexpr (BS == 16) {
            uint4 v = *reinterpret_cast<const uint4*>(row);
            c.w[r][0] = v.x; c.w[r][1] = v.y; c.w[r][2] = v.z; c.w[r][3] = v.w;
        } else if constexpr (BS == 8) {
            uint2 v = *reinterpret_cast<const uint2*>(row);
            c.w[r][0] = v.x; c.w[r][1] = v.y;
        } else {
            c.w[r][0] = *reinterpret_cast<const uint32_t*>(row);
        }
    }
}

enum { BODY_FIRST = 0, BODY_MID = 1, BODY_LAST = 2 };

// A CTA that starts while its SM's other CTA is in the bodies gets the ALU pipe only when that one leaves it idle, so every
// ALU-pipe instruction of the prologue (shifts, compares, the integer-division sequences) stretches the time until this CTA's
// warps reach their own bodies and the SM has four searching warps per scheduler again.  The prologue therefore does its
// arithmetic on the FMA pipe where it can: divisions by launch constants as multiply-high + shift, the byte-shifted window
// copies as multiply-high + multiply-add.
__device__ __forceinline__ uint32_t div_magic(uint32_t x, uint32_t magic, uint32_t shift) { return (__umulhi(x, magic) + x) >> shift; }
static void div_magic_constants(uint32_t d, uint32_t& magic, uint32_t& shift) {   // exact for x < 2^31
    shift = 0;
    while ((1u << shift) < d) shift++;
    magic = (uint32_t)((((unsigned long long)1 << 32) * (((unsigned long long)1 << shift) - d)) / d + 1);
}
// (lo >> s) | (hi << (32 - s)) for s = 8, 16, 24 with k = 1 << (32 - s) in a register: mul.hi + mad.lo, both on the FMA pipe
__device__ __forceinline__ uint32_t funnel_fma(uint32_t lo, uint32_t hi, uint32_t k) {
    uint32_t t, r;
    asm("mul.hi.u32 %0, %1, %2;" : "=r"(t) : "r"(lo), "r"(k));
    asm("mad.lo.u32 %0, %1, %2, %3;" : "=r"(r) : "r"(hi), "r"(k), "r"(t));
    return r;
}

// Per-pass constants of the per-thread argmin.
//   PACKED : key = SAD << (l1bits+mbits) | L1 << mbits | m   -- one IMAD + one VIMNMX per candidate;
//            usable when the three fields fit 32 bits (always for the headline configurations).
//   general: key = SAD << 9 | L1, m tracked separately (strict-less keeps the first m).
struct KeyCfg {
    uint32_t scale;   // PACKED: 1 << (l1bits + mbits)
    int mbits;        // PACKED: bits of the m field
};

__device__ __forceinline__ uint32_t imad_u32(uint32_t a, uint32_t b, uint32_t c) {
    uint32_t d;
    asm("mad.lo.u32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));  // IMAD: FMA pipe, not the ALU pipe
    return d;
}

// One unrolled body of BS window rows.  `rowp` points at this thread's first word of the body's first
// window row; `wpitch` is the window pitch in words.  The candidate completing after row t has
// m = mbase + t (m = vertical offset + R).
//   PACKED : `utab[m]` (shared memory, rebuilt per pass) holds (|mvy(m)| << mbits) | m, with bit 31 set
//            when the vertical offset leaves the plane, so a finished candidate costs one LDS, two IMAD
//            (FMA pipe) and ONE ALU-pipe instruction (VIMNMX); no compare, select or branch.
//   general: per-thread arithmetic (used only when SAD | L1 | m does not fit 31 bits).
//   SADMAP : additionally store the SAD of every in-range candidate to smap[m * n1] (FastME's look-up table).
template <int BS, int MODE, bool PACKED, bool SADMAP, bool PIPE>
__device__ __forceinline__ void me_body(const CurBlock<BS>& cur, uint32_t (&acc)[BS], const uint32_t* rowp, int wpitch,
                                        int mbase, const uint32_t* utab_m, int mlo, int mhi, int mvy0, int sc, uint32_t tthr,
                                        uint32_t one, const KeyCfg kc, uint32_t& best, uint32_t& bestm, uint16_t* smap, int n1) {
    constexpr int WPR = BS / 4;
    uint32_t pend = 0xFFFFFFFFu;
    // Window words are requested one word ahead of their use (PIPE, constant pitch only): volatile loads and volatile
    // VABSDIFF4s keep the interleaving written here, so a warp meets the shared-memory latency once per body instead of
    // once per word (left to itself ptxas sinks every LDS to just in front of its first use).
    const volatile uint32_t* vrow = rowp;
    uint32_t wcur = 0;
    if (PIPE) wcur = vrow[0];
#pragma unroll
    for (int t = 0; t < BS; t++) {
        uint32_t w[WPR];
        if (!PIPE) {
#pragma unroll
            for (int j = 0; j < WPR; j++) w[j] = rowp[t * wpitch + j];
        }
#pragma unroll
        for (int wi = 0; wi < WPR; wi++) {
            uint32_t wnext = 0;
            if (PIPE) {
                w[wi] = wcur;
                if (!(t == BS - 1 && wi == WPR - 1)) wnext = vrow[(wi + 1 < WPR) ? (t * wpitch + wi + 1) : ((t + 1) * wpitch)];
            }
#pragma unroll
            for (int j = 0; j < BS; j++) {
                // candidate offset m = y - j; FIRST body: m >= 0 <=> j <= t; LAST body: m <= 2R <=> j >= t
                bool on = true;
                if (MODE == BODY_FIRST) on = (j <= t);
                if (MODE == BODY_LAST) on = (j >= t);
                if (on) {
                    const int slot = (t - j) & (BS - 1);
                    // first word of a fresh candidate (j == 0, wi == 0) starts from zero: no reset needed
                    const uint32_t c0 = (j == 0 && wi == 0) ? 0u : acc[slot];
                    acc[slot] = PIPE ? sad4_keep(w[wi], cur.w[j][wi], c0) : sad4(w[wi], cur.w[j][wi], c0);
                }
            }
            if (PIPE) wcur = wnext;
        }
        // the candidate whose last row (j = BS-1) was just added is complete
        const bool completes = (MODE != BODY_FIRST) || (t == BS - 1);
        if (completes) {
            const int slot = (t + 1) & (BS - 1);
            if (PACKED) {
                // the thread's own |mvx| term is the same for all of its candidates: it joins the winner after the bodies
                const uint32_t u = utab_m[t];   // utab_m = utab + mbase (per thread)
                const uint32_t key = imad_u32(acc[slot], kc.scale, u);
                // two finished candidates per ALU-pipe instruction (VIMNMX3): the steady and ramp-down bodies finish one
                // candidate per row, so even rows park their key and odd rows fold both into the running minimum
                if (MODE == BODY_FIRST) best = min(best, key);
                else if ((t & 1) == 0) pend = key;
                else best = __vimin3_u32(best, pend, key);
                if (SADMAP && !(u >> 31)) smap[(mbase + t) * n1] = (uint16_t)acc[slot];
            } else {
                const int m = mbase + t;
                const uint32_t amvy = (uint32_t)abs(mvy0 + sc * m);
                const uint32_t key = imad_u32(acc[slot], 512u, imad_u32(one, amvy, tthr));
                if (m >= mlo && m <= mhi && key < best) { best = key; bestm = (uint32_t)m; }
            }
        }
    }
}

// Tiled full-search kernel.  One CTA = NB x NBY blocks (NB side by side, NBY stacked); grid = tiles x lanes, linear.
// Dynamic smem: the TMA window (NBY*BS+2R rows) + 3 shifted copies + the CTA's current blocks.
// Threads: NB*2R, one per candidate column dx in [-R, R-1] of one block column; each slides over all 2R+1 vertical
// offsets once per stacked block row (current block reloaded into registers from shared memory).  The last column
// dx = +R of every block is split into vertical segments of BS+1 candidates (ramp-up + ramp-down body only); 32 segments
// are one "extra job", and a warp that has finished its own columns draws extra jobs from a shared-memory counter until
// none is left -- no warp waits at the closing barrier while work remains, and the CTA needs no extra warps.
// WPC: window pitch in words when it is known at compile time (the headline shapes), 0 = read it from the arguments.
// With a constant pitch every LDS of an unrolled body takes an immediate offset from one base register: no address
// arithmetic (VIADD, ALU pipe) between the VABSDIFF4s.
// Registers: 96 let two CTAs of up to 341 threads share an SM.  The two constant-pitch shapes run 256 threads, so two CTAs
// fit at 128 registers: the spare ones are what lets ptxas keep the window loads a whole word ahead of their use (me_body,
// PIPE); at 96 it sinks a quarter of them back to just in front of their first use.
template <int BS, int NB, int NBY, bool PACKED, bool SADMAP, int WPC>
__global__ void __maxnreg__(WPC ? 128 : 96) me_tiled_kernel(const __grid_constant__ CUtensorMap ref_map, MeArgs a) {
    extern __shared__ __align__(128) uint8_t smem[];
    __shared__ uint64_t bar;
    __shared__ unsigned long long sbest[NBY][NB];
    __shared__ int xjob;   // extra jobs handed out so far (all windows of the CTA)
    // per-pass candidate tables, indexed by m + BS (the constant-pitch shapes have Rv <= 64: with eight stacked rows two CTAs
    // only fit an SM with the smaller table)
    constexpr int UTN = (WPC ? 2 * 64 : 2 * 128) + 1 + 2 * BS;
    __shared__ uint32_t utab[NBY][UTN];

    const int tid = threadIdx.x;
    const int R = a.R;        // horizontal range = the search range (plane units)
    const int Rv = a.Rv;      // vertical range walked by the bodies: R rounded up so that 2*Rv is a multiple of BS;
                              // offsets beyond R are masked through utab (bit 31) like offsets outside the plane
    const int WW = a.win_pitch;               // bytes, multiple of 16
    const int copy_stride = a.win_copy_bytes + 32;  // +32 B: copy k starts 8 banks after copy k-1 (conflict-free LDS)
    // Tile from the linear CTA index.  CTAs [0, n_full) own a whole tile of NB x NBY blocks; the tiles of the last,
    // partly filled wave are cut into NBY CTAs of one block row each (a quarter of the work at NBY = 4), so the tail of
    // the launch is a fraction of a tile time instead of a whole one.
    int tile = blockIdx.x, yy0 = 0, yy1 = NBY;
    if (tile >= a.n_full) {
        const int u = tile - a.n_full;
        tile = a.n_full + u / NBY;
        yy0 = u % NBY;
        yy1 = yy0 + 1;
    }
    const int per_z = a.tiles_x * a.tiles_y;
    const int z = (int)div_magic((uint32_t)tile, a.perz_magic, a.perz_shift), t2 = tile - z * per_z;
    const int ty = (int)div_magic((uint32_t)t2, a.tx_magic, a.tx_shift), tx = t2 - ty * a.tiles_x;
    const int bx0 = tx * NB;
    const int by0 = ty * NBY;
    const int rows = (yy1 - yy0) * BS + 2 * Rv;       // window rows this CTA reads (the TMA box always has NBY*BS + 2*Rv)
    const int box_bytes = WW * (NBY * BS + 2 * Rv);
    const int wy0 = (by0 + yy0) * BS - Rv;           // plane row of window row 0
    // SAD-map mode has no reduction across references, so every (lane, reference) pair gets its own CTAs
    // (z = lanes * max_refs): small frames (CIF: 30 tiles per lane) then fill the GPU
    const int lane = SADMAP ? z / a.max_refs : z;
    const MeLane& L = a.lanes[lane];
    const int r_begin = SADMAP ? z % a.max_refs : 0;
    const int r_end = SADMAP ? min(r_begin + 1, L.nref) : L.nref;
    if (r_begin >= r_end) return;
    if (by0 + yy0 >= a.bh) return;                   // sub-tile below the last block row
    uint8_t* scur = smem + 4 * (size_t)copy_stride;   // [NBY*BS][NB*BS] current pixels of the CTA's blocks

    const int nmain = NB * 2 * R;
    const int nseg = (2 * Rv + 1 + BS) / (BS + 1);
    // Jobs of a warp: first its main job (one candidate column per thread, every block row of the CTA); then, as long as
    // there are any left, "extra" jobs of 32 segments of the last columns, drawn from a shared-memory counter -- whichever
    // warps finish their main job first take them, so no warp sits at the closing barrier while work is left.
    const int nxjobs = (NB * NBY * nseg + 31) >> 5;
    const int wlane = tid & 31;

    if (tid == 0) {
        mbar_init(&bar, 1);
        fence_mbar_init();
        // the first window is requested before anything else, so the TMA latency hides behind the staging of the
        // current blocks and the candidate table
        mbar_arrive_expect_tx(&bar, (uint32_t)box_bytes);
        tma_load_3d(smem, &ref_map, &bar, bx0 * BS - R - a.win_lm, wy0, L.ref_plane[r_begin]);
    }
    for (int i = tid; i < NB * NBY; i += blockDim.x) sbest[i / NB][i % NB] = ~0ull;
    if (tid == 0) xjob = 0;
    {   // stage the current blocks (NB*BS x NBY*BS bytes) in shared memory, 16 B per thread-iteration
        const uint8_t* cp = a.cur_base + (size_t)L.cur_plane * a.cur_plane_bytes;
        constexpr int VPR = NB * BS / 16 > 0 ? NB * BS / 16 : 1;   // 16-byte vectors per row
        constexpr int RB = NB * BS;                               // bytes per staged row
        if constexpr (RB % 16 == 0) {
            for (int i = tid; i < NBY * BS * VPR; i += blockDim.x) {
                const int y = i / VPR, v = i - y * VPR;
                const int gy = by0 * BS + y, gx = bx0 * BS + v * 16;
                uint4 val = make_uint4(0, 0, 0, 0);
                if (gy < a.H && gx < a.W) val = *reinterpret_cast<const uint4*>(cp + (size_t)gy * a.cur_pitch + gx);
                *reinterpret_cast<uint4*>(scur + y * RB + v * 16) = val;
            }
        }
    }
    __syncthreads();

    const int wpitch = WPC ? WPC : (WW >> 2);

    uint32_t one;
    asm volatile("mov.u32 %0, 1;" : "=r"(one));  // opaque 1 so `one*u + t` stays an IMAD
    KeyCfg kc;
    kc.mbits = a.key_mbits;
    kc.scale = 1u << (a.key_mbits + a.key_l1bits);

    uint32_t parity = 0;
    int win = 0;   // windows searched so far
    const int nmid = (2 * Rv) / BS - 1;

    for (int r = r_begin; r < r_end; r++) {
        for (int ph = 0; ph < a.nphase; ph++) {
            const int px = ph & 1, py = ph >> 1;
            if (tid == 0 && (r != r_begin || ph != 0)) {
                mbar_arrive_expect_tx(&bar, (uint32_t)box_bytes);
                // box origin: 16-byte aligned column (bx0*BS - R - win_lm), rows <= 256
                tma_load_3d(smem, &ref_map, &bar, bx0 * BS - R - a.win_lm, wy0, L.ref_plane[r] + ph);
            }
            if (PACKED) {
                // utab[yy][m + BS]: L1 contribution and index of vertical offset m, bit 31 = outside the plane or the range
                // one table column per thread (its |mvy| term once), all block rows of the CTA: no division, few compares
                const int per = 2 * Rv + 1 + 2 * BS;
                for (int ii = tid; ii < per; ii += blockDim.x) {
                    const int m = ii - BS;
                    const uint32_t amvy = (uint32_t)abs(py - a.sc * Rv + a.sc * m);
                    const uint32_t ent = (amvy << kc.mbits) | (uint32_t)m;
#pragma unroll
                    for (int yy = 0; yy < NBY; yy++) {
                        const int oy = (by0 + yy) * BS;
                        const int mlo = max(Rv - R, Rv - oy);
                        const int mhi = min(Rv + R - py, a.H - py - BS - oy + Rv);
                        utab[yy][ii] = (m >= mlo && m <= mhi) ? ent : 0x80000000u;
                    }
                }
            }
            mbar_wait(&bar, parity);
            parity ^= 1u;
            {   // byte-shifted copies 1..3 of the window
                const uint32_t* c0 = reinterpret_cast<const uint32_t*>(smem);
                uint32_t* c1 = reinterpret_cast<uint32_t*>(smem + copy_stride);
                uint32_t* c2 = reinterpret_cast<uint32_t*>(smem + 2 * (size_t)copy_stride);
                uint32_t* c3 = reinterpret_cast<uint32_t*>(smem + 3 * (size_t)copy_stride);
                const int nq = (WW * rows) >> 4;   // 16-byte groups (WW is a multiple of 16)
                uint32_t k8, k16, k24;   // opaque, so that the multiplications are not turned back into shifts
                asm volatile("mov.u32 %0, 0x01000000;" : "=r"(k8));
                asm volatile("mov.u32 %0, 0x00010000;" : "=r"(k16));
                asm volatile("mov.u32 %0, 0x00000100;" : "=r"(k24));
                for (int g4 = tid; g4 < nq; g4 += blockDim.x) {
                    const uint4 v = reinterpret_cast<const uint4*>(c0)[g4];
                    const uint32_t nx = c0[4 * g4 + 4];   // first word of the next group (the 32-byte gap after the last one)
                    reinterpret_cast<uint4*>(c1)[g4] = make_uint4(funnel_fma(v.x, v.y, k8), funnel_fma(v.y, v.z, k8),
                                                                  funnel_fma(v.z, v.w, k8), funnel_fma(v.w, nx, k8));
                    reinterpret_cast<uint4*>(c2)[g4] = make_uint4(funnel_fma(v.x, v.y, k16), funnel_fma(v.y, v.z, k16),
                                                                  funnel_fma(v.z, v.w, k16), funnel_fma(v.w, nx, k16));
                    reinterpret_cast<uint4*>(c3)[g4] = make_uint4(funnel_fma(v.x, v.y, k24), funnel_fma(v.y, v.z, k24),
                                                                  funnel_fma(v.z, v.w, k24), funnel_fma(v.w, nx, k24));
                }
            }
            __syncthreads();

            // every warp draws until it misses once, so a window uses nwarps + nxjobs tickets, the first nxjobs of them good
            const int xbeg = win * (((int)blockDim.x >> 5) + nxjobs), xend = xbeg + nxjobs;
            for (int job = -1;;) {
            const bool is_extra = job >= 0;
            int b, dx, m0 = 0, yye = 0;
            bool active;
            if (!is_extra) {
                b = (int)div_magic((uint32_t)tid, a.r2_magic, a.r2_shift);
                dx = tid - b * 2 * R - R;
                active = tid < nmain;
            } else {
                const int e = (job - xbeg) * 32 + wlane;
                yye = e / (NB * nseg);
                const int e2 = e - yye * NB * nseg;
                b = e2 / nseg;
                const int seg = e2 - b * nseg;
                dx = R;
                m0 = min(seg * (BS + 1), 2 * Rv - BS);  // overlapping the previous segment is harmless for an argmin
                active = yye >= yy0 && yye < yy1;     // (also false beyond the last segment: yye >= NBY)
            }
            active = active && (bx0 + b < a.bw);
            const int ox = (bx0 + b) * BS;
            // this thread's window column (left edge of the candidate) and the aligned copy it reads
            const int X = a.win_lm + b * BS + dx + R;
            const uint32_t* colp = reinterpret_cast<const uint32_t*>(smem + (size_t)(X & 3) * copy_stride) + (X >> 2);
            const int mvx = a.sc * dx + px;
            const bool xvalid = active && (ox + dx >= 0) && (ox + dx + BS <= a.W - px) && (dx <= R - px);
            if (xvalid) {
                const uint32_t absmx = (uint32_t)abs(mvx);
                const uint32_t tthr = PACKED ? 0u : absmx;   // packed keys take |mvx| after the bodies
                const int mvy0 = py - a.sc * Rv;  // mvy = mvy0 + sc*m
                // main threads: every block row of the CTA; extra threads: their one (row, segment)
                const int yy_lo = is_extra ? yye : yy0, yy_hi = is_extra ? yye + 1 : yy1;
                for (int yy = yy_lo; yy < yy_hi; yy++) {
                    if (by0 + yy >= a.bh) break;
                    const int oy = (by0 + yy) * BS;
                    const int mlo = max(Rv - R, Rv - oy);
                    const int mhi = min(Rv + R - py, a.H - py - BS - oy + Rv);
                    CurBlock<BS> cur;
                    load_cur<BS>(cur, scur + (size_t)yy * BS * (NB * BS) + b * BS, NB * BS);
                    uint32_t acc[BS];
#pragma unroll
                    for (int i = 0; i < BS; i++) acc[i] = 0;
                    uint32_t best = 0xFFFFFFFFu, bestm = 0;
                    const int n1 = 2 * R + 1;
                    uint16_t* smap = nullptr;
                    if (SADMAP) {   // map row = vertical offset + R (Rv == R whenever a map is asked for, see me_can_map)
                        const size_t blk = (size_t)(by0 + yy) * a.bw + bx0 + b;
                        smap = a.sad_map + ((((size_t)lane * a.max_refs + r) * a.nphase + ph) * a.nblk + blk) * (size_t)a.map_stride + (dx + R);
                    }
                    // Vertical offsets outside the plane are masked through utab, but sliding over them costs as much as over
                    // valid ones (top and bottom block rows: up to half of the 2R+1 offsets, 2.2 % of a 1080p r=32 frame).  A
                    // main thread therefore starts at the first offset that can be valid and runs only as many steady bodies as
                    // the valid span needs (warp-uniform: mlo / mhi depend on the block row only); a segment of the extra warps
                    // that lies outside the span is skipped.
                    int m0y = m0, nm = 0;
                    if (!is_extra) {
                        nm = max(0, (max(mhi - mlo, 0) + BS - 1) / BS - 1);
                        nm = min(nm, nmid);
                        m0y = max(0, min(mlo, 2 * Rv - BS * (nm + 1)));
                    } else if (m0 + BS < mlo || m0 > mhi) {
                        continue;
                    }
                    const uint32_t* rowp = colp + ((yy - yy0) * BS + m0y) * wpitch;
                    int mbase = m0y - (BS - 1);
                    const uint32_t* ut = &utab[yy][0] + BS + mbase;
                    me_body<BS, BODY_FIRST, PACKED, SADMAP, WPC != 0>(cur, acc, rowp, wpitch, mbase, ut, mlo, mhi, mvy0, a.sc, tthr, one, kc, best, bestm, smap, n1);
                    rowp += BS * wpitch;
                    mbase += BS;
                    ut += BS;
                    for (int i = 0; i < nm; i++) {
                        me_body<BS, BODY_MID, PACKED, SADMAP, WPC != 0>(cur, acc, rowp, wpitch, mbase, ut, mlo, mhi, mvy0, a.sc, tthr, one, kc, best, bestm, smap, n1);
                        rowp += BS * wpitch;
                        mbase += BS;
                        ut += BS;
                    }
                    me_body<BS, BODY_LAST, PACKED, SADMAP, WPC != 0>(cur, acc, rowp, wpitch, mbase, ut, mlo, mhi, mvy0, a.sc, tthr, one, kc, best, bestm, smap, n1);
                    if (PACKED ? (best < 0x80000000u) : (best != 0xFFFFFFFFu)) {
                        uint32_t hi;
                        if (PACKED) {
                            best += absmx << kc.mbits;   // the L1 field holds |mvx| + |mvy| <= 2 * Rh: no carry into the SAD field
                            bestm = best & ((1u << kc.mbits) - 1u);
                            const uint32_t l1 = (best >> kc.mbits) & ((1u << a.key_l1bits) - 1u);
                            hi = ((best >> (kc.mbits + a.key_l1bits)) << 9) | l1;
                        } else {
                            hi = best;
                        }
                        const int mvy = mvy0 + a.sc * (int)bestm;
                        const uint32_t lo = ((uint32_t)r << 20) | ((uint32_t)(mvy + a.Rh) << 10) | (uint32_t)(mvx + a.Rh);
                        atomicMin(&sbest[yy][b], ((unsigned long long)hi << 32) | lo);
                    }
                }
            }
            // next job of this warp
            int nj = 0;
            if (wlane == 0) nj = atomicAdd(&xjob, 1);
            nj = __shfl_sync(0xffffffffu, nj, 0);
            if (nj >= xend) break;
            job = nj;
            }
            win++;
            __syncthreads();  // everyone is done with the window before the next TMA overwrites it
        }
    }
    for (int i = tid; i < NB * NBY; i += blockDim.x) {
        const int yy = i / NB, bb = i - yy * NB;
        if (yy >= yy0 && yy < yy1 && bx0 + bb < a.bw && by0 + yy < a.bh) {
            const unsigned long long k = sbest[yy][bb];
            const uint32_t hi = (uint32_t)(k >> 32), lo = (uint32_t)k;
            int4 o;
            o.x = (int)(lo & 1023u) - a.Rh;
            o.y = (int)((lo >> 10) & 1023u) - a.Rh;
            o.z = (int)(lo >> 20);
            o.w = (int)(hi >> 9);
            a.out[(size_t)lane * a.nblk + (size_t)(by0 + yy) * a.bw + bx0 + bb] = o;
        }
    }
}

// Generic fallback for (block, range) pairs the tiled kernel does not cover (2R not a multiple of
// BS, BS = 2/32, ...): one CTA per block, threads stride over candidates, bytes straight from L1/L2.
// Same key and tie-break; correctness path, not a performance path.
__global__ void __launch_bounds__(256) me_generic_kernel(MeArgs a, const uint8_t* ref_base, size_t ref_plane_bytes, int ref_pitch) {
    __shared__ unsigned long long sbest;
    const int bs = a.bs;
    const int bx = blockIdx.x, by = blockIdx.y, lane = blockIdx.z;
    const MeLane& L = a.lanes[lane];
    const int ox = bx * bs, oy = by * bs;
    if (threadIdx.x == 0) sbest = ~0ull;
    __syncthreads();
    const uint8_t* cur = a.cur_base + (size_t)L.cur_plane * a.cur_plane_bytes + (size_t)oy * a.cur_pitch + ox;
    // the current block as 32-bit words in shared memory; reference rows are read as aligned words and funnel-shifted, so a
    // candidate costs bs*bs/4 VABSDIFF4 instead of bs*bs byte operations
    __shared__ uint32_t scur[32 * 32 / 4];
    const int wpr = bs >> 2;
    for (int i = threadIdx.x; i < bs * wpr; i += blockDim.x)
        scur[i] = *reinterpret_cast<const uint32_t*>(cur + (size_t)(i / wpr) * a.cur_pitch + 4 * (i % wpr));
    __syncthreads();
    const int R = a.R, n1 = 2 * R + 1;
    unsigned long long best = ~0ull;
    const int per_plane = n1 * n1;
    const int total = L.nref * a.nphase * per_plane;
    for (int c = threadIdx.x; c < total; c += blockDim.x) {
        const int rp = c / per_plane, q = c - rp * per_plane;
        const int r = rp / a.nphase, ph = rp - r * a.nphase;
        const int px = ph & 1, py = ph >> 1;
        const int dy = q / n1 - R, dx = q - (q / n1) * n1 - R;
        if (ox + dx < 0 || oy + dy < 0 || ox + dx + bs > a.W - px || oy + dy + bs > a.H - py || dx > R - px || dy > R - py) continue;
        const uint8_t* ref = ref_base + (size_t)(L.ref_plane[r] + ph) * ref_plane_bytes + (size_t)(oy + dy) * ref_pitch + (ox + dx);
        const uintptr_t ad = reinterpret_cast<uintptr_t>(ref);
        const uint32_t sh = (uint32_t)(ad & 3) * 8;      // the same for every row: the pitch is a multiple of 16
        const uint32_t* rowp = reinterpret_cast<const uint32_t*>(ad & ~(uintptr_t)3);
        uint32_t s = 0;
        for (int y = 0; y < bs; y++, rowp += ref_pitch >> 2) {
            uint32_t lo = rowp[0];
            for (int w = 0; w < wpr; w++) {
                const uint32_t hi = rowp[w + 1];
                s = sad4(__funnelshift_r(lo, hi, sh), scur[y * wpr + w], s);
                lo = hi;
            }
        }
        const int mvx = a.sc * dx + px, mvy = a.sc * dy + py;
        const uint32_t hi = (s << 9) + (uint32_t)(abs(mvx) + abs(mvy));
        const uint32_t lo = ((uint32_t)r << 20) | ((uint32_t)(mvy + a.Rh) << 10) | (uint32_t)(mvx + a.Rh);
        const unsigned long long k = ((unsigned long long)hi << 32) | lo;
        if (k < best) best = k;
    }
    atomicMin(&sbest, best);
    __syncthreads();
    if (threadIdx.x == 0) {
        const uint32_t hi = (uint32_t)(sbest >> 32), lo = (uint32_t)sbest;
        int4 o;
        o.x = (int)(lo & 1023u) - a.Rh;
        o.y = (int)((lo >> 10) & 1023u) - a.Rh;
        o.z = (int)(lo >> 20);
        o.w = (int)(hi >> 9);
        a.out[(size_t)lane * a.nblk + (size_t)by * a.bw + bx] = o;
    }
}

// resident CTAs of the whole GPU for a kernel / block size / dynamic shared memory size (cached per device)
template <typename K>
static int resident_slots(K kernel, int threads, size_t smem, int* cache) {
    int& v = cache[current_device_slot()];
    if (v > 0) return v;
    int per_sm = 0, dev = 0, sms = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, threads, smem) != cudaSuccess || per_sm < 1) per_sm = 1;
    cudaGetDevice(&dev);
    if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || sms < 1) sms = 148;
    v = per_sm * sms;
    return v;
}

template <int BS, int NB, int NBY, bool PACKED, bool SADMAP, int WPC>
cudaError_t launch_tiled_pmw(const CUtensorMap& map, MeArgs a, int lanes, cudaStream_t st) {
    const int R = a.R;
    const MeTileCfg cfg = me_tile_config(BS, R);
    a.Rv = cfg.Rv;
    const int nmain = NB * 2 * R;
    const int nseg = (2 * a.Rv + 1 + BS) / (BS + 1);
    (void)nseg;
    const int threads = (nmain + 31) & ~31;   // the last columns are extra jobs of the same warps (see the kernel)
    a.win_pitch = cfg.win_pitch;
    a.win_lm = cfg.win_lm;
    a.win_copy_bytes = ((cfg.win_pitch * (NBY * BS + 2 * cfg.Rv) + 127) / 128) * 128;   // (NBY may be the tall shape's)
    const size_t smem = 4 * (size_t)(a.win_copy_bytes + 32) + (size_t)NBY * BS * NB * BS + 16;
    static size_t configured_dev[BVC_MAX_DEVICES] = {};
    size_t& configured = configured_dev[current_device_slot()];
    if (smem > configured) {
        cudaError_t e = cudaFuncSetAttribute(me_tiled_kernel<BS, NB, NBY, PACKED, SADMAP, WPC>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        // largest shared-memory carve-out: CTAs of other kernels (another lane group's tail) never wait for a re-partition
        e = cudaFuncSetAttribute(me_tiled_kernel<BS, NB, NBY, PACKED, SADMAP, WPC>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
        if (e != cudaSuccess) return e;
        configured = smem;
    }
    a.tiles_x = (a.bw + NB - 1) / NB;
    a.tiles_y = (a.bh + NBY - 1) / NBY;
    div_magic_constants((uint32_t)(a.tiles_x * a.tiles_y), a.perz_magic, a.perz_shift);
    div_magic_constants((uint32_t)a.tiles_x, a.tx_magic, a.tx_shift);
    div_magic_constants((uint32_t)(2 * R), a.r2_magic, a.r2_shift);
    const long long total = (long long)a.tiles_x * a.tiles_y * (SADMAP ? lanes * a.max_refs : lanes);
    // whole waves of full tiles, then the remaining tiles as NBY one-row CTAs each (see the kernel)
    static int slots_dev[BVC_MAX_DEVICES] = {};
    const int slots = resident_slots(me_tiled_kernel<BS, NB, NBY, PACKED, SADMAP, WPC>, threads, smem, slots_dev);
    long long rem = (NBY > 1 && a.tail_split) ? total % slots : 0;
    if (rem == total && total * NBY <= slots) rem = 0;   // a launch that does not even fill the GPU once: splitting cannot help the makespan
    a.n_full = (int)(total - rem);
    const long long grid = a.n_full + rem * NBY;
    if (grid > 0x7fffffffLL) return cudaErrorInvalidValue;
    me_tiled_kernel<BS, NB, NBY, PACKED, SADMAP, WPC><<<(unsigned)grid, threads, smem, st>>>(map, a);
    return cudaGetLastError();
}
template <int BS, int NB, int NBY, bool PACKED, bool SADMAP>
cudaError_t launch_tiled_pm(const CUtensorMap& map, MeArgs a, int lanes, cudaStream_t st) {
    // constant-pitch instantiations for the two headline shapes: 1080p r=32 (4x4 blocks, 128-byte window rows) and
    // 4K r=64 (2x2 blocks, 160-byte rows)
    if constexpr (BS == 16 && PACKED && !SADMAP && ((NB == 4 && (NBY == 4 || NBY == 8)) || (NB == 2 && NBY == 2))) {
        constexpr int WP = (NB == 4) ? 32 : 40;
        if (me_tile_config(BS, a.R).win_pitch == WP * 4) return launch_tiled_pmw<BS, NB, NBY, PACKED, SADMAP, WP>(map, a, lanes, st);
    }
    return launch_tiled_pmw<BS, NB, NBY, PACKED, SADMAP, 0>(map, a, lanes, st);
}
template <int BS, int NB, int NBY, bool PACKED>
cudaError_t launch_tiled_p(const CUtensorMap& map, MeArgs a, int lanes, cudaStream_t st) {
    if (a.sad_map) {
        if (!PACKED) return cudaErrorInvalidValue;   // me_can_map() guarantees the packed key
        return launch_tiled_pm<BS, NB, NBY, true, true>(map, a, lanes, st);
    }
    return launch_tiled_pm<BS, NB, NBY, PACKED, false>(map, a, lanes, st);
}

static int bitlen(unsigned v) { int n = 0; while (v) { n++; v >>= 1; } return n; }

template <int BS, int NB, int NBY>
cudaError_t launch_tiled(const CUtensorMap& map, MeArgs a, int lanes, cudaStream_t st) {
    // packed key: SAD | L1 | m in 31 bits (bit 31 marks offsets outside the plane / the range)
    const int Rv = me_tile_config(BS, a.R).Rv;
    const int sadbits = bitlen(255u * BS * BS);
    a.key_l1bits = bitlen(2u * a.Rh);
    a.key_mbits = bitlen(2u * Rv);
    if (sadbits + a.key_l1bits + a.key_mbits <= 31 && Rv <= 128) return launch_tiled_p<BS, NB, NBY, true>(map, a, lanes, st);
    return launch_tiled_p<BS, NB, NBY, false>(map, a, lanes, st);
}

// Tile shape: NB blocks side by side (NB*BS a multiple of 16 so the left margin of the aligned TMA box is
// the same for every CTA) and NBY stacked.  Constraints: <= 544 threads, TMA box <= 256 x 256, four window copies
// <= 110 KB; shapes that let two CTAs share an SM are preferred.
struct TileShape { int nb, nby; };
TileShape pick_shape(int bs, int R, int Rv, bool tall) {
    const int cand16[] = {4, 2, 1}, cand8[] = {8, 4, 2}, cand4[] = {8, 4, 4};
    const int* c = bs == 16 ? cand16 : bs == 8 ? cand8 : cand4;
    const int lm = (16 - R % 16) % 16;
    const int nseg = (2 * Rv + 1 + bs) / (bs + 1);
    // pass 0: shapes that leave room for two CTAs per SM (<= 341 threads at 96 registers, <= 106 KB of shared memory each):
    // at r = 64 that is 2 x 2 blocks (0.85 of the VABSDIFF4 peak on the 4K workload) instead of 4 x 1 in one 544-thread
    // CTA (0.80); pass 1: anything that fits.
    for (int pass = 0; pass < 2; pass++) {
        for (int i = 0; i < 3; i++) {
            const int nb = c[i];
            const int pitch = ((lm + nb * bs + 2 * R + 15) / 16) * 16;
            const int nbys[] = {8, 4, 2, 1};
            for (int j = 0; j < 4; j++) {
                const int nby = nbys[j];
                // eight stacked rows (half as many CTA prologues, 24 instead of 32 window rows per block row): the "tall" shape
                // of the headline geometry, taken by launches of at least four waves of such CTAs (launch_me_fullsearch) --
                // with the two or three GOP lanes a rank has under strong scaling the smaller CTAs fill the GPU better
                if (nby == 8 && !(tall && bs == 16 && nb == 4 && R == 32)) continue;
                const int rows = nby * bs + 2 * Rv;
                // (selection still counts the former extra warps, so every other geometry keeps the shape it was measured with)
                const int threads = ((nb * 2 * R + 31) & ~31) + (nby == 8 ? 0 : ((nb * nby * nseg + 31) & ~31));
                const int smem = 4 * pitch * rows + nb * nby * bs * bs;
                if (pitch > 256 || rows > 256) continue;
                if (pass == 0 ? (threads <= 341 && smem <= 106 * 1024) : (threads <= 544 && smem <= 110 * 1024)) return {nb, nby};
            }
        }
    }
    return {0, 0};
}

}  // namespace

// Which search kernel serves (block size, range in plane units):
//   2R <  BS : the narrow kernel (me_narrow.cu), exact work, one thread per (block, candidate column)
//   2R >= BS : the tiled kernel; when 2R is not a multiple of BS the bodies walk a vertical range Rv > R (2*Rv the next
//              multiple of BS) and the extra offsets are masked, so the executed / algorithmic VABSDIFF4 ratio is
//              (2*Rv+1)/(2R+1) (1.0 for every BASELINE configuration)
//   else     : the generic kernel (R = 0, windows beyond the TMA box limits)
MeTileCfg me_tile_config(int bs, int R, bool tall) {
    MeTileCfg c{};
    if (!(bs == 4 || bs == 8 || bs == 16)) return c;
    if (R < 1) return c;
    if (2 * R < bs) return me_narrow_config(bs, R);
    const int Rv = (2 * R + bs - 1) / bs * bs / 2;
    const TileShape t = pick_shape(bs, R, Rv, tall);
    if (t.nb == 0) return c;
    c.tiled = true;
    c.Rv = Rv;
    c.nb = t.nb;
    c.nby = t.nby;
    c.win_lm = (16 - R % 16) % 16;
    c.win_pitch = ((c.win_lm + t.nb * bs + 2 * R + 15) / 16) * 16;
    c.rows = t.nby * bs + 2 * Rv;
    return c;
}

template <int BS, int NB>
static cudaError_t launch_by_nby(const MeTileCfg& cfg, const CUtensorMap& map, const MeArgs& a, int lanes, cudaStream_t st) {
    if constexpr (BS == 16 && NB == 4) {
        if (cfg.nby == 8) return launch_tiled<BS, NB, 8>(map, a, lanes, st);
    }
    if (cfg.nby == 4) return launch_tiled<BS, NB, 4>(map, a, lanes, st);
    if (cfg.nby == 2) return launch_tiled<BS, NB, 2>(map, a, lanes, st);
    return launch_tiled<BS, NB, 1>(map, a, lanes, st);
}

cudaError_t launch_me_fullsearch(const CUtensorMap* ref_map, const MeArgs& args, int lanes, const uint8_t* ref_base,
                                 size_t ref_plane_bytes, int ref_pitch, cudaStream_t st, const CUtensorMap* tall_map) {
    MeArgs a = args;
    MeTileCfg cfg = me_tile_config(a.bs, a.R);
    if (tall_map && cfg.tiled && !a.sad_map) {
        const MeTileCfg tc = me_tile_config(a.bs, a.R, true);
        const long long ctas = (long long)((a.bw + tc.nb - 1) / tc.nb) * ((a.bh + tc.nby - 1) / tc.nby) * lanes;
        if (tc.tiled && tc.nby == 8 && a.tall_mode >= 0 && (a.tall_mode > 0 || ctas >= 4 * 296)) { cfg = tc; ref_map = tall_map; }
    }
    if (cfg.narrow && ref_map && !a.sad_map) return launch_me_narrow(*ref_map, a, lanes, st);
    if (cfg.tiled && ref_map) {
        if (a.bs == 16) {
            if (cfg.nb == 4) return launch_by_nby<16, 4>(cfg, *ref_map, a, lanes, st);
            if (cfg.nb == 2) return launch_by_nby<16, 2>(cfg, *ref_map, a, lanes, st);
            return launch_by_nby<16, 1>(cfg, *ref_map, a, lanes, st);
        } else if (a.bs == 8) {
            if (cfg.nb == 8) return launch_by_nby<8, 8>(cfg, *ref_map, a, lanes, st);
            if (cfg.nb == 4) return launch_by_nby<8, 4>(cfg, *ref_map, a, lanes, st);
            return launch_by_nby<8, 2>(cfg, *ref_map, a, lanes, st);
        } else {
            if (cfg.nb == 8) return launch_by_nby<4, 8>(cfg, *ref_map, a, lanes, st);
            return launch_by_nby<4, 4>(cfg, *ref_map, a, lanes, st);
        }
    }
    if (a.sad_map) return cudaErrorInvalidValue;   // callers ask me_can_map() first
    dim3 grid(a.bw, a.bh, lanes);
    me_generic_kernel<<<grid, 256, 0, st>>>(a, ref_base, ref_plane_bytes, ref_pitch);
    return cudaGetLastError();
}

bool me_can_map(int bs, int R) {
    const MeTileCfg c = me_tile_config(bs, R);
    if (!c.tiled || c.Rv != R || R > 128) return false;
    // the packed key (SAD | L1 | m in 31 bits) must fit; R here is already in plane units, L1 in MV units <= 4R
    return bitlen(255u * bs * bs) + bitlen(4u * R) + bitlen(2u * R) <= 31;
}

}  // namespace bvc
