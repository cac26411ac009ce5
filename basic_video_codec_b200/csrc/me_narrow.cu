// me_narrow.cu -- K1/K3 for narrow search ranges (2R < block size): the reference's own runs use r = 2, 3, 4 with
// i = 8 / 16 (results.csv, assign1/ex4_plots.py) and BASELINE config 3 is i = 16, r = 4 with half-pel vectors.
//
// Same contract as me_fullsearch.cu (find_lowest_mae_block + get_ref_block_at_mv + is_out_of_range + common.mae,
// reference encoder/block_predictor.py:61-143, common.py:43-45; winner = min (SAD, |mvx|+|mvy|, ref, mvy, mvx)).
//
// With 2R+1 < BS vertical offsets there is no steady state for the rotating-accumulator body of the tiled kernel,
// and 2R candidate columns per block do not fill warps.  Here:
//   * a tile = 128 pixels x 4 block rows (NBX = 128/BS blocks side by side); its window (128 + 2R) x (4*BS + 2R) comes
//     in by one TMA tile load per pass = (reference, phase plane).  The kernel is persistent (one CTA per resident slot,
//     tiles strided over the CTAs) and runs without CTA barriers: a ring of four TMA windows, per-slot mbarriers for
//     "window landed" and a per-slot count of the warps that are done with it; the warp that counts last requests the
//     window four passes ahead.  Every warp streams through its passes on its own.  A candidate row is five aligned
//     LDS.32 (the four lanes of a block read the same words: broadcast) realigned by BS/4 funnel shifts -- 1 SHF per 9
//     VABSDIFF4 on the same pipe at R = 4, cheaper than deriving shifted copies of every window, which needed two
//     barriers per pass and kept the pipe at 59 % (profiles/r2_ncu_narrow_*.csv, history in profiles/r2_experiments.md).
//   * one thread = one candidate column dx of one block, ALL 2R+1 vertical offsets at once: NM = 2R+1 accumulators,
//     the BS + 2R window rows fully unrolled, every row (BS/4 LDS.32) feeding up to NM candidates.  The executed
//     VABSDIFF4 count equals the algorithmic count (NM * BS * BS/4 per column).  The current block stays in
//     registers for all passes.
//   * lanes of a warp = NBX blocks x G = BS/4 adjacent columns ("column group"): 8 distinct words per load at BS = 16,
//     no bank conflicts.  The 2R+1 columns are NQ full groups plus NL left-over columns; a left-over column takes one
//     lane per block of the tile (4-way bank conflicts on 1/9 of the work at R = 4, every lane busy).
//   * argmin as in the tiled kernel: packed key SAD | L1 | m per candidate (one IMAD + one VIMNMX), a 64-bit key
//     (SAD, L1, ref, mvy, mvx) per thread across passes, one shared-memory atomicMin per thread at the end.
#include "bvc_common.cuh"
#include "bvc_kernels.h"

namespace bvc {
namespace {

constexpr int NARROW_PITCH = 160;   // window row bytes: lm + 128 + 2R <= 151 for R <= 7
// Block rows per tile.  The search loop wants 128 registers per thread (the current block alone is 64): with less the
// compiler cannot keep a window row's loads in flight behind the previous row's VABSDIFF4s (r = 4 at 96 registers: 0.31 of
// the peak; r = 7 at 128: 0.60).  128 registers = at most 16 warps per SM, so a tile is as many block rows as give the
// most warps <= 16 in ONE CTA per SM (TMA box <= 256 rows); ties go to the smaller tile (less padding at the frame bottom).
constexpr int narrow_warps(int bs, int R, int nby) {
    const int g = bs / 4, nbx = 32 / g, nc = 2 * R + 1, nq = nc / g, nl = nc - nq * g;
    return nby * nq + nl * ((nby * nbx + 31) / 32);
}
constexpr int narrow_nby(int bs, int R) {
    int best = 1, bw = 0;
    for (int nby = 1; nby <= 15; nby++) {
        const int w = narrow_warps(bs, R, nby);
        if (w <= 16 && nby * bs + 2 * R <= 256 && w > bw) { best = nby; bw = w; }
    }
    return best;
}
constexpr int NARROW_STAGES = 4;    // window ring depth
constexpr int KEY_MBITS = 4;        // m <= 14
constexpr int KEY_L1BITS = 6;       // |mvx| + |mvy| <= 28 (half-pel units, R <= 7)

template <int BS, int R>
struct NarrowCfg {
    static constexpr int G = BS / 4;              // columns per block inside a column-group warp
    static constexpr int NBX = 32 / G;            // blocks side by side (128 pixels)
    static constexpr int NBY = narrow_nby(BS, R);
    static constexpr int NM = 2 * R + 1;          // vertical offsets = accumulators per thread
    static constexpr int NC = 2 * R + 1;          // candidate columns per block
    static constexpr int NQ = NC / G;             // full column groups
    static constexpr int NL = NC - NQ * G;        // left-over columns
    static constexpr int LW = (NBY * NBX + 31) / 32;   // warps per left-over column (one lane per block of the tile)
    static constexpr int WARPS = NBY * NQ + NL * LW;
    static constexpr int THREADS = WARPS * 32;
    static constexpr int ROWS = NBY * BS + 2 * R;
    static constexpr int LM = (16 - R % 16) % 16; // left margin: the TMA box origin must be 16-byte aligned
    static constexpr int WP = NARROW_PITCH / 4;   // window pitch in words
    static constexpr int CB = (NARROW_PITCH * ROWS + 32 + 127) / 128 * 128;   // bytes reserved per window (TMA destination)
    static constexpr int SMEM = NARROW_STAGES * CB;
    static constexpr int MIN_CTAS = 1;
    static_assert(WARPS <= 16, "128 registers per thread");
    static_assert(2 * R < BS, "the tiled kernel serves 2R >= BS");
    static_assert(LM + 128 + 2 * R <= NARROW_PITCH, "window row does not fit the pitch");
    static_assert(NM - 1 < (1 << KEY_MBITS), "m field");
    static_assert(NM <= 32, "the candidate table of a warp is built by its first NM lanes");
};

// Position of a CTA in its sequence of passes: the CTA owns the contiguous tile range [tile, tend); pass p of npass =
// references x phase planes of the tile's lane.  Tile coordinates (lane z, tile row ty, tile column tx) advance by carry,
// so the per-pass bookkeeping has no integer divisions and -- when every lane of the launch sees the same number of
// references (always, on the clip path) -- no memory access.
struct PassIter {
    int tile, tend, p, npass;
    int z, ty, tx;
};

__device__ __forceinline__ int lane_npass(const MeArgs& a, int z) {
    return (a.uniform_nref > 0 ? a.uniform_nref : a.lanes[z].nref) * a.nphase;
}
__device__ __forceinline__ void iter_step_tile(PassIter& it, const MeArgs& a) {
    it.tile++;
    if (++it.tx == a.tiles_x) { it.tx = 0; if (++it.ty == a.tiles_y) { it.ty = 0; it.z++; } }
}
__device__ __forceinline__ bool iter_settle(PassIter& it, const MeArgs& a) {   // first tile at or after it.tile with passes to do
    while (it.tile < it.tend) {
        it.npass = lane_npass(a, it.z);
        if (it.npass > 0) return true;
        iter_step_tile(it, a);
    }
    return false;
}
__device__ __forceinline__ bool iter_next(PassIter& it, const MeArgs& a) {
    if (++it.p < it.npass) return true;
    it.p = 0;
    iter_step_tile(it, a);
    return iter_settle(it, a);
}

template <int BS, int R>
struct TilePos {
    int z, bx0, by0;
    __device__ __forceinline__ explicit TilePos(const PassIter& it)
        : z(it.z), bx0(it.tx * NarrowCfg<BS, R>::NBX), by0(it.ty * NarrowCfg<BS, R>::NBY) {}
};

// Persistent, barrier-free kernel.  Grid = resident CTAs (or fewer tiles); every CTA walks the passes of its tiles.  The
// windows live in a ring of NARROW_STAGES TMA buffers: a warp waits for the window of its pass on that buffer's mbarrier,
// searches it, and counts itself off the buffer; the warp that counts last requests the window NARROW_STAGES passes ahead
// into it.  Warps of a CTA are therefore never more than NARROW_STAGES passes apart but otherwise run independently: no
// __syncthreads, no warp ever waits for another warp's search -- the ALU pipe only idles when all of an SM's warps wait
// for memory at once.  A tile's winners meet in shared memory (64-bit atomicMin); the warp that finishes a tile last
// writes them out.
template <int BS, int R>
__global__ void __launch_bounds__(NarrowCfg<BS, R>::THREADS, NarrowCfg<BS, R>::MIN_CTAS) me_narrow_kernel(const __grid_constant__ CUtensorMap ref_map, MeArgs a) {
    using C = NarrowCfg<BS, R>;
    constexpr int WPR = BS / 4, NM = C::NM, WP = C::WP, NST = NARROW_STAGES;
    extern __shared__ __align__(128) uint8_t smem[];
    __shared__ uint64_t full[NST];                       // TMA completion per ring slot
    __shared__ int released[NST];                        // warps done with the slot's current window
    __shared__ unsigned long long sbest[NST][C::NBY][C::NBX];   // winners per block, ring over tiles
    __shared__ int tile_done[NST];                       // warps done with the tile
    __shared__ uint32_t utab[C::WARPS][NM + 1];          // per-warp candidate table of the current pass

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    constexpr uint32_t box_bytes = (uint32_t)(NARROW_PITCH * C::ROWS);

    PassIter it;
    {   // this CTA's share of the tiles: a contiguous range (neighbouring tiles share window rows through L2)
        const long long n = a.n_tiles, g = gridDim.x, i = blockIdx.x;
        it.tile = (int)(n * i / g);
        it.tend = (int)(n * (i + 1) / g);
        it.p = 0; it.npass = 0;
        const int per_z = a.tiles_x * a.tiles_y;
        it.z = it.tile / per_z;
        const int t2 = it.tile - it.z * per_z;
        it.ty = t2 / a.tiles_x;
        it.tx = t2 - it.ty * a.tiles_x;
    }
    if (!iter_settle(it, a)) return;

    // ---- role of this thread: (block row yy, block b, column index dxi = dx + R); the same for every tile ----
    int yy, b, dxi;
    if (warp < C::NBY * C::NQ) {
        yy = warp / (C::NQ > 0 ? C::NQ : 1);
        const int q = warp - yy * C::NQ;
        b = lane / C::G;
        dxi = q * C::G + (lane % C::G);
    } else {
        const int w2 = warp - C::NBY * C::NQ;
        const int l = w2 / C::LW, part = w2 - l * C::LW;
        const int i = part * 32 + lane;
        yy = i / C::NBX;
        b = i - yy * C::NBX;
        dxi = C::NQ * C::G + l;
    }
    const int dx = dxi - R;
    // window column of the candidate's left edge: words X>>2 .. of every row, realigned by a funnel shift of (X&3) bytes.
    // The lanes of one block read the same words (broadcast), different blocks different banks.
    const int X = C::LM + b * BS + dxi;
    const uint32_t shift = (uint32_t)(X & 3) * 8u;
    const int word_off = (X >> 2) + yy * BS * WP;

    auto issue_tma = [&](const PassIter& ps, int slot) {
        const int r = a.nphase == 4 ? ps.p >> 2 : ps.p, ph = ps.p & (a.nphase - 1);   // nphase is 1 or 4
        mbar_arrive_expect_tx(&full[slot], box_bytes);
        tma_load_3d(smem + slot * C::CB, &ref_map, &full[slot], ps.tx * C::NBX * BS - R - C::LM, ps.ty * C::NBY * BS - R,
                    a.lanes[ps.z].ref_plane[r] + ph);
    };

    PassIter ahead = it;                      // the pass NST ahead of `it` (every warp advances its own copy in step)
    bool has_ahead = true;
    if (tid == 0) {
        for (int i = 0; i < NST; i++) { mbar_init(&full[i], 1); released[i] = 0; tile_done[i] = 0; }
        fence_mbar_init();
    }
    for (int i = tid; i < NST * C::NBY * C::NBX; i += C::THREADS) (&sbest[0][0][0])[i] = ~0ull;
    __syncthreads();                          // the only CTA-wide barrier: ring state initialised
    for (int i = 0; i < NST; i++) {
        if (has_ahead && tid == 0) issue_tma(ahead, i);
        if (has_ahead) has_ahead = iter_next(ahead, a);
    }

    uint32_t one;
    asm volatile("mov.u32 %0, 1;" : "=r"(one));   // opaque 1: `one*u + t` stays an IMAD (FMA pipe), off the ALU pipe
    constexpr uint32_t scale = 1u << (KEY_MBITS + KEY_L1BITS);
    uint32_t cur[BS][WPR];
    unsigned long long tbest = ~0ull;
    TilePos<BS, R> tp(it);
    bool blk_ok = false;
    int ktile = 0;                            // tiles this CTA has started: ring index of sbest / tile_done

    for (int s = 0;; s++) {
        const int slot = s % NST;
        if (it.p == 0) {
            // a new tile: this thread's current block, straight from global memory into registers (the four lanes of a
            // block read the same 16 bytes per row; the latency is covered by the other warps, which do not wait for us)
            tp = TilePos<BS, R>(it);
            blk_ok = (yy < C::NBY) && (tp.bx0 + b < a.bw) && (tp.by0 + yy < a.bh);
            tbest = ~0ull;
            const uint8_t* cp = a.cur_base + (size_t)a.lanes[tp.z].cur_plane * a.cur_plane_bytes +
                                (size_t)((tp.by0 + yy) * BS) * a.cur_pitch + (tp.bx0 + b) * BS;
#pragma unroll
            for (int r = 0; r < BS; r++) {
                const uint8_t* row = cp + (size_t)r * a.cur_pitch;
                if constexpr (BS == 16) {
                    const uint4 v = blk_ok ? *reinterpret_cast<const uint4*>(row) : make_uint4(0, 0, 0, 0);
                    cur[r][0] = v.x; cur[r][1] = v.y; cur[r][2] = v.z; cur[r][3] = v.w;
                } else if constexpr (BS == 8) {
                    const uint2 v = blk_ok ? *reinterpret_cast<const uint2*>(row) : make_uint2(0, 0);
                    cur[r][0] = v.x; cur[r][1] = v.y;
                } else {
                    cur[r][0] = blk_ok ? *reinterpret_cast<const uint32_t*>(row) : 0u;
                }
            }
        }
        const int r = a.nphase == 4 ? it.p >> 2 : it.p, ph = it.p & (a.nphase - 1);
        const int px = ph & 1, py = ph >> 1;
        // utab[warp][m]: L1 contribution and index of vertical offset m (dy = m - R); bit 31 = leaves the plane / the range.
        // Column-group warps hold one block row, left-over warps several: the table is per lane's row there.
        if (warp < C::NBY * C::NQ) {
            if (lane < NM) {
                const int oy2 = (tp.by0 + yy) * BS, dy = lane - R;
                const bool ok = (dy <= R - py) && (oy2 + dy >= 0) && (oy2 + dy + BS <= a.H - py);
                utab[warp][lane] = ok ? (((uint32_t)abs(a.sc * dy + py) << KEY_MBITS) | (uint32_t)lane) : 0x80000000u;
            }
            __syncwarp();
        }
        const int ox = (tp.bx0 + b) * BS;
        const bool xvalid = blk_ok && (ox + dx >= 0) && (ox + dx + BS <= a.W - px) && (dx <= R - px);
        mbar_wait(&full[slot], (uint32_t)((s / NST) & 1));
        if (xvalid) {
            const uint32_t* colp = reinterpret_cast<const uint32_t*>(smem + slot * C::CB) + word_off;
            uint32_t acc[NM];
#pragma unroll
            for (int y = 0; y < BS + NM - 1; y++) {
                uint32_t w[WPR + 1];
#pragma unroll
                for (int j = 0; j <= WPR; j++) w[j] = colp[y * WP + j];
#pragma unroll
                for (int j = 0; j < WPR; j++) w[j] = __funnelshift_r(w[j], w[j + 1], shift);
#pragma unroll
                for (int wi = 0; wi < WPR; wi++) {
#pragma unroll
                    for (int m = 0; m < NM; m++) {
                        const int j = y - m;   // row j of the current block meets window row y for vertical offset m
                        if (j >= 0 && j < BS) acc[m] = sad4(w[wi], cur[j][wi], (j == 0 && wi == 0) ? 0u : acc[m]);
                    }
                }
            }
            const int mvx = a.sc * dx + px;
            const uint32_t tthr = (uint32_t)abs(mvx) << KEY_MBITS;
            uint32_t best = 0xFFFFFFFFu, pend = 0xFFFFFFFFu;
            const int oy2 = (tp.by0 + yy) * BS;
#pragma unroll
            for (int m = 0; m < NM; m++) {
                uint32_t u;
                if (warp < C::NBY * C::NQ) {
                    u = utab[warp][m];
                } else {   // left-over warps: lanes of different block rows, the entry is computed in place
                    const int dy = m - R;
                    const bool ok = (dy <= R - py) && (oy2 + dy >= 0) && (oy2 + dy + BS <= a.H - py);
                    u = ok ? (((uint32_t)abs(a.sc * dy + py) << KEY_MBITS) | (uint32_t)m) : 0x80000000u;
                }
                uint32_t u_plus_t;
                asm("mad.lo.u32 %0, %1, %2, %3;" : "=r"(u_plus_t) : "r"(one), "r"(u), "r"(tthr));
                uint32_t key;
                asm("mad.lo.u32 %0, %1, %2, %3;" : "=r"(key) : "r"(acc[m]), "r"(scale), "r"(u_plus_t));
                // two candidates per ALU-pipe instruction (VIMNMX3); NM = 2R + 1 is odd, the last one goes alone
                if ((m & 1) == 0 && m + 1 < NM) pend = key;
                else if (m & 1) best = __vimin3_u32(best, pend, key);
                else best = min(best, key);
            }
            if (best < 0x80000000u) {
                const uint32_t m = best & ((1u << KEY_MBITS) - 1u);
                const uint32_t l1 = (best >> KEY_MBITS) & ((1u << KEY_L1BITS) - 1u);
                const uint32_t hi = ((best >> (KEY_MBITS + KEY_L1BITS)) << 9) | l1;
                const int mvy = a.sc * ((int)m - R) + py;
                const uint32_t lo = ((uint32_t)r << 20) | ((uint32_t)(mvy + a.Rh) << 10) | (uint32_t)(mvx + a.Rh);
                const unsigned long long k = ((unsigned long long)hi << 32) | lo;
                tbest = k < tbest ? k : tbest;
            }
        }
        // ---- end of a tile: winners to shared memory; the warp that arrives last writes the tile out ----
        if (it.p == it.npass - 1) {
            const int ring = ktile % NST;
            if (tbest != ~0ull) atomicMin(&sbest[ring][yy][b], tbest);
            __threadfence_block();
            __syncwarp();
            int last = 0;
            if (lane == 0) last = atomicAdd(&tile_done[ring], 1) == C::WARPS - 1;
            last = __shfl_sync(0xffffffffu, last, 0);
            if (last) {
                __threadfence_block();
                for (int i = lane; i < C::NBY * C::NBX; i += 32) {
                    const int y2 = i / C::NBX, bb = i - y2 * C::NBX;
                    const unsigned long long k = sbest[ring][y2][bb];
                    sbest[ring][y2][bb] = ~0ull;         // the ring slot is used again NST tiles from now
                    if (tp.bx0 + bb < a.bw && tp.by0 + y2 < a.bh) {
                        const uint32_t hi = (uint32_t)(k >> 32), lo = (uint32_t)k;
                        int4 o;
                        o.x = (int)(lo & 1023u) - a.Rh;
                        o.y = (int)((lo >> 10) & 1023u) - a.Rh;
                        o.z = (int)(lo >> 20);
                        o.w = (int)(hi >> 9);
                        a.out[(size_t)tp.z * a.nblk + (size_t)(tp.by0 + y2) * a.bw + tp.bx0 + bb] = o;
                    }
                }
                if (lane == 0) tile_done[ring] = 0;
                __threadfence_block();
            }
            ktile++;
        }
        // ---- count this warp off the window; the last one requests the window NST passes ahead into the slot ----
        __syncwarp();
        if (lane == 0) {
            __threadfence_block();
            if (atomicAdd(&released[slot], 1) == C::WARPS - 1) {
                released[slot] = 0;
                if (has_ahead) {
                    __threadfence_block();
                    fence_proxy_async_smem();      // generic-proxy reads of the slot before the async-proxy overwrite
                    issue_tma(ahead, slot);
                }
            }
        }
        if (has_ahead) has_ahead = iter_next(ahead, a);
        if (!iter_next(it, a)) break;
    }
}

template <int BS, int R>
cudaError_t launch_narrow(const CUtensorMap& map, MeArgs a, int lanes, cudaStream_t st) {
    using C = NarrowCfg<BS, R>;
    static int slots_dev[BVC_MAX_DEVICES] = {};
    int& slots = slots_dev[current_device_slot()];
    if (slots == 0) {
        cudaError_t e = cudaFuncSetAttribute(me_narrow_kernel<BS, R>, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM);
        if (e != cudaSuccess) return e;
        e = cudaFuncSetAttribute(me_narrow_kernel<BS, R>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
        if (e != cudaSuccess) return e;
        int per_sm = 0, dev = 0, sms = 0;
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, me_narrow_kernel<BS, R>, C::THREADS, C::SMEM) != cudaSuccess || per_sm < 1) per_sm = 1;
        cudaGetDevice(&dev);
        if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || sms < 1) sms = 148;
        slots = per_sm * sms;
    }
    a.tiles_x = (a.bw + C::NBX - 1) / C::NBX;
    a.tiles_y = (a.bh + C::NBY - 1) / C::NBY;
    const long long total = (long long)a.tiles_x * a.tiles_y * lanes;
    if (total > 0x7fffffffLL) return cudaErrorInvalidValue;
    a.n_tiles = (int)total;
    const unsigned grid = (unsigned)(total < slots ? total : slots);   // persistent: one CTA per resident slot
    me_narrow_kernel<BS, R><<<grid, C::THREADS, C::SMEM, st>>>(map, a);
    return cudaGetLastError();
}

}  // namespace

MeTileCfg me_narrow_config(int bs, int R) {
    MeTileCfg c{};
    if (!(bs == 4 || bs == 8 || bs == 16) || R < 1 || 2 * R >= bs) return c;
    c.narrow = true;
    c.nb = 128 / bs;
    c.nby = narrow_nby(bs, R);
    c.win_lm = (16 - R % 16) % 16;
    c.win_pitch = NARROW_PITCH;
    c.rows = narrow_nby(bs, R) * bs + 2 * R;
    c.Rv = R;
    return c;
}

cudaError_t launch_me_narrow(const CUtensorMap& ref_map, const MeArgs& a, int lanes, cudaStream_t st) {
    switch (a.bs * 16 + a.R) {
        case 16 * 16 + 1: return launch_narrow<16, 1>(ref_map, a, lanes, st);
        case 16 * 16 + 2: return launch_narrow<16, 2>(ref_map, a, lanes, st);
        case 16 * 16 + 3: return launch_narrow<16, 3>(ref_map, a, lanes, st);
        case 16 * 16 + 4: return launch_narrow<16, 4>(ref_map, a, lanes, st);
        case 16 * 16 + 5: return launch_narrow<16, 5>(ref_map, a, lanes, st);
        case 16 * 16 + 6: return launch_narrow<16, 6>(ref_map, a, lanes, st);
        case 16 * 16 + 7: return launch_narrow<16, 7>(ref_map, a, lanes, st);
        case 8 * 16 + 1: return launch_narrow<8, 1>(ref_map, a, lanes, st);
        case 8 * 16 + 2: return launch_narrow<8, 2>(ref_map, a, lanes, st);
        case 8 * 16 + 3: return launch_narrow<8, 3>(ref_map, a, lanes, st);
        case 4 * 16 + 1: return launch_narrow<4, 1>(ref_map, a, lanes, st);
    }
    return cudaErrorInvalidValue;
}

}  // namespace bvc
