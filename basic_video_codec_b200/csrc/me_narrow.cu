// me_narrow.cu -- K1/K3 for narrow search ranges (2R < block size): the reference's own runs use r = 2, 3, 4 with
// i = 8 / 16 (results.csv, assign1/ex4_plots.py) and BASELINE config 3 is i = 16, r = 4 with half-pel vectors.
//
// Same contract as me_fullsearch.cu (find_lowest_mae_block + get_ref_block_at_mv + is_out_of_range + common.mae,
// reference encoder/block_predictor.py:61-143, common.py:43-45; winner = min (SAD, |mvx|+|mvy|, ref, mvy, mvx)).
//
// With 2R+1 < BS vertical offsets there is no steady state for the rotating-accumulator body of the tiled kernel,
// and 2R candidate columns per block do not fill warps.  Here:
//   * one CTA = a tile of 128 pixels x 4 block rows (NBX = 128/BS blocks side by side); the window
//     (128 + 2R) x (4*BS + 2R) comes in by one TMA tile load per (reference, phase plane), double buffered: the
//     load of pass p+1 is in flight while pass p is searched.  Three byte-shifted copies are derived per pass so the
//     search reads aligned words only (no PRMT / SHF next to the VABSDIFF4s, they share the ALU pipe).
//   * one thread = one candidate column dx of one block, ALL 2R+1 vertical offsets at once: NM = 2R+1 accumulators,
//     the BS + 2R window rows fully unrolled, every row (BS/4 LDS.32) feeding up to NM candidates.  The executed
//     VABSDIFF4 count equals the algorithmic count (NM * BS * BS/4 per column).  The current block stays in
//     registers for all passes.
//   * lanes of a warp = NBX blocks x G = BS/4 adjacent columns ("column group"), which makes the 32 LDS of a warp
//     hit 32 different banks when shifted copy k starts base_k = k + [k >= 4 - R%4] words after a 128-byte boundary
//     (the +1 compensates the word carry of columns left of the block's first aligned word).  The 2R+1 columns are
//     NQ full groups plus NL left-over columns; a left-over column takes one lane per block of the tile (4-way bank
//     conflicts on 1/9 of the work at R = 4, every lane busy).
//   * argmin as in the tiled kernel: packed key SAD | L1 | m per candidate (one IMAD + one VIMNMX), a 64-bit key
//     (SAD, L1, ref, mvy, mvx) per thread across passes, one shared-memory atomicMin per thread at the end.
#include "bvc_common.cuh"
#include "bvc_kernels.h"

namespace bvc {
namespace {

constexpr int NARROW_PITCH = 160;   // window row bytes: lm + 128 + 2R <= 151 for R <= 7
constexpr int NARROW_NBY = 4;
constexpr int KEY_MBITS = 4;        // m <= 14
constexpr int KEY_L1BITS = 6;       // |mvx| + |mvy| <= 28 (half-pel units, R <= 7)

template <int BS, int R>
struct NarrowCfg {
    static constexpr int G = BS / 4;              // columns per block inside a column-group warp
    static constexpr int NBX = 32 / G;            // blocks side by side (128 pixels)
    static constexpr int NBY = NARROW_NBY;
    static constexpr int NM = 2 * R + 1;          // vertical offsets = accumulators per thread
    static constexpr int NC = 2 * R + 1;          // candidate columns per block
    static constexpr int NQ = NC / G;             // full column groups
    static constexpr int NL = NC - NQ * G;        // left-over columns
    static constexpr int LW = NBY * NBX / 32;     // warps per left-over column (one lane per block of the tile)
    static constexpr int WARPS = NBY * NQ + NL * LW;
    static constexpr int THREADS = WARPS * 32;
    static constexpr int MIN_CTAS = THREADS <= 320 ? 2 : 1;   // two CTAs per SM cover each other's prologues when the registers allow
    static constexpr int ROWS = NBY * BS + 2 * R;
    static constexpr int LM = (16 - R % 16) % 16; // left margin: the TMA box origin must be 16-byte aligned
    static constexpr int WP = NARROW_PITCH / 4;   // window pitch in words
    static constexpr int CB = (NARROW_PITCH * ROWS + 16 + 127) / 128 * 128;   // bytes reserved per window copy
    static constexpr int RHO = R % 4;
    __host__ __device__ static constexpr int copy_base_words(int k) { return k + ((k >= 4 - RHO) ? 1 : 0); }
    // raw window x2 (TMA destinations, 128-byte aligned) | copies 1..3 (skewed by base_k words) | current blocks
    static constexpr int OFF_COPIES = 2 * CB;
    static constexpr int OFF_CUR = 5 * CB + 128;
    static constexpr int SMEM = OFF_CUR + NBY * BS * 128;
    static_assert(2 * R < BS, "the tiled kernel serves 2R >= BS");
    static_assert(LM + 128 + 2 * R <= NARROW_PITCH, "window row does not fit the pitch");
    static_assert(NM - 1 < (1 << KEY_MBITS), "m field");
};

template <int BS, int R>
__global__ void __launch_bounds__(NarrowCfg<BS, R>::THREADS, NarrowCfg<BS, R>::MIN_CTAS) me_narrow_kernel(const __grid_constant__ CUtensorMap ref_map, MeArgs a) {
    using C = NarrowCfg<BS, R>;
    constexpr int WPR = BS / 4, NM = C::NM, WP = C::WP;
    extern __shared__ __align__(128) uint8_t smem[];
    __shared__ uint64_t bar[2];
    __shared__ unsigned long long sbest[C::NBY][C::NBX];
    __shared__ uint32_t utab[C::NBY][NM + 1];

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int per_z = a.tiles_x * a.tiles_y;
    const int z = (int)blockIdx.x / per_z, t2 = (int)blockIdx.x - z * per_z;
    const int ty = t2 / a.tiles_x, tx = t2 - ty * a.tiles_x;
    const int bx0 = tx * C::NBX, by0 = ty * C::NBY;
    const MeLane& L = a.lanes[z];
    const int npass = L.nref * a.nphase;
    if (npass <= 0) return;
    const int wx0 = bx0 * BS - R - C::LM, wy0 = by0 * BS - R;
    constexpr uint32_t box_bytes = (uint32_t)(NARROW_PITCH * C::ROWS);

    // ---- role of this thread: (block row yy, block b, column index dxi = dx + R) ----
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
    const int ox = (bx0 + b) * BS;
    const bool blk_ok = (bx0 + b < a.bw) && (by0 + yy < a.bh);

    if (tid == 0) {
        mbar_init(&bar[0], 1);
        mbar_init(&bar[1], 1);
        fence_mbar_init();
        mbar_arrive_expect_tx(&bar[0], box_bytes);
        tma_load_3d(smem, &ref_map, &bar[0], wx0, wy0, L.ref_plane[0]);
    }
    for (int i = tid; i < C::NBY * C::NBX; i += C::THREADS) sbest[i / C::NBX][i % C::NBX] = ~0ull;
    uint8_t* scur = smem + C::OFF_CUR;
    {   // current pixels of the tile: NBY*BS rows of 128 bytes
        const uint8_t* cp = a.cur_base + (size_t)L.cur_plane * a.cur_plane_bytes;
        for (int i = tid; i < C::NBY * BS * 8; i += C::THREADS) {
            const int y = i >> 3, v = i & 7;
            const int gy = by0 * BS + y, gx = bx0 * BS + v * 16;
            uint4 val = make_uint4(0, 0, 0, 0);
            if (gy < a.H && gx < a.W) val = *reinterpret_cast<const uint4*>(cp + (size_t)gy * a.cur_pitch + gx);   // W is a multiple of BS, the pitch of 16
            *reinterpret_cast<uint4*>(scur + y * 128 + v * 16) = val;
        }
    }
    __syncthreads();

    // current block -> registers, once for all passes
    uint32_t cur[BS][WPR];
#pragma unroll
    for (int r = 0; r < BS; r++) {
        const uint8_t* row = scur + (yy * BS + r) * 128 + b * BS;
        if constexpr (BS == 16) {
            const uint4 v = *reinterpret_cast<const uint4*>(row);
            cur[r][0] = v.x; cur[r][1] = v.y; cur[r][2] = v.z; cur[r][3] = v.w;
        } else if constexpr (BS == 8) {
            const uint2 v = *reinterpret_cast<const uint2*>(row);
            cur[r][0] = v.x; cur[r][1] = v.y;
        } else {
            cur[r][0] = *reinterpret_cast<const uint32_t*>(row);
        }
    }

    // window column of the candidate's left edge; (X & 3) picks the shifted copy
    const int X = C::LM + b * BS + dxi;
    const int ksel = X & 3;
    const int word_off = (X >> 2) + yy * BS * WP;
    const uint32_t* copyp = nullptr;
    if (ksel) copyp = reinterpret_cast<const uint32_t*>(smem + C::OFF_COPIES + (ksel - 1) * C::CB) + C::copy_base_words(ksel) + word_off;

    uint32_t one;
    asm volatile("mov.u32 %0, 1;" : "=r"(one));   // opaque 1: `one*u + t` stays an IMAD (FMA pipe), off the ALU pipe
    constexpr uint32_t scale = 1u << (KEY_MBITS + KEY_L1BITS);
    unsigned long long tbest = ~0ull;

    for (int p = 0; p < npass; p++) {
        const int r = p / a.nphase, ph = p - r * a.nphase;
        const int px = ph & 1, py = ph >> 1;
        const uint8_t* raw = smem + (p & 1) * C::CB;
        // utab[yy][m]: L1 contribution and index of vertical offset m (dy = m - R); bit 31 = leaves the plane / the range
        for (int i = tid; i < C::NBY * NM; i += C::THREADS) {
            const int y2 = i / NM, m = i - y2 * NM;
            const int oy2 = (by0 + y2) * BS, dy = m - R;
            const bool ok = (dy <= R - py) && (oy2 + dy >= 0) && (oy2 + dy + BS <= a.H - py);
            const uint32_t amvy = (uint32_t)abs(a.sc * dy + py);
            utab[y2][m] = ok ? ((amvy << KEY_MBITS) | (uint32_t)m) : 0x80000000u;
        }
        mbar_wait(&bar[p & 1], (uint32_t)((p >> 1) & 1));
        {   // byte-shifted copies 1..3 of the window (32-bit stores: the copies are skewed by base_k words)
            const uint32_t* c0 = reinterpret_cast<const uint32_t*>(raw);
            uint32_t* c1 = reinterpret_cast<uint32_t*>(smem + C::OFF_COPIES) + C::copy_base_words(1);
            uint32_t* c2 = reinterpret_cast<uint32_t*>(smem + C::OFF_COPIES + C::CB) + C::copy_base_words(2);
            uint32_t* c3 = reinterpret_cast<uint32_t*>(smem + C::OFF_COPIES + 2 * C::CB) + C::copy_base_words(3);
            constexpr int nq = (NARROW_PITCH * C::ROWS) >> 4;
            for (int g4 = tid; g4 < nq; g4 += C::THREADS) {
                const uint4 v = reinterpret_cast<const uint4*>(c0)[g4];
                const uint32_t nx = c0[4 * g4 + 4];   // first word of the next group (16 spare bytes follow the window)
                const uint32_t w[5] = {v.x, v.y, v.z, v.w, nx};
#pragma unroll
                for (int i = 0; i < 4; i++) {
                    c1[4 * g4 + i] = __funnelshift_r(w[i], w[i + 1], 8);
                    c2[4 * g4 + i] = __funnelshift_r(w[i], w[i + 1], 16);
                    c3[4 * g4 + i] = __funnelshift_r(w[i], w[i + 1], 24);
                }
            }
        }
        __syncthreads();
        if (tid == 0 && p + 1 < npass) {   // next pass's window into the other raw buffer while this one is searched
            const int r1 = (p + 1) / a.nphase, ph1 = (p + 1) - r1 * a.nphase;
            mbar_arrive_expect_tx(&bar[(p + 1) & 1], box_bytes);
            tma_load_3d(smem + ((p + 1) & 1) * C::CB, &ref_map, &bar[(p + 1) & 1], wx0, wy0, L.ref_plane[r1] + ph1);
        }
        const bool xvalid = blk_ok && (ox + dx >= 0) && (ox + dx + BS <= a.W - px) && (dx <= R - px);
        if (xvalid) {
            const uint32_t* colp = ksel ? copyp : reinterpret_cast<const uint32_t*>(raw) + word_off;
            uint32_t acc[NM];
#pragma unroll
            for (int y = 0; y < BS + NM - 1; y++) {
                uint32_t w[WPR];
#pragma unroll
                for (int j = 0; j < WPR; j++) w[j] = colp[y * WP + j];
#pragma unroll
                for (int m = 0; m < NM; m++) {
                    const int j = y - m;   // row j of the current block meets window row y for vertical offset m
                    if (j >= 0 && j < BS) {
#pragma unroll
                        for (int wi = 0; wi < WPR; wi++) acc[m] = sad4(w[wi], cur[j][wi], (j == 0 && wi == 0) ? 0u : acc[m]);
                    }
                }
            }
            const int mvx = a.sc * dx + px;
            const uint32_t tthr = (uint32_t)abs(mvx) << KEY_MBITS;
            uint32_t best = 0xFFFFFFFFu;
#pragma unroll
            for (int m = 0; m < NM; m++) {
                uint32_t u_plus_t;
                asm("mad.lo.u32 %0, %1, %2, %3;" : "=r"(u_plus_t) : "r"(one), "r"(utab[yy][m]), "r"(tthr));
                uint32_t key;
                asm("mad.lo.u32 %0, %1, %2, %3;" : "=r"(key) : "r"(acc[m]), "r"(scale), "r"(u_plus_t));
                best = min(best, key);
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
        __syncthreads();   // everyone is done with the copies (and with raw[p&1]) before they are rebuilt
    }
    if (tbest != ~0ull) atomicMin(&sbest[yy][b], tbest);
    __syncthreads();
    for (int i = tid; i < C::NBY * C::NBX; i += C::THREADS) {
        const int y2 = i / C::NBX, bb = i - y2 * C::NBX;
        if (bx0 + bb < a.bw && by0 + y2 < a.bh) {
            const unsigned long long k = sbest[y2][bb];
            const uint32_t hi = (uint32_t)(k >> 32), lo = (uint32_t)k;
            int4 o;
            o.x = (int)(lo & 1023u) - a.Rh;
            o.y = (int)((lo >> 10) & 1023u) - a.Rh;
            o.z = (int)(lo >> 20);
            o.w = (int)(hi >> 9);
            a.out[(size_t)z * a.nblk + (size_t)(by0 + y2) * a.bw + bx0 + bb] = o;
        }
    }
}

template <int BS, int R>
cudaError_t launch_narrow(const CUtensorMap& map, MeArgs a, int lanes, cudaStream_t st) {
    using C = NarrowCfg<BS, R>;
    static bool once_dev[BVC_MAX_DEVICES] = {};
    bool& once = once_dev[current_device_slot()];
    if (!once) {
        cudaError_t e = cudaFuncSetAttribute(me_narrow_kernel<BS, R>, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM);
        if (e != cudaSuccess) return e;
        e = cudaFuncSetAttribute(me_narrow_kernel<BS, R>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
        if (e != cudaSuccess) return e;
        once = true;
    }
    a.tiles_x = (a.bw + C::NBX - 1) / C::NBX;
    a.tiles_y = (a.bh + C::NBY - 1) / C::NBY;
    const long long grid = (long long)a.tiles_x * a.tiles_y * lanes;
    if (grid > 0x7fffffffLL) return cudaErrorInvalidValue;
    me_narrow_kernel<BS, R><<<(unsigned)grid, C::THREADS, C::SMEM, st>>>(map, a);
    return cudaGetLastError();
}

}  // namespace

MeTileCfg me_narrow_config(int bs, int R) {
    MeTileCfg c{};
    if (!(bs == 4 || bs == 8 || bs == 16) || R < 1 || 2 * R >= bs) return c;
    c.narrow = true;
    c.nb = 128 / bs;
    c.nby = NARROW_NBY;
    c.win_lm = (16 - R % 16) % 16;
    c.win_pitch = NARROW_PITCH;
    c.rows = NARROW_NBY * bs + 2 * R;
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
