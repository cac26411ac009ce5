// bvc_kernels.h -- host-visible launchers of the CUDA kernels (internal to libbvc_b200.so).
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "bvc_common.cuh"

namespace bvc {

// ---- K1/K3 motion estimation -----------------------------------------------------------------
struct MeLane {
    int cur_plane;                 // index of the current frame in the input pool
    int nref;                      // references available (deque length), oldest first
    int ref_plane[BVC_MAX_REFS];   // plane index in the reference pool (first phase plane when frac)
};
struct MeArgs {
    const uint8_t* cur_base;
    size_t cur_plane_bytes;
    int cur_pitch;
    const MeLane* lanes;           // device array [lanes]
    int4* out;                     // device [lanes][nblk] = (mvx, mvy, ref, sad)
    int W, H, bs, bw, bh, nblk;
    int R;                         // integer search range on the (phase) planes
    int sc;                        // 1 integer-pel, 2 half-pel MV units
    int nphase;                    // 1 or 4
    int Rh;                        // range in MV units (= R*sc)
    int win_pitch, win_copy_bytes, win_lm; // filled by the launcher
    int Rv;                        // vertical range the tiled bodies walk (>= R, 2*Rv a multiple of bs); launcher
    int tiles_x, tiles_y, n_full;  // linear tile grid: CTAs [0, n_full) own whole tiles, the rest one block row each (launcher)
    int tail_split;                // 1: cut the tiles of the last, partly filled wave into one-row CTAs
    int tall_mode;                 // tall tile shape: 0 = by launch size (launch_me_fullsearch), 1 = always, -1 = never
    int uniform_nref;              // > 0: every lane of the launch has this many references (no per-lane look-up in the kernels)
    int n_tiles;                   // narrow kernel: tiles_x * tiles_y * lanes, walked by a persistent grid (launcher)
    int key_l1bits, key_mbits;     // packed argmin key layout (launcher)
    // divisions by launch constants as multiply-high + shift (launcher): by tiles_x * tiles_y, by tiles_x, by 2R
    uint32_t perz_magic, perz_shift, tx_magic, tx_shift, r2_magic, r2_shift;
    // SAD map (FastME): when non-null the tiled kernel stores the SAD of every in-range candidate instead of reducing
    // them: uint16 [lane][ref][phase][blk][map_stride >= (2R+1)^2] (row = vertical offset + R, column = horizontal offset + R),
    // pre-filled with 0xFFFF = "leaves the plane".  max_refs = the ref stride (nRefFrames).
    uint16_t* sad_map;
    int max_refs;
    int map_stride;                // uint16 elements per (lane, ref, phase, block): (2R+1)^2 rounded up to a multiple of 8
};
struct MeTileCfg {
    bool tiled;     // me_tiled_kernel (2R >= bs)
    bool narrow;    // me_narrow_kernel (2R < bs)
    int nb;         // blocks per CTA, side by side
    int nby;        // block rows per CTA, stacked
    int win_pitch;  // TMA box width in bytes
    int rows;       // TMA box height
    int win_lm;     // left margin: window column of x0-R inside the 16-byte aligned box
    int Rv;         // vertical range walked (tiled kernel: R rounded up so that 2*Rv % bs == 0)
};
MeTileCfg me_tile_config(int bs, int R, bool tall = false);   // tall: see pick_shape (me_fullsearch.cu)
// narrow-range search (me_narrow.cu): 2R < bs, one thread per (block, candidate column), exact work
MeTileCfg me_narrow_config(int bs, int R);
cudaError_t launch_me_narrow(const CUtensorMap& ref_map, const MeArgs& args, int lanes, cudaStream_t st);
cudaError_t launch_me_fullsearch(const CUtensorMap* ref_map, const MeArgs& args, int lanes, const uint8_t* ref_base,
                                 size_t ref_plane_bytes, int ref_pitch, cudaStream_t st, const CUtensorMap* tall_map = nullptr);

// ---- K4 FastME ---------------------------------------------------------------------------------
cudaError_t launch_fastme(const MeArgs& a, int lanes, const uint8_t* ref_base, size_t ref_plane_bytes, int ref_pitch,
                          long long* cmp_out, cudaStream_t st);
// FastME on a precomputed SAD map (a.sad_map, radius a.R plane units around the block, filled by launch_me_fullsearch):
// the serial MVP chain becomes table look-ups; candidates outside the map are evaluated directly.
cudaError_t launch_fastme_walk(const MeArgs& a, int lanes, const uint8_t* ref_base, size_t ref_plane_bytes, int ref_pitch,
                               long long* cmp_out, cudaStream_t st);
// The same with transfer tables: every block's walk tabulated for all predictors within +-15 (parallel), the MVP chain as one
// look-up per block, then the walks replayed in parallel for SAD / comparison counts.  scratch: fastme_table_bytes().
size_t fastme_table_bytes(int lanes, int nblk);
cudaError_t launch_fastme_table(const MeArgs& a, int lanes, const uint8_t* ref_base, size_t ref_plane_bytes, int ref_pitch,
                                void* scratch, long long* cmp_out, cudaStream_t st);
// Without a SAD map: the serial walk with the reference windows (TMA, box fastme_window_box()) and current pixels of the
// blocks ahead staged in shared memory while the current block is walked (one CTA per frame, one warp per reference).
// Returns cudaErrorInvalidValue when the windows do not fit shared memory (fastme_window_smem() > 200 KB).
size_t fastme_window_smem(const MeArgs& a, int max_refs);
void fastme_window_box(int bs, int nphase, int* box_w, int* box_h, int* box_d);
cudaError_t launch_fastme_window(const CUtensorMap* win_map, const MeArgs& a, int lanes, int max_refs, const uint8_t* ref_base,
                                 size_t ref_plane_bytes, int ref_pitch, long long* cmp_out, cudaStream_t st);
// true when the tiled search kernel can produce the SAD map for (block size, map radius in plane units)
bool me_can_map(int bs, int R);

// ---- K2 half-pel phase planes ----------------------------------------------------------------
// src: one W x H plane; dst: 4 consecutive phase planes (P00 = copy, P10 = horizontal, P01 = vertical,
// P11 = diagonal), each W x H with the last column / row of the odd phases left 0.
cudaError_t launch_halfpel(const uint8_t* const* src_planes, uint8_t* const* dst_planes, int nplanes, int W, int H,
                           int pitch, size_t plane_bytes, cudaStream_t st);
// interleave 4 phase planes into the reference's (2H x 2W) layout (drop-in / test hook)
cudaError_t launch_halfpel_interleave(const uint8_t* phases, int W, int H, int pitch, size_t plane_bytes, uint8_t* out2x,
                                      cudaStream_t st);

// ---- K5 / K6 / K7: transform, reconstruction, intra wavefront, entropy ---------------------------
struct FrameLane {
    int cur_plane;                 // input pool index
    int slot;                      // per-frame slot of the stream arena (= frame index inside a clip call)
    int out_plane;                 // reference-pool plane receiving the reconstruction
    int nref;
    int ref_plane[BVC_MAX_REFS];
};
struct TqArgs {
    const uint8_t* cur_base;
    size_t cur_plane_bytes;
    int cur_pitch;
    uint8_t* ref_base;             // reference pool (pred is read from it, recon is written into it)
    size_t ref_plane_bytes;
    int ref_pitch;
    const FrameLane* lanes;        // device [lanes]
    const int4* mv;                // device [lanes][nblk] (P frames)
    int32_t* modes;                // device [lanes][nblk] (I frames: out)
    int32_t* isad;                 // device [lanes][nblk] (I frames: mode-decision SAD, out)
    const int32_t* qp_rows;        // device [lanes][bh]
    int16_t* levels;               // device [lanes][H][W] (frame layout), may be null
    int8_t* resid_mc;              // debug planes [lanes][H][W], may be null
    int8_t* resid_nomc;
    uint32_t* blk_bits;            // device [lanes][nblk][blk_words] per-block coefficient bit strings
    int32_t* blk_nbits;            // device [lanes][nblk]
    int blk_words;                 // words reserved per block in blk_bits
    int W, H, bs, bw, bh, nblk;
    int frac;                      // MVs in half-pel units, pred from phase planes
    int multi_ref;                 // nRefFrames > 1: pred from refs[mv.ref] else refs[0]
    uint32_t* top_mail;            // I frames: device [lanes][bh][bw][bs] mailboxes: bottom row of the block above, every pixel
                                   // as pixel | epoch << 8 (see tq_iframe_kernel)
    uint32_t epoch;                // I frames: tag of this frame's mailbox entries (1 .. 2^24-1, never the previous frame's)
    int* ticket;                   // I frames: one self-resetting counter per launch in flight: a CTA's block row follows the
                                   // order in which CTAs actually start, so a row never waits for a CTA that is not resident
    int row_begin, row_count;      // block rows to encode in this launch (0, bh = whole frame); the row-by-row
                                   // rate-control loop (Frame.get_rc_qp, Frame.py:168-188) launches one row at a time
    int quad;                      // I frames, BS >= 8: four warps per block pair (tq_iframe_quad_kernel) instead of one (2: always)
    int cta_cap;                   // P frames: at most this many CTAs, each looping over work units (0 = one CTA per unit)
    uint32_t ux_magic, ux_shift;   // P frames: division by the work units per lane / by bw (filled in by launch_tq_pframe)
    uint32_t bw_magic, bw_shift;
};
cudaError_t launch_tq_pframe(const TqArgs& a, int lanes, cudaStream_t st);
// with_entropy = false: the wavefront only (levels stay in a.levels); launch_tq_ientropy codes them later, on any stream
cudaError_t launch_tq_iframe(const TqArgs& a, int lanes, cudaStream_t st, bool with_entropy = true);
cudaError_t launch_tq_ientropy(const TqArgs& a, int lanes, cudaStream_t st);

// Block-level hook: residual (int16) + pred (int16) -> level/recon/idct/coef, nblocks blocks of bs x bs
int tq_blk_words(int bs);
cudaError_t launch_tq_blocks(const int16_t* res, const int16_t* pred, int nblocks, int bs, int qp, int16_t* level,
                             uint8_t* recon, double* idct, double* coef, cudaStream_t st);

// ---- K7b: stream assembly ----------------------------------------------------------------------
struct PackArgs {
    const int4* mv;                // P
    const int32_t* modes;          // I
    const int32_t* qp_rows;        // [lanes][bh]
    const uint32_t* blk_bits;
    const int32_t* blk_nbits;
    int blk_words;
    long long* coef_off;             // device [lanes][nblk+1] scratch: bit offset of every block's string inside its tile
    int32_t* pred_off;               // device [lanes][nblk] scratch: the same for its prediction symbols
    long long* tile_tot;             // device [lanes][tiles+1][2] scratch: (coefficient, prediction) bits per tile of 1024 blocks
    long long* tile_base;            // device [lanes][tiles+1][2] scratch: exclusive scan of tile_tot, frame totals in entry `tiles`
    int tiles;                       // pack_tiles(nblk)
    const FrameLane* lanes;        // device [lanes] (slot = where this frame's streams go)
    uint32_t* coef_stream;         // device [slots][coef_cap_words]
    uint32_t* pred_stream;         // device [slots][pred_cap_words]
    long long* frame_bits;           // device [slots][2] = (pred_bits, coef_bits) (out)
    long long* row_bits;             // device [lanes][bh] (out) bits_per_row
    size_t coef_cap_words, pred_cap_words;
    int* slot_overflow;            // device flag: a frame's stream does not fit its slot (may be null)
    int bw, bh, nblk;
    int base_qp;
    int intra;                     // 1: modes, 0: motion vectors
    int with_ref;                  // nRefFrames > 1: code the reference index difference
};
int pack_tiles(int nblk);
cudaError_t launch_pack(const PackArgs& a, int lanes, cudaStream_t st);
// bits the reference accounts to one block row (PFrame.py:76-83): coefficient bits of its blocks + its
// prediction symbols (row QP symbol included).  out: device long long[lanes].
cudaError_t launch_row_bits(const PackArgs& a, int lanes, int row, long long* out, cudaStream_t st);

// ---- rate control on the device (RCflag = 1, Frame.get_rc_qp encoder/Frame.py:168-188) ----------------
// The QP of block row k+1 follows from the bits rows <= k consumed (calculate_constant_row_bit_budget +
// find_rc_qp_for_row, encoder/RateControl/RateControl.py:9-20,34-43).  All arithmetic in IEEE double like the
// reference's Python floats: remaining -= row_bits; budget = remaining / rows_left; first table entry (QP ascending)
// whose expected row size is <= budget, else the largest QP.
struct RcArgs {
    int n;                      // table entries (0 = rate control off)
    int qp[16];                 // ascending
    long long bits[16];         // expected bits per block row at that QP (the lookup's 'I' column: Frame.py:169 always asks for 'I')
    double frame_budget;        // targetBR / frame_rate
    double* remaining;          // device [lanes]
    int32_t* qp_rows;           // device [lanes][bh]: row k+1 is written when row k is accounted
};
// start of a frame: remaining = frame_budget, QP of row 0
cudaError_t launch_rc_begin(const RcArgs& rc, int lanes, int bh, cudaStream_t st);
// bits of block row `row` (as launch_row_bits), then the QP of row + 1
cudaError_t launch_row_bits_rc(const PackArgs& a, const RcArgs& rc, int lanes, int row, long long* out, cudaStream_t st);

// ---- decoder (decode.cu) -------------------------------------------------------------------------
// One exp-Golomb bit stream inside the container image (two per frame: 2f = prediction data, 2f+1 = coefficients).
struct EgStream {
    long long byte0;     // first byte of the stream in the container
    long long nbits;     // 8 * payload bytes
    long long chunk0;    // first 512-bit chunk of this stream in the chunk tables
    long long sym0;      // first symbol of this stream in the symbol array (host fills it after the chain pass)
    int nsym, neob;      // symbols / end-of-block markers found (chain pass)
    int frame, kind;     // kind: 0 prediction data, 1 coefficients
};
struct DecArgs {
    uint8_t* ref_base;             // reference pool: predictions are read from it, decoded frames written into it
    size_t ref_plane_bytes;
    int ref_pitch;
    const FrameLane* lanes;        // device [lanes]; slot = frame index in the clip
    const int4* mv_all;            // [nframes][nblk]
    const int32_t* modes_all;      // [nframes][nblk]
    const int32_t* qp_all;         // [nframes][bh]
    const int16_t* syms;           // all symbols of the clip
    const long long* coef_sym0;    // [nframes] first symbol of the frame's coefficient stream
    const int* blk_start;          // [nframes][nblk+1] first symbol of every block, relative to coef_sym0
    int16_t* levels_out;           // [nframes][H][W] or null
    const uint8_t* frame_ok;       // [nframes] 0 = the frame's coefficient stream is malformed: leave the frame alone
    uint32_t* top_mail;            // I frames: [lanes][bh][bw][bs] bottom rows handed down, pixel | epoch << 8 (see tq_iframe_kernel)
    uint32_t epoch;
    int* ticket;                   // I frames: start-order ticket counter (see TqArgs::ticket)
    int* err_flag;
    int W, H, bs, bw, bh, nblk;
    int frac;
};
int eg_chunk_bits();
// Tokenizing runs step by step (the streams of the frames one decode step needs): chunks [chunk_begin, chunk_begin + nchunks)
// of the chunk tables, streams stream_list[0 .. nstreams)
cudaError_t launch_eg_tokenize_spec(const uint8_t* data, EgStream* streams, const int* stream_list, int nstreams, const int* chunk_stream,
                                    long long chunk_begin, long long nchunks, uint8_t* exit_tab, uint16_t* nsym_tab, uint8_t* neob_tab,
                                    uint8_t* entry_tab, int* symbase, int* eobbase, int* err_flag, cudaStream_t st);
cudaError_t launch_eg_chunk_map(const EgStream* streams, int nstreams, int* chunk_stream, cudaStream_t st);
cudaError_t launch_eg_offsets(EgStream* streams, const int* stream_list, int nstreams, long long slab_base, long long* coef_sym0,
                              uint8_t* frame_ok, int nblk, int pred_only, int* err_flag, cudaStream_t st);
cudaError_t launch_eg_tokenize_emit(const uint8_t* data, const EgStream* streams, const int* chunk_stream, long long chunk_begin,
                                    long long nchunks, const uint8_t* entry_tab, const int* symbase, const int* eobbase, int16_t* syms,
                                    int* blk_start, int nblk, cudaStream_t st);
cudaError_t launch_pred_decode(const EgStream* streams, const int16_t* syms, const uint8_t* intra_flags, const int* frame_list, int nframes,
                               int4* mv_all, int32_t* modes_all, int32_t* qp_all, int bw, int bh, int base_qp, int with_ref, int* err_flag,
                               cudaStream_t st);
cudaError_t launch_dec_pframe(const DecArgs& a, int lanes, cudaStream_t st);
cudaError_t launch_dec_iframe(const DecArgs& a, int lanes, cudaStream_t st);
cudaError_t launch_fill_plane(uint8_t* p, size_t n, uint8_t v, cudaStream_t st);

// ---- container assembly ---------------------------------------------------------------------------
struct ContainerArgs {
    const long long* frame_bits;   // device [nframes][2]
    const uint32_t* coef_stream;   // device [nframes][coef_cap_words]
    const uint32_t* pred_stream;
    size_t coef_cap_words, pred_cap_words;
    long long* frame_off;          // device [nframes+1] byte offset of every frame record (out), total at the end
    int* overflow;                 // device flag: a payload does not fit its length field
    uint8_t* out;                  // device container image
    long long out_cap;
    int nframes, i_period;
};
cudaError_t launch_container(const ContainerArgs& a, cudaStream_t st);

}  // namespace bvc
