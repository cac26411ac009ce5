/*
 * bvc_oracle.h -- CPU restatement (plain C) of the encoder hot path of dheri/basic_video_codec.
 *
 * TEST INFRASTRUCTURE ONLY.  This library is the parity checker for the CUDA product in
 * basic_video_codec_b200/csrc.  Only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs may load it.  The product never links or calls it.
 *
 * Parity status: PINNED for everything except the DCT rounding noise.  Integer work (SAD search,
 * tie-break, FastME walk, half-pel buffer, intra predictors/mode rule, zig-zag/RLE/exp-Golomb,
 * container) is checked bit-for-bit against the imported Python reference
 * (oracle/ref_harness.py, tests/golden/).  The 2-D DCT is the *defined* fp64 transform of DESIGN.md
 * (the reference delegates to scipy.fftpack, a third-party FFT: encoder/dct.py:9-18); the Python
 * reference is run with this transform monkeypatched in to produce the bit-exact goldens, and
 * against scipy-fp64 the coefficients agree to < 1e-9 (tests/test_oracle_vs_reference.py).
 *
 * All `file:line` citations are relative to the reference repository root.
 */
#ifndef BVC_ORACLE_H
#define BVC_ORACLE_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct bvo_config {
    int width, height;   /* padded plane size (multiples of block) */
    int block;           /* i  : 2,4,8,16,32 */
    int range;           /* r  : integer-pel search range (ignored when fastme) */
    int qp;              /* base quantisation parameter */
    int nref;            /* nRefFrames (deque maxlen), >= 1 */
    int fastme;          /* 0/1 */
    int frac;            /* fracMeEnabled 0/1: MVs in half-pel units */
    int i_period;        /* I_Period */
} bvo_config;

/* ---- pieces (block_predictor.py, dct.py, entropy_encoder.py) ---- */

/* build_pre_interpolated_buffer, block_predictor.py:145-177.  out is (2H x 2W). */
void bvo_halfpel_plane(const uint8_t *ref, int W, int H, uint8_t *out);

/* find_lowest_mae_block, block_predictor.py:61-91 for one block at (ox,oy).
 * planes[k]: integer plane (H x W) when !frac, half-pel plane (2H x 2W) when frac.
 * Returns SAD of the winner; mv[3] = {mvx, mvy, ref}.  valid_out (optional) = number of
 * in-range candidates actually evaluated. */
int32_t bvo_full_search_block(const uint8_t *cur, int W, int H, int ox, int oy, int bs,
                              const uint8_t *const *planes, int nref, int range, int frac,
                              int32_t mv[3], int64_t *valid_out);

/* find_fast_me_block, block_predictor.py:11-58 (incl. the late-binding closure behaviour). */
int32_t bvo_fast_me_block(const uint8_t *cur, int W, int H, int ox, int oy, int bs,
                          const uint8_t *const *planes, int nref, int frac,
                          int mvp_x, int mvp_y, int32_t mv[3], int64_t *comparisons);

/* frame-level ME (all blocks, raster order; FastME carries the MVP chain, PFrame.py:105-110).
 * mv: nblk*3 int32, sad: nblk int32. Returns total_mae_comparisons (PFrame.py:68). */
int64_t bvo_me_frame(const bvo_config *cfg, const uint8_t *cur, const uint8_t *const *planes,
                     int nref_avail, int32_t *mv, int32_t *sad);

/* generate_quantization_matrix, dct.py:21-32: returns the shift s with Q = 2^s. */
int bvo_q_shift(int bs, int qp, int u, int v);

/* The defined fp64 forward transform: coef (bs*bs doubles, row-major [u][v]) of an int16 residual. */
void bvo_fdct(const int16_t *res, int bs, double *coef);
/* The defined fp64 inverse transform of rescaled levels. */
void bvo_idct(const double *coef, int bs, double *out);
/* direct access to the constant tables (for the table-agreement test) */
const double *bvo_dct_ct(int bs);
const double *bvo_dct_w(int bs);   /* bs*bs scale table W[u][v] */

/* apply_dct_and_quantization + reconstruct_block, Frame.py:190-202, for one block.
 * res/pred: bs*bs (row-major, stride bs).  Outputs: level (int16), recon (uint8), idct (double),
 * coef (double, pre-quantisation; may be NULL). */
void bvo_transform_block(const int16_t *res, const int16_t *pred, int bs, int qp,
                         int16_t *level, uint8_t *recon, double *idct, double *coef);

/* ---- bit writer (bitarray semantics: MSB first, zero padded to bytes) ---- */
typedef struct bvo_bits {
    uint8_t *data;
    size_t   nbits;
    size_t   cap_bytes;
} bvo_bits;
void bvo_bits_init(bvo_bits *b);
void bvo_bits_free(bvo_bits *b);
void bvo_put_eg(bvo_bits *b, int32_t v);                 /* entropy_encoder.py:8-29 */
int  bvo_eg_len(int32_t v);
/* zigzag_order entropy_encoder.py:115-135; out has bs*bs entries */
void bvo_zigzag(const int16_t *blk, int stride, int bs, int16_t *out);
/* rle_encode entropy_encoder.py:65-88; out needs 2*n+1 entries; returns count */
int  bvo_rle(const int16_t *zz, int n, int32_t *out);

/* ---- frame level ---- */
typedef struct bvo_frame_out {
    /* all caller-allocated, H*W (or nblk) sized */
    uint8_t  *recon;          /* reconstructed_frame */
    int16_t  *levels;         /* quantized_dct_residual_frame (block tiled in frame layout) */
    int32_t  *mv;             /* P: nblk*3 ; I: unused */
    int32_t  *sad;            /* P: nblk min SAD ; I: nblk mode-decision "SAD" (with the uint8 wrap) */
    int32_t  *modes;          /* I: nblk intra modes */
    int8_t   *resid_mc;       /* P: int8 cast of idct residual (PFrame.py:39,63); I: uint8 residual. may be NULL */
    int8_t   *resid_nomc;     /* P: cur - refs[0] as int8 (PFrame.py:40,64,116). may be NULL */
    bvo_bits  pred_bits;      /* entropy_encoded_prediction_data */
    bvo_bits  coef_bits;      /* entropy_encoded_DCT_coffs */
    int64_t  *bits_per_row;   /* rows entries (may be NULL) */
    double    avg_mae;
    int64_t   mae_comparisons;
    /* optional per-row QP feedback (rate control, Frame.get_rc_qp Frame.py:168-188): called before block row
     * `row` is coded with the bits the previous row consumed (0 for row 0); returns the row's QP.
     * qp_used (optional, rows entries) receives the QPs. */
    int32_t (*qp_cb)(void *user, int32_t row, int64_t prev_row_bits);
    void     *qp_user;
    int32_t  *qp_used;
} bvo_frame_out;

/* PFrame.encode_mc_q_dct, PFrame.py:29-97.  refs: deque order (index 0 = oldest).
 * hp_refs: half-pel planes (only read when cfg->frac).  qp_rows: per block-row QP (rows entries). */
void bvo_encode_pframe(const bvo_config *cfg, const uint8_t *cur, const uint8_t *const *refs,
                       const uint8_t *const *hp_refs, int nref_avail, const int32_t *qp_rows,
                       bvo_frame_out *out);
/* IFrame.encode_mc_q_dct, IFrame.py:22-83 */
void bvo_encode_iframe(const bvo_config *cfg, const uint8_t *cur, const int32_t *qp_rows,
                       bvo_frame_out *out);

/* encode_video frame loop + container (encoder.py:75-121,154-155,174-186), RCflag = 0.
 * frames: nframes planes of cfg->width*cfg->height (already padded).  first_index is the 1-based
 * index of frames[0] in the clip (the I/P decision is (idx-1) % I_Period == 0); use 1 for a clip.
 * Writes the container bytes into a malloc'd buffer (*out, *out_len); recon_out (optional)
 * receives nframes reconstructed planes.  nthreads > 1 encodes independent GOPs concurrently
 * (OpenMP); the byte stream is identical.  Returns 0 on success. */
int bvo_encode_clip(const bvo_config *cfg, const uint8_t *frames, int nframes, int first_index,
                    uint8_t **out, size_t *out_len, uint8_t *recon_out, int nthreads);
/* decode_video (decoder.py:26-87): parses the container, entropy-decodes both streams of every frame
 * (Frame.py:81-110, PFrame.py:166-228, IFrame.py:132-166) and rebuilds the frames (PFrame.py:252-317,
 * IFrame.py:85-114).  frames_out: max_frames planes.  Optional per-frame outputs: levels_out (H*W int16),
 * pred_out (nblk*3: mvx,mvy,ref or mode,0,0), qp_out (rows), kinds_out (1 = I).  Returns 0, or -1 for a
 * malformed stream (ValueError / IndexError in the reference). */
int bvo_decode_clip(const bvo_config *cfg, const uint8_t *data, size_t len, int max_frames, uint8_t *frames_out,
                    int *nframes_out, int16_t *levels_out, int32_t *pred_out, int32_t *qp_out, uint8_t *kinds_out);
void bvo_free(void *p);

#ifdef __cplusplus
}
#endif
#endif
