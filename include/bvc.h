/*
 * bvc.h -- C ABI of libbvc_b200.so: the B200-native encoder hot path of dheri/basic_video_codec.
 *
 * The reference is pure Python and has no FFI; its seam is the object protocol between
 * encoder/encoder.py (the frame loop) and the Frame classes.  Every entry point below names the
 * reference interface it replaces (paths relative to the reference repository).  A maintainer binds
 * these with ctypes (INTEGRATION.md shows the stub); basic_video_codec_b200/ ships that binding plus
 * PFrame / IFrame / encode_video mirrors with the reference's names and attributes.
 *
 * Conventions: plain pointers and sizes, no exceptions across the boundary, int status
 * (0 = BVC_OK, negative = error) and bvc_last_error() for the message.  All planes are row-major
 * uint8 luma (H x W, W bytes per row).  A context is bound to one GPU and one geometry and is not
 * thread safe; contexts are independent (one per GPU / host thread for GOP sharding).
 * There is no CPU fallback: every function fails with BVC_ERR_CUDA when no sm_100 device is usable.
 */
#ifndef BVC_H
#define BVC_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define BVC_OK 0
#define BVC_ERR_INVALID (-1)     /* ValueError in the reference (params.py:28-36, block_predictor.py:70-71) */
#define BVC_ERR_CUDA (-2)
#define BVC_ERR_OVERFLOW (-3)    /* OverflowError: payload does not fit the container length field (encoder.py:108-117) */
#define BVC_ERR_NOMEM (-4)
#define BVC_ERR_UNSUPPORTED (-5)

typedef struct bvc_ctx bvc_ctx;

/* EncoderConfig (encoder/params.py:6-23) + InputParameters.width/height (input_parameters.py:4-11).
 * RCflag: per-row QPs are passed explicitly to the frame-level / row-level calls; bvc_set_rate_control puts RCflag 1 on the clip path. */
typedef struct bvc_params {
    int width, height;   /* luma size; must be multiples of block_size (pad_frame, common.py:22-32, is applied by the Python layer) */
    int block_size;      /* i  : 4, 8 or 16 */
    int search_range;    /* r  : integer-pel range; ignored when fast_me */
    int qp;              /* quantization_factor (base QP) */
    int nref_frames;     /* nRefFrames: length of the reference window, 1..8 */
    int fast_me;         /* fastME */
    int frac_me;         /* fracMeEnabled: half-pel motion vectors */
    int i_period;        /* I_Period */
} bvc_params;

/* Outputs of one encoded frame: the attributes encoder/encoder.py reads off a Frame after
 * encode_mc_q_dct() (encoder.py:89-155).  All pointers are caller-allocated host buffers; any of
 * them may be NULL to skip that output (and its device-to-host copy). */
typedef struct bvc_frame_out {
    uint8_t *recon;          /* reconstructed_frame, H*W */
    int16_t *levels;         /* quantized_dct_residual_frame, H*W, block-tiled in frame layout */
    int32_t *mv;             /* P: mv_field in raster order, nblk*3 (mvx, mvy, ref_idx) */
    int32_t *sad;            /* per-block min SAD (= MAE * i*i); I: mode-decision SAD */
    int32_t *modes;          /* I: intra_modes, nblk */
    int8_t  *resid_mc;       /* residual_frame (debug plane: PFrame.py:39,63 / IFrame.py:30,57) */
    int8_t  *resid_nomc;     /* residual_wo_mc_frame (P only, PFrame.py:40,64) */
    uint8_t *pred_bytes;     /* entropy_encoded_prediction_data.tobytes() */
    size_t   pred_cap;
    uint8_t *coef_bytes;     /* entropy_encoded_DCT_coffs.tobytes() */
    size_t   coef_cap;
    int64_t *bits_per_row;   /* bits_per_row, H/i entries */
    /* scalars written by the call */
    int64_t  pred_nbits, coef_nbits;
    double   avg_mae;
    int64_t  mae_comparisons; /* total_mae_comparisons */
} bvc_frame_out;

/* lifetime ---------------------------------------------------------------------------------- */
/* max_lanes: how many independent GOPs the context may encode in lock-step (clip API); 1 is enough
 * for the frame-level calls. */
int bvc_create(bvc_ctx **ctx, int device, const bvc_params *params, int max_lanes);
void bvc_destroy(bvc_ctx *ctx);
const char *bvc_last_error(const bvc_ctx *ctx); /* ctx may be NULL: error of the last failed bvc_create */
int bvc_set_qp(bvc_ctx *ctx, int qp);

/* frame level -------------------------------------------------------------------------------- */
/* IFrame.encode_mc_q_dct (encoder/IFrame.py:22-83).  qp_rows: H/i per-row QPs or NULL (= base QP). */
int bvc_encode_iframe(bvc_ctx *ctx, const uint8_t *cur, const int32_t *qp_rows, bvc_frame_out *out);
/* PFrame.encode_mc_q_dct (encoder/PFrame.py:29-97).  refs: the reference window in deque order,
 * index 0 = oldest (encoder.py:33,154); nref_avail = len(reference_frames). */
int bvc_encode_pframe(bvc_ctx *ctx, const uint8_t *cur, const uint8_t *const *refs, int nref_avail,
                      const int32_t *qp_rows, bvc_frame_out *out);

/* row level: the rate-control feedback loop (RCflag = 1) ------------------------------------------- */
/* Frame.get_rc_qp (encoder/Frame.py:168-188) derives the QP of block row k from the bits rows < k
 * consumed (encoder/PFrame.py:53-83, encoder/IFrame.py:38-70).  bvc_frame_begin uploads the frame and
 * its reference window and runs motion estimation once (it does not depend on the QP);
 * bvc_frame_encode_row transforms / reconstructs / entropy-codes block row `row` with `qp` and returns
 * the bits that row added to both streams (row_bits_consumed); rows must be given in order 0..H/i-1;
 * bvc_frame_end assembles the streams and fills `out` exactly like bvc_encode_pframe / _iframe. */
int bvc_frame_begin(bvc_ctx *ctx, const uint8_t *cur, const uint8_t *const *refs, int nref_avail, int intra);
int bvc_frame_encode_row(bvc_ctx *ctx, int row, int qp, int64_t *row_bits);
int bvc_frame_end(bvc_ctx *ctx, bvc_frame_out *out);

/* block-level hooks ---------------------------------------------------------------------------- */
/* PFrame.get_motion_vector over a whole frame: find_lowest_mae_block or find_fast_me_block per the
 * context's parameters (encoder/block_predictor.py:11-91).  mv: nblk*3, sad: nblk. */
int bvc_me_search(bvc_ctx *ctx, const uint8_t *cur, const uint8_t *const *refs, int nref_avail,
                  int32_t *mv, int32_t *sad, int64_t *comparisons);
/* build_pre_interpolated_buffer (encoder/block_predictor.py:145-177): out is (2H x 2W). */
int bvc_interp_halfpel(bvc_ctx *ctx, const uint8_t *ref, uint8_t *out2x);
/* apply_dct_and_quantization + reconstruct_block (encoder/Frame.py:190-202) on nblocks dense
 * bs x bs blocks.  idct / coef (fp64) may be NULL. */
int bvc_dct_quant_recon(int device, const int16_t *residual, const int16_t *pred, int nblocks, int bs,
                        int qp, int16_t *level, uint8_t *recon, double *idct, double *coef);

/* clip level ------------------------------------------------------------------------------------ */
/* The encode_video frame loop for RCflag = 0 (encoder/encoder.py:75-121,154-155,174-186): frame idx
 * (1-based) with (idx-1) % I_Period == 0 is an I frame and clears the reference window, so GOPs are
 * independent and are encoded max_lanes at a time.  `frames` = nframes planes (host memory, ideally
 * pinned).  The container bytes (encoded.bin layout, encoder.py:104-121) are written to out[0..*out_len).
 * recon (optional) receives the nframes reconstructed planes.  When out_cap is too small the call returns BVC_ERR_NOMEM
 * with *out_len = the bytes needed, so the caller can retry. */
int bvc_encode_clip(bvc_ctx *ctx, const uint8_t *frames, int nframes, uint8_t *out, size_t out_cap,
                    size_t *out_len, uint8_t *recon);
/* Same, for a clip whose planes are already resident in HBM (bvc_clip_upload), so the timed region
 * of a throughput measurement holds no input transfer. */
int bvc_clip_upload(bvc_ctx *ctx, const uint8_t *frames, int nframes);
int bvc_encode_clip_resident(bvc_ctx *ctx, int nframes, uint8_t *out, size_t out_cap, size_t *out_len,
                             uint8_t *recon);
/* Rate control on the clip path: RCflag = 1 (Frame.get_rc_qp encoder/Frame.py:168-188 with
 * calculate_constant_row_bit_budget / find_rc_qp_for_row, encoder/RateControl/RateControl.py:9-20,34-43).  After this call
 * bvc_encode_clip* encode every frame block row by block row; the launch that accounts a row's bits picks the next row's
 * QP, for all GOP lanes at once, without a host round trip.  frame_bit_budget = targetBR / frame_rate (encoder.py:181-185);
 * the table is the lookup's 'I' column (Frame.py:169 always asks for 'I'): n <= 16 entries, QPs ascending, expected bits
 * per block row.  rc_flag 0 switches back to the base QP; RCflag 2 / 3 (two passes, scene changes: consecutive frames and
 * GOPs are coupled) stay on the frame-level calls and return BVC_ERR_UNSUPPORTED here. */
int bvc_set_rate_control(bvc_ctx *ctx, int rc_flag, double frame_bit_budget, int n, const int32_t *qps, const int64_t *row_bits);

/* Sharded jobs (GOPs of one clip on several GPUs, encoder.py:174-186 makes them independent): the same
 * encode, but the finished container stays in device memory and only its length comes back; bvc_container_download then
 * copies bytes [offset, offset+len) of it to `dst` -- typically straight into this rank's slice of a host buffer shared by
 * all ranks, once the ranks have exchanged their lengths, so the concatenated stream is written exactly once.
 * frames == NULL encodes the resident clip (bvc_clip_upload).  cap_hint: device bytes to reserve for the container
 * (0 = nframes * W * H / 2 + 1 MB); BVC_ERR_NOMEM with *out_len = the bytes needed when it does not fit. */
int bvc_encode_clip_device(bvc_ctx *ctx, const uint8_t *frames, int nframes, size_t cap_hint, size_t *out_len);
int bvc_container_download(bvc_ctx *ctx, uint8_t *dst, size_t offset, size_t len);
/* Page-lock (cudaHostRegister) a host range the caller owns -- e.g. a POSIX shared-memory mapping -- so that uploads
 * from it / downloads into it run at the link rate; bvc_host_unregister before unmapping. */
int bvc_host_register(void *ptr, size_t bytes);
int bvc_host_unregister(void *ptr);

/* Device memory of the clip calls does not grow with the clip: inputs pass through a ring of 6 steps x max_lanes planes, bit
 * streams through one wave's (max_lanes x I_Period frames) slots.  A slot reserves 6 bits per pixel for a frame's
 * coefficient stream by default (at least 1 MB, at most the worst case of 26 bits per pixel); a frame that needs more
 * makes the call fail with BVC_ERR_NOMEM -- raise the reservation here (0 = default, SIZE_MAX = worst case) and repeat. */
int bvc_set_stream_slot_bytes(bvc_ctx *ctx, size_t bytes);

/* Input stage: like bvc_clip_upload, for an I420 (YUV 4:2:0 planar) file image of src_w x src_h frames.  Only the luma
 * planes are transferred (read_y_component, assign1/ex2.py:14-28) and they are padded bottom / right with 128 to the
 * context's width / height (pad_frame, common.py:22-32), which must be src_w / src_h rounded up to block_size. */
int bvc_clip_upload_i420(bvc_ctx *ctx, const uint8_t *yuv, int src_w, int src_h, int nframes);

/* decoder ------------------------------------------------------------------------------------------ */
/* decode_video (decoder.py:26-87): parses the container (`data`, encoded.bin layout), entropy-decodes both streams
 * of every frame (Frame.entropy_decode_dct_coffs encoder/Frame.py:81-110, PFrame / IFrame.entropy_decode_prediction_data
 * encoder/PFrame.py:166-228, encoder/IFrame.py:132-166) and rebuilds the frames (construct_frame_from_dct_and_mv
 * encoder/PFrame.py:252-317, IFrame.decode_mc_q_dct encoder/IFrame.py:85-114).  At most max_frames frames
 * (frames_to_process) are decoded; GOPs (I-frame boundaries) are decoded max_lanes at a time.  The context supplies
 * width/height/block_size, the base qp, nref_frames and frac_me exactly like params.encoder_config does.
 * frames_out: max_frames planes (H*W each) or NULL.  Optional per-frame outputs, any may be NULL: levels_out (H*W int16,
 * quantized_dct_residual_frame), pred_out (nblk*3 int32: mvx,mvy,ref for P frames / mode,0,0 for I frames),
 * qp_rows_out (H/i int32: rc_qp_per_row), kinds_out (1 = I frame).  A malformed stream (ValueError / IndexError in
 * the reference) returns BVC_ERR_INVALID. */
int bvc_decode_clip(bvc_ctx *ctx, const uint8_t *data, size_t len, int max_frames, uint8_t *frames_out, int *nframes_out,
                    int16_t *levels_out, int32_t *pred_out, int32_t *qp_rows_out, uint8_t *kinds_out);
/* One frame: Frame.entropy_decode_prediction_data + entropy_decode_dct_coffs + decode_mc_q_dct on the payloads of one
 * container record.  refs: the reference window in deque order (index 0 = oldest), ignored for an I frame.
 * coef == NULL: prediction data only (entropy_decode_prediction_data on its own): fills pred_out / qp_rows_out. */
int bvc_decode_frame(bvc_ctx *ctx, int intra, const uint8_t *pred, size_t pred_len, const uint8_t *coef, size_t coef_len,
                     const uint8_t *const *refs, int nref_avail, uint8_t *recon, int16_t *levels, int32_t *pred_out,
                     int32_t *qp_rows_out);

/* Lane groups of the clip path: the max_lanes GOP lanes of a step are split into `groups` (1..4) groups with
 * their own CUDA streams, so that the tail of one group's motion search is filled by the other groups' kernels.
 * 1 = every kernel of a step back to back on one stream (the per-kernel timings of bvc_last_kernel_times are
 * exclusive only then).  Default 2 (environment override: BVC_LANE_GROUPS).  The output does not depend on it. */
int bvc_set_lane_groups(bvc_ctx *ctx, int groups);

/* FastME evaluation (find_fast_me_block, encoder/block_predictor.py:11-58, and the serial predictor chain of
 * encoder/PFrame.py:34,44,105-110).  The output does not depend on the setting.
 *   0 (default) automatic: 3 when at least 26 frames are in flight per step (16x16 blocks), else 4.
 *   1 every candidate is evaluated from global memory when the serial walk reaches it (first version).
 *   2 SAD map (all candidates within 16 MV units of every block, computed by the tiled full-search kernel), serial walk.
 *   3 window walk: serial walk, the reference windows of the blocks ahead staged in shared memory by TMA; no scratch memory.
 *   4 SAD map + transfer tables: the walk of every block tabulated for all predictors within +-15 in parallel, the chain
 *     is one look-up per block.  Scratch: 2 * 33^2 bytes per block and reference for the map, 2.2 KB per block for the tables. */
int bvc_set_fastme_direct(bvc_ctx *ctx, int on);

/* instrumentation --------------------------------------------------------------------------- */
/* kernels launched by this context since creation (bench.py's gpu_launches) */
int64_t bvc_launch_count(const bvc_ctx *ctx);
/* Device time of the last clip call, from CUDA events recorded on the context's compute stream around
 * every kernel launch: ms[k] / launches[k] per kernel class (BVC_K_*), and the whole call (first
 * enqueue to last download) in *clip_ms.  The per-class figures are recorded only with one lane group
 * (bvc_set_lane_groups(ctx, 1)), where kernels run back to back on one stream; otherwise they are 0.
 * Any pointer may be NULL. */
#define BVC_K_ME 0       /* motion estimation (full search or FastME) */
#define BVC_K_TQ_P 1     /* P-frame residual/transform/quantise/reconstruct + per-block entropy */
#define BVC_K_TQ_I 2     /* I-frame intra wavefront (same body) */
#define BVC_K_PACK 3     /* stream assembly (scan + emit) */
#define BVC_K_HALFPEL 4  /* half-pel phase planes */
#define BVC_NUM_KERNEL_CLASSES 5
int bvc_last_kernel_times(const bvc_ctx *ctx, double *ms, int64_t *launches, double *clip_ms);
/* algorithmic pixel-absdiffs (valid candidates * i*i * refs) of one P frame with `nref_avail` references */
int64_t bvc_me_work_per_frame(const bvc_ctx *ctx, int nref_avail);
/* Issue-rate ceilings of the device, measured now (dependent-free chains on every SM, best of five launches, CUDA
 * events): VABSDIFF4.U8.ACC thread-ops/s (x4 = pixel-absdiffs/s, the search kernels' roofline) and DFMA thread-ops/s
 * (the transform's).  Any pointer may be NULL. */
int bvc_measure_peaks(int device, double *vabsdiff4_thread_ops_per_s, double *dfma_thread_ops_per_s, int *sm_count);

#ifdef __cplusplus
}
#endif
#endif
