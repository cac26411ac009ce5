// bvc_api.cu -- the C ABI (include/bvc.h): context, device memory layout, per-step kernel sequencing,
// GOP-lane scheduling and host-side container assembly.
//
// Device layout (one context = one GPU, one geometry):
//   in_pool   [nframes][H][pitch]            input luma planes (whole clip resident; 1.25 GB for 600 x 1080p)
//   ref_pool  [lanes][nref+1][pps][H][pitch] reconstruction ring per GOP lane; pps = 4 phase planes when
//                                            half-pel ME is on (phase 0 = the reconstruction itself), else 1.
//                                            One 3-D TMA tensor map (x, y, plane) covers the whole pool.
//   per lane  mv int4[nblk] | modes | levels (frame API only) | blk_bits[nblk][blk_words] | blk_nbits |
//             coef/pred streams (double buffered so the D2H of step s overlaps step s+1)
// A "step" encodes frame k of every active GOP lane with one launch per kernel:
//   I step: tq_iframe (wavefront)            -> pack_scan -> pack_emit [-> halfpel]
//   P step: me (full search | fastme) -> tq_pframe -> pack_scan -> pack_emit [-> halfpel]
#include <cuda.h>
#include <cuda_runtime.h>

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/bvc.h"
#include "bvc_kernels.h"

#define BVC_MAX_GROUPS 4

namespace bvc {
cudaError_t launch_fastme(const MeArgs& a, int lanes, const uint8_t* ref_base, size_t ref_plane_bytes, int ref_pitch,
                          long long* cmp_out, cudaStream_t st);
}

using namespace bvc;

static std::string g_create_error;

struct bvc_ctx {
    int device = 0;
    bvc_params p{};
    Geom g{};
    int max_lanes = 1;
    int pps = 1;          // planes per ring slot
    int slots = 2;        // ring slots per lane (nref + 1)
    cudaStream_t st = nullptr, st_h2d = nullptr, st_d2h = nullptr;   // compute / frame-level, input upload, decoded-plane download
    // clip path: the GOP lanes of a step are split into `ngroups` lane groups with their own streams, so the tail
    // of one group's motion search (a last, partly filled wave of 77 us CTAs) is filled by the other group's
    // kernels.  st_grp: motion search (low priority), st_post: everything after it (high priority).  B200's block
    // scheduler does not place a second kernel's CTAs next to a kernel that still has CTAs to dispatch
    // (profiles/microbench/cosched.cu), so this buys the tails (~1.5 %), not a transform-under-search overlap.
    int ngroups = 2;
    int iquad = 1;        // I-frame wavefront with four warps per block pair (BVC_IQUAD=0: the one-warp kernel)
    int tq_cta_cap = 0;   // clip path with lane groups: CTAs of the P transform launch (0 = one per work unit), BVC_TQ_CTAS
    int me_tall = 0;      // motion search: tall tile shape by launch size (0), always (BVC_ME_TALL=1), never (BVC_ME_TALL=-1)
    int tail_split = 1;   // motion search: tiles of the last, partly filled wave as one-row CTAs (BVC_TAIL_SPLIT=0 turns it off)
    cudaStream_t st_grp[BVC_MAX_GROUPS] = {}, st_post[BVC_MAX_GROUPS] = {};
    // st_pack: entropy coding of I levels + stream assembly.  Nothing of the next step's search needs them (it needs the
    // reconstruction only), so they leave the per-group critical path ME(k) -> transform(k) -> ME(k+1) and run while the
    // next search is on the GPU (high priority: they take the slots the search frees).  The motion-vector array is double
    // buffered by step parity so that ME(k+1) does not overwrite what the assembly of step k still reads.
    cudaStream_t st_pack[BVC_MAX_GROUPS] = {};
    cudaEvent_t ev_me[BVC_MAX_GROUPS] = {}, ev_post[BVC_MAX_GROUPS] = {}, ev_tq[BVC_MAX_GROUPS] = {}, ev_pack[BVC_MAX_GROUPS] = {};
    std::string err;
    int64_t launches = 0;

    uint8_t* in_pool = nullptr;
    size_t in_planes = 0;
    uint8_t* ref_pool = nullptr;
    size_t ref_planes = 0;
    CUtensorMap ref_map{};
    bool have_map = false;
    CUtensorMap ref_map_tall{};   // box of the tall tile shape (me_tile_config(.., true)), when the geometry has one
    bool have_tall = false;
    CUtensorMap fw_map{};     // FastME window walk: box = fastme_window_box()
    bool have_fw_map = false;

    int4* d_mv = nullptr;
    int32_t *d_modes = nullptr, *d_isad = nullptr, *d_qp_rows = nullptr, *d_blk_nbits = nullptr;
    int16_t* d_levels = nullptr;       // [lanes][H][W]: I frames (all lanes) and the frame-level calls (lane 0)
    int8_t *d_resid_mc = nullptr, *d_resid_nomc = nullptr;
    uint32_t* d_blk_bits = nullptr;
    int blk_words = 0;
    long long *d_coef_off = nullptr, *d_row_bits = nullptr, *d_cmp = nullptr;
    int32_t* d_pred_off = nullptr;                                // stream assembly scratch, see PackArgs
    long long *d_tile_tot = nullptr, *d_tile_base = nullptr;
    int pack_tiles_n = 0;
    // stream arena: one slot per frame of a clip call (slot 0 for the frame-level calls)
    uint32_t *d_coef_stream = nullptr, *d_pred_stream = nullptr;
    long long *d_frame_bits = nullptr, *d_frame_off = nullptr;
    int* d_overflow = nullptr;
    size_t stream_slots = 0;
    size_t coef_cap_words = 0, pred_cap_words = 0;
    uint8_t* d_container = nullptr;   // whole-clip container kept on the device (bvc_encode_clip_device)
    size_t container_cap = 0;
    uint8_t* d_frag[2] = {nullptr, nullptr};   // staging of one wave's container fragment
    size_t frag_cap = 0, frag_min = 0;
    int nfrag = 0;
    void* h_totals = nullptr;         // pinned: per-wave fragment sizes and overflow flags
    size_t h_totals_cap = 0;
    size_t slot_bytes = 0;            // bvc_set_stream_slot_bytes: 0 = default
    uint32_t* d_top_mail = nullptr;   // [max_lanes][bh][bw][bs]: bottom rows handed down the I-frame wavefront (tq_iframe_kernel)
    uint32_t epoch = 0;               // tag of the last I frame's mailbox entries
    int* d_ticket = nullptr;      // [max_lanes]: start-order counters of the wavefront kernels, one per lane group in flight
    MeLane* d_me_lanes = nullptr;
    FrameLane* d_fr_lanes = nullptr;
    size_t lane_desc_cap = 0;
    const uint8_t** d_hp_src = nullptr;
    uint8_t** d_hp_dst = nullptr;
    size_t hp_desc_cap = 0;

    // decoder scratch (grow-only device buffers, see dbuf())
    struct DBuf { void* p = nullptr; size_t cap = 0; };
    DBuf sad_map;         // FastME look-up table (uint16 [lanes][nref][phase][blk][n1*n1])
    DBuf fastme_tab;      // FastME transfer tables + per-block predictors (fastme_table_bytes)
    int fastme_direct = 0;  // bvc_set_fastme_direct: 0 auto (window walk / transfer tables), 1 direct evaluation, 2 SAD map + serial walk,
                            // 3 window walk, 4 SAD map + transfer tables
    DBuf dec_in, dec_streams, dec_chunk_stream, dec_exit, dec_nsym, dec_neob, dec_entry, dec_symbase, dec_eobbase, dec_intra,
        dec_mv, dec_modes, dec_qp, dec_blk_start, dec_sym0, dec_syms, dec_levels, dec_lanes, dec_frame_ok, dec_step_frames,
        dec_step_streams;

    // host staging
    void* h_desc = nullptr;       // pinned descriptor staging
    size_t h_desc_cap = 0;

    // instrumentation of the last clip call: CUDA events on the compute stream around every kernel class
    bool timing = true;
    std::vector<cudaEvent_t> ev_pool;
    size_t ev_used = 0;
    struct Span { int cls, e0, e1; };
    std::vector<Span> spans;
    double last_ms[BVC_NUM_KERNEL_CLASSES] = {0, 0, 0, 0, 0};
    int64_t last_launches[BVC_NUM_KERNEL_CLASSES] = {0, 0, 0, 0, 0};
    double last_clip_ms = 0;
    int resident_frames = 0;
    size_t container_len = 0;   // bytes of the container the last clip call left in d_container
    // rate control on the clip path (bvc_set_rate_control): RCflag 1, per-row feedback chained on the device
    RcArgs rc{};
    double* d_rc_remaining = nullptr;
    // row-by-row (rate control) state
    bool row_open = false, row_intra = false;
    int row_nref = 0, row_next = 0;
    FrameLane row_fl{};
    long long* d_rowbits = nullptr;
};

static size_t default_coef_cap_words(const bvc_ctx* c);

// record an event on `st` (default: the compute stream) and return its index (-1 when timing is off)
static int tick(bvc_ctx* c, cudaStream_t st = nullptr) {
    if (!c->timing) return -1;
    if (c->ev_used == c->ev_pool.size()) {
        cudaEvent_t e;
        if (cudaEventCreate(&e) != cudaSuccess) return -1;
        c->ev_pool.push_back(e);
    }
    const int i = (int)c->ev_used++;
    cudaEventRecord(c->ev_pool[i], st ? st : c->st);
    return i;
}
static void span(bvc_ctx* c, int cls, int e0, int e1) {
    if (e0 >= 0 && e1 >= 0) c->spans.push_back({cls, e0, e1});
}

#define CK(call)                                                                                         \
    do {                                                                                                 \
        cudaError_t e__ = (call);                                                                        \
        if (e__ != cudaSuccess) {                                                                        \
            char b__[512];                                                                               \
            snprintf(b__, sizeof b__, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e__), __FILE__, __LINE__); \
            c->err = b__;                                                                                \
            return BVC_ERR_CUDA;                                                                         \
        }                                                                                                \
    } while (0)

static int fail(bvc_ctx* c, int code, const char* msg) {
    if (c) c->err = msg; else g_create_error = msg;
    return code;
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static int encode_map(bvc_ctx* c, CUtensorMap* out, int box_w, int box_h, int box_d = 1) {
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres));
    if (!fn || qres != cudaDriverEntryPointSuccess) return fail(c, BVC_ERR_CUDA, "cuTensorMapEncodeTiled entry point not available");
    cuuint64_t dims[3] = {(cuuint64_t)c->g.W, (cuuint64_t)c->g.H, (cuuint64_t)c->ref_planes};
    cuuint64_t strides[2] = {(cuuint64_t)c->g.pitch, (cuuint64_t)c->g.plane_bytes};
    cuuint32_t box[3] = {(cuuint32_t)box_w, (cuuint32_t)box_h, (cuuint32_t)box_d};
    cuuint32_t estr[3] = {1, 1, 1};
    CUresult r = ((EncodeTiledFn)fn)(out, CU_TENSOR_MAP_DATA_TYPE_UINT8, 3, c->ref_pool, dims, strides, box, estr,
                                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        char b[128];
        snprintf(b, sizeof b, "cuTensorMapEncodeTiled failed with CUresult %d", (int)r);
        return fail(c, BVC_ERR_CUDA, b);
    }
    return BVC_OK;
}

static int make_ref_map(bvc_ctx* c) {
    c->have_map = false;
    c->have_tall = false;
    c->have_fw_map = false;
    if (c->p.fast_me && c->g.bs % 4 == 0 && c->g.bs <= 32) {
        int bw = 0, bh = 0, bd = 1;
        fastme_window_box(c->g.bs, c->p.frac_me ? 4 : 1, &bw, &bh, &bd);
        int rc = encode_map(c, &c->fw_map, bw, bh, bd);
        if (rc != BVC_OK) return rc;
        c->have_fw_map = true;
    }
    // FastME: the tiled kernel fills the SAD map of radius 16 MV units (launch_fastme_any); its TMA box is sized for that
    const int R = c->p.fast_me ? (c->p.frac_me ? 8 : 16) : c->p.search_range;
    if (c->p.fast_me && !me_can_map(c->g.bs, R)) return BVC_OK;
    MeTileCfg cfg = me_tile_config(c->g.bs, R);
    if (!cfg.tiled && !cfg.narrow) return BVC_OK;
    int rc = encode_map(c, &c->ref_map, cfg.win_pitch, cfg.rows);
    if (rc != BVC_OK) return rc;
    c->have_map = true;
    const MeTileCfg tall = me_tile_config(c->g.bs, R, true);
    if (!c->p.fast_me && cfg.tiled && tall.tiled && tall.nby != cfg.nby) {
        rc = encode_map(c, &c->ref_map_tall, tall.win_pitch, tall.rows);
        if (rc != BVC_OK) return rc;
        c->have_tall = true;
    }
    return BVC_OK;
}

template <typename T>
static cudaError_t dalloc(T** p, size_t n) { return cudaMalloc((void**)p, n * sizeof(T)); }

extern "C" int bvc_create(bvc_ctx** out, int device, const bvc_params* p, int max_lanes) {
    if (!out || !p) return fail(nullptr, BVC_ERR_INVALID, "null argument");
    *out = nullptr;
    const int bs = p->block_size;
    if (!(bs == 4 || bs == 8 || bs == 16)) return fail(nullptr, BVC_ERR_UNSUPPORTED, "block_size must be 4, 8 or 16");
    if (p->width < bs || p->height < bs)  // block_predictor.py:70-71
        return fail(nullptr, BVC_ERR_INVALID, "frame smaller than block_size");
    if (p->width % bs || p->height % bs) return fail(nullptr, BVC_ERR_UNSUPPORTED, "width/height must be multiples of block_size (pad first)");
    int lg = 0; while ((1 << lg) < bs) lg++;
    if (p->qp < 0 || p->qp > lg + 7) return fail(nullptr, BVC_ERR_INVALID, "qp > log2(block_size) + 7");  // params.py:29-30
    if (p->nref_frames < 1 || p->nref_frames > BVC_MAX_REFS) return fail(nullptr, BVC_ERR_UNSUPPORTED, "nref_frames must be 1..8");
    if (!p->fast_me && (p->search_range < 0 || p->search_range * (p->frac_me ? 2 : 1) > 255))
        return fail(nullptr, BVC_ERR_UNSUPPORTED, "search_range out of supported range");
    if (p->i_period < 1) return fail(nullptr, BVC_ERR_INVALID, "i_period must be >= 1");
    if (max_lanes < 1) max_lanes = 1;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || device < 0 || device >= ndev)
        return fail(nullptr, BVC_ERR_CUDA, "no usable CUDA device (this library has no CPU fallback)");
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) != cudaSuccess || prop.major < 10)
        return fail(nullptr, BVC_ERR_CUDA, "device is not sm_100 class (kernels are built for sm_100a only)");

    bvc_ctx* c = new bvc_ctx();
    c->device = device;
    c->p = *p;
    c->max_lanes = max_lanes;
    Geom& g = c->g;
    g.W = p->width; g.H = p->height; g.bs = bs;
    g.pitch = (g.W + 15) / 16 * 16;
    g.bw = g.W / bs; g.bh = g.H / bs; g.nblk = g.bw * g.bh;
    g.plane_bytes = ((size_t)g.pitch * g.H + 255) / 256 * 256;
    c->pps = p->frac_me ? 4 : 1;
    c->slots = p->nref_frames + 1;
    c->blk_words = tq_blk_words(bs);
    auto boot = [&]() -> int {
        CK(cudaSetDevice(device));
        CK(cudaStreamCreateWithFlags(&c->st, cudaStreamNonBlocking));
        CK(cudaStreamCreateWithFlags(&c->st_h2d, cudaStreamNonBlocking));
        CK(cudaStreamCreateWithFlags(&c->st_d2h, cudaStreamNonBlocking));
        int prio_lo = 0, prio_hi = 0;
        CK(cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi));   // numerically lower = higher priority
        for (int gi = 0; gi < BVC_MAX_GROUPS; gi++) {
            CK(cudaStreamCreateWithPriority(&c->st_grp[gi], cudaStreamNonBlocking, prio_lo));
            CK(cudaStreamCreateWithPriority(&c->st_post[gi], cudaStreamNonBlocking, prio_hi));
            CK(cudaStreamCreateWithPriority(&c->st_pack[gi], cudaStreamNonBlocking, prio_hi));
            CK(cudaEventCreateWithFlags(&c->ev_me[gi], cudaEventDisableTiming));
            CK(cudaEventCreateWithFlags(&c->ev_post[gi], cudaEventDisableTiming));
            CK(cudaEventCreateWithFlags(&c->ev_tq[gi], cudaEventDisableTiming));
            CK(cudaEventCreateWithFlags(&c->ev_pack[gi], cudaEventDisableTiming));
        }
        if (const char* e = getenv("BVC_LANE_GROUPS")) c->ngroups = std::max(1, std::min(BVC_MAX_GROUPS, atoi(e)));
        if (const char* e = getenv("BVC_TAIL_SPLIT")) c->tail_split = atoi(e) != 0;
        if (const char* e = getenv("BVC_ME_TALL")) c->me_tall = atoi(e);
        if (const char* e = getenv("BVC_IQUAD")) c->iquad = std::max(0, std::min(2, atoi(e)));
        if (const char* e = getenv("BVC_TQ_CTAS")) c->tq_cta_cap = std::max(0, atoi(e));
        const size_t L = (size_t)max_lanes, nb = (size_t)g.nblk;
        c->ref_planes = L * c->slots * c->pps;
        CK(cudaMalloc((void**)&c->ref_pool, c->ref_planes * g.plane_bytes + 4096));
        CK(cudaMemset(c->ref_pool, 0, c->ref_planes * g.plane_bytes + 4096));
        CK(dalloc(&c->d_mv, 2 * L * nb));   // two sets, by step parity
        CK(dalloc(&c->d_modes, L * nb));
        CK(dalloc(&c->d_isad, L * nb));
        CK(dalloc(&c->d_qp_rows, L * g.bh));
        CK(dalloc(&c->d_blk_nbits, L * nb));
        CK(dalloc(&c->d_blk_bits, L * nb * c->blk_words));
        CK(dalloc(&c->d_levels, L * (size_t)g.W * g.H));   // every lane: I frames hand their levels to the entropy kernel
        CK(dalloc(&c->d_resid_mc, (size_t)g.W * g.H));
        CK(dalloc(&c->d_resid_nomc, (size_t)g.W * g.H));
        CK(dalloc(&c->d_coef_off, L * (nb + 1)));
        c->pack_tiles_n = pack_tiles(g.nblk);
        CK(dalloc(&c->d_pred_off, L * nb));
        CK(dalloc(&c->d_tile_tot, L * (size_t)(c->pack_tiles_n + 1) * 2));
        CK(dalloc(&c->d_tile_base, L * (size_t)(c->pack_tiles_n + 1) * 2));
        CK(dalloc(&c->d_row_bits, L * g.bh));
        CK(dalloc(&c->d_cmp, L));
        CK(dalloc(&c->d_rowbits, 1));
        CK(dalloc(&c->d_ticket, L));
        CK(dalloc(&c->d_top_mail, L * nb * bs));
        CK(cudaMemset(c->d_top_mail, 0, L * nb * bs * sizeof(uint32_t)));   // epoch 0 = never posted
        CK(dalloc(&c->d_rc_remaining, L));
        CK(cudaMemset(c->d_ticket, 0, L * sizeof(int)));
        c->coef_cap_words = default_coef_cap_words(c);
        c->pred_cap_words = nb * 3 + g.bh + 8;
        CK(dalloc(&c->d_overflow, 2));   // [0] a payload does not fit its container length field, [1] a stream slot is too small
        std::vector<int32_t> q(L * g.bh, p->qp);
        CK(cudaMemcpy(c->d_qp_rows, q.data(), q.size() * 4, cudaMemcpyHostToDevice));
        int rc = make_ref_map(c);
        if (rc != BVC_OK) return rc;
        return BVC_OK;
    };
    int rc = boot();
    if (rc != BVC_OK) {
        g_create_error = c->err;
        bvc_destroy(c);
        return rc;
    }
    *out = c;
    return BVC_OK;
}

extern "C" void bvc_destroy(bvc_ctx* c) {
    if (!c) return;
    cudaSetDevice(c->device);
    if (c->st) cudaStreamSynchronize(c->st);
    if (c->st_h2d) cudaStreamSynchronize(c->st_h2d);
    if (c->st_d2h) cudaStreamSynchronize(c->st_d2h);
    for (int gi = 0; gi < BVC_MAX_GROUPS; gi++) {
        if (c->st_grp[gi]) { cudaStreamSynchronize(c->st_grp[gi]); cudaStreamDestroy(c->st_grp[gi]); }
        if (c->st_post[gi]) { cudaStreamSynchronize(c->st_post[gi]); cudaStreamDestroy(c->st_post[gi]); }
        if (c->st_pack[gi]) { cudaStreamSynchronize(c->st_pack[gi]); cudaStreamDestroy(c->st_pack[gi]); }
        if (c->ev_me[gi]) cudaEventDestroy(c->ev_me[gi]);
        if (c->ev_post[gi]) cudaEventDestroy(c->ev_post[gi]);
        if (c->ev_tq[gi]) cudaEventDestroy(c->ev_tq[gi]);
        if (c->ev_pack[gi]) cudaEventDestroy(c->ev_pack[gi]);
    }
    cudaFree(c->in_pool); cudaFree(c->ref_pool); cudaFree(c->d_mv); cudaFree(c->d_modes); cudaFree(c->d_isad);
    cudaFree(c->d_qp_rows); cudaFree(c->d_blk_nbits); cudaFree(c->d_blk_bits); cudaFree(c->d_levels);
    cudaFree(c->d_resid_mc); cudaFree(c->d_resid_nomc); cudaFree(c->d_coef_off); cudaFree(c->d_row_bits);
    cudaFree(c->d_pred_off); cudaFree(c->d_tile_tot); cudaFree(c->d_tile_base);
    cudaFree(c->d_cmp); cudaFree(c->d_rowbits); cudaFree(c->d_ticket); cudaFree(c->d_top_mail); cudaFree(c->d_rc_remaining); cudaFree(c->d_me_lanes);
    cudaFree(c->d_fr_lanes); cudaFree(c->d_hp_src); cudaFree(c->d_hp_dst);
    cudaFree(c->d_coef_stream); cudaFree(c->d_pred_stream); cudaFree(c->d_frame_bits); cudaFree(c->d_frame_off);
    cudaFree(c->d_overflow); cudaFree(c->d_container); cudaFree(c->d_frag[0]); cudaFree(c->d_frag[1]);
    if (c->h_totals) cudaFreeHost(c->h_totals);
    for (bvc_ctx::DBuf* b : {&c->dec_in, &c->dec_streams, &c->dec_chunk_stream, &c->dec_exit, &c->dec_nsym, &c->dec_neob, &c->dec_entry,
                             &c->dec_symbase, &c->dec_eobbase, &c->dec_intra, &c->dec_mv, &c->dec_modes, &c->dec_qp, &c->dec_blk_start,
                             &c->dec_sym0, &c->dec_syms, &c->dec_levels, &c->dec_lanes, &c->dec_frame_ok, &c->dec_step_frames,
                             &c->dec_step_streams, &c->sad_map, &c->fastme_tab})
        cudaFree(b->p);
    if (c->h_desc) cudaFreeHost(c->h_desc);
    for (auto e : c->ev_pool) cudaEventDestroy(e);
    if (c->st) cudaStreamDestroy(c->st);
    if (c->st_h2d) cudaStreamDestroy(c->st_h2d);
    if (c->st_d2h) cudaStreamDestroy(c->st_d2h);
    delete c;
}

extern "C" const char* bvc_last_error(const bvc_ctx* c) { return c ? c->err.c_str() : g_create_error.c_str(); }
extern "C" int64_t bvc_launch_count(const bvc_ctx* c) { return c ? c->launches : 0; }
extern "C" int bvc_last_kernel_times(const bvc_ctx* c, double* ms, int64_t* launches, double* clip_ms) {
    if (!c) return BVC_ERR_INVALID;
    for (int i = 0; i < BVC_NUM_KERNEL_CLASSES; i++) {
        if (ms) ms[i] = c->last_ms[i];
        if (launches) launches[i] = c->last_launches[i];
    }
    if (clip_ms) *clip_ms = c->last_clip_ms;
    return BVC_OK;
}

extern "C" int bvc_set_qp(bvc_ctx* c, int qp) {
    if (!c) return BVC_ERR_INVALID;
    int lg = 0; while ((1 << lg) < c->g.bs) lg++;
    if (qp < 0 || qp > lg + 7) return fail(c, BVC_ERR_INVALID, "qp > log2(block_size) + 7");
    CK(cudaSetDevice(c->device));
    c->p.qp = qp;
    std::vector<int32_t> q((size_t)c->max_lanes * c->g.bh, qp);
    CK(cudaMemcpyAsync(c->d_qp_rows, q.data(), q.size() * 4, cudaMemcpyHostToDevice, c->st));
    CK(cudaStreamSynchronize(c->st));
    return BVC_OK;
}

extern "C" int bvc_set_lane_groups(bvc_ctx* c, int groups) {
    if (!c || groups < 1 || groups > BVC_MAX_GROUPS) return BVC_ERR_INVALID;
    c->ngroups = groups;
    return BVC_OK;
}

extern "C" int bvc_set_fastme_direct(bvc_ctx* c, int on) {
    if (!c) return BVC_ERR_INVALID;
    c->fastme_direct = (on >= 2 && on <= 4) ? on : (on ? 1 : 0);
    return BVC_OK;
}

extern "C" int64_t bvc_me_work_per_frame(const bvc_ctx* c, int nref_avail) {
    if (!c || c->p.fast_me) return 0;
    const Geom& g = c->g;
    const int sc = c->p.frac_me ? 2 : 1, R = c->p.search_range * sc;
    // valid candidates per block factorise into x and y counts (block_predictor.py:116-143)
    auto count = [&](int o, int n, int size) {  // positions m in [-R,R] with 0 <= sc*o+m and sc*o+m+sc*n <= sc*size
        int lo = std::max(-R, -sc * o), hi = std::min(R, sc * (size - n - o));
        return std::max(0, hi - lo + 1);
    };
    int64_t tot = 0;
    for (int by = 0; by < g.bh; by++)
        for (int bx = 0; bx < g.bw; bx++) tot += (int64_t)count(bx * g.bs, g.bs, g.W) * count(by * g.bs, g.bs, g.H);
    return tot * g.bs * g.bs * nref_avail;
}

// ---------------------------------------------------------------------------------------------
static int ensure_in_pool(bvc_ctx* c, size_t planes) {
    if (planes <= c->in_planes) return BVC_OK;
    if (c->in_pool) CK(cudaFree(c->in_pool));
    c->in_pool = nullptr;
    c->in_planes = 0;
    c->resident_frames = 0;   // whatever bvc_clip_upload left in the old pool is gone
    CK(cudaMalloc((void**)&c->in_pool, planes * c->g.plane_bytes + 4096));
    c->in_planes = planes;
    return BVC_OK;
}
static int ensure_lane_desc(bvc_ctx* c, size_t n) {
    if (n <= c->lane_desc_cap) return BVC_OK;
    if (c->d_me_lanes) CK(cudaFree(c->d_me_lanes));
    if (c->d_fr_lanes) CK(cudaFree(c->d_fr_lanes));
    CK(dalloc(&c->d_me_lanes, n));
    CK(dalloc(&c->d_fr_lanes, n));
    if (c->d_hp_src) CK(cudaFree(c->d_hp_src));
    if (c->d_hp_dst) CK(cudaFree(c->d_hp_dst));
    CK(dalloc(&c->d_hp_src, n * BVC_MAX_REFS));
    CK(dalloc(&c->d_hp_dst, n * BVC_MAX_REFS));
    c->lane_desc_cap = n;
    return BVC_OK;
}
static int ensure_streams(bvc_ctx* c, size_t slots) {
    if (slots <= c->stream_slots) return BVC_OK;
    cudaFree(c->d_coef_stream); cudaFree(c->d_pred_stream); cudaFree(c->d_frame_bits); cudaFree(c->d_frame_off);
    c->d_coef_stream = c->d_pred_stream = nullptr; c->d_frame_bits = c->d_frame_off = nullptr;
    c->stream_slots = 0;
    CK(dalloc(&c->d_coef_stream, slots * c->coef_cap_words));
    CK(dalloc(&c->d_pred_stream, slots * c->pred_cap_words));
    CK(dalloc(&c->d_frame_bits, slots * 2));
    CK(dalloc(&c->d_frame_off, slots + BVC_MAX_GROUPS + 1));   // one offset table (n + 1 entries) per container part
    c->stream_slots = slots;
    return BVC_OK;
}
static int ensure_container(bvc_ctx* c, size_t bytes) {
    if (bytes <= c->container_cap) return BVC_OK;
    cudaFree(c->d_container);
    c->d_container = nullptr; c->container_cap = 0;
    CK(cudaMalloc((void**)&c->d_container, bytes + 64));
    c->container_cap = bytes;
    return BVC_OK;
}
// Words reserved per frame for the coefficient stream.  The worst case (every coefficient at the largest magnitude) is 26
// bits per pixel; the default reserves 6 bits per pixel (at least 1 MB, never more than the worst case) -- pack_scan
// reports a frame that does not fit and the call fails with BVC_ERR_NOMEM instead of overrunning
// (bvc_set_stream_slot_bytes raises the reservation).
static size_t default_coef_cap_words(const bvc_ctx* c) {
    const size_t worst = (size_t)c->g.nblk * c->blk_words + 8;
    const size_t want = c->slot_bytes ? c->slot_bytes : std::max((size_t)c->g.W * c->g.H / 4 * 3, (size_t)1 << 20);
    return std::min(worst, want / 4 + 8);
}
static int ensure_fragments(bvc_ctx* c, size_t bytes, int n) {
    if (bytes <= c->frag_cap && n <= c->nfrag) return BVC_OK;
    for (int i = 0; i < 2; i++) { cudaFree(c->d_frag[i]); c->d_frag[i] = nullptr; }
    c->frag_cap = 0; c->nfrag = 0;
    const size_t cap = std::max(bytes, c->frag_cap);
    for (int i = 0; i < n; i++) CK(cudaMalloc((void**)&c->d_frag[i], cap + 64));
    c->frag_cap = cap; c->nfrag = n;
    return BVC_OK;
}
static int ensure_pinned(bvc_ctx* c, void** p, size_t* cap, size_t need) {
    if (need <= *cap) return BVC_OK;
    if (*p) CK(cudaFreeHost(*p));
    *p = nullptr; *cap = 0;
    size_t n = std::max(need, (size_t)1 << 20);
    CK(cudaHostAlloc(p, n, cudaHostAllocDefault));
    *cap = n;
    return BVC_OK;
}

static inline int ring_plane(const bvc_ctx* c, int lane, int slot) { return (lane * c->slots + slot) * c->pps; }
static inline uint8_t* plane_ptr(const bvc_ctx* c, int plane) { return c->ref_pool + (size_t)plane * c->g.plane_bytes; }

static int upload_plane(bvc_ctx* c, uint8_t* dst, const uint8_t* src) {
    CK(cudaMemcpy2DAsync(dst, c->g.pitch, src, c->g.W, c->g.W, c->g.H, cudaMemcpyHostToDevice, c->st));
    return BVC_OK;
}
static int download_plane(bvc_ctx* c, uint8_t* dst, const uint8_t* src, cudaStream_t st = nullptr) {
    if (c->g.pitch == c->g.W) CK(cudaMemcpyAsync(dst, src, (size_t)c->g.W * c->g.H, cudaMemcpyDeviceToHost, st ? st : c->st));
    else CK(cudaMemcpy2DAsync(dst, c->g.W, src, c->g.pitch, c->g.W, c->g.H, cudaMemcpyDeviceToHost, st ? st : c->st));
    return BVC_OK;
}

struct StepPlan {
    int nl = 0;
    int k = 0;            // position of the step's frames inside their GOPs (parity picks the motion-vector set)
    int nref = 0;         // references every lane of the step sees (frame k of its GOP: min(k, nRefFrames))
    bool intra = false;
    size_t desc_off = 0;  // offset (in lanes) into the device descriptor arrays
};

// FastME for lanes [L0, L0+nl): SAD map by the tiled search kernel + table walk, or the direct kernel.
static int launch_fastme_any(bvc_ctx* c, const MeArgs& m, int nl, size_t L0, cudaStream_t st, int step_lanes) {
    const Geom& g = c->g;
    const int Rm = c->p.frac_me ? 8 : 16;   // 16 MV units around the block: where the walk can look before it stops
    auto scratch_for = [&](size_t total, size_t offset, char** out) -> int {
        if (total + 256 > c->fastme_tab.cap) {
            if (c->fastme_tab.p) CK(cudaFree(c->fastme_tab.p));
            c->fastme_tab.p = nullptr; c->fastme_tab.cap = 0;
            CK(cudaMalloc(&c->fastme_tab.p, total + 256));
            c->fastme_tab.cap = total + 256;
        }
        // lane groups run concurrently on their own streams: each gets its own slice of the scratch
        *out = static_cast<char*>(c->fastme_tab.p) + offset;
        return BVC_OK;
    };
    // Serial walk over TMA-staged windows against SAD map + transfer tables: the walk costs ~1 us per block whatever the number
    // of frames in flight (one CTA each), the map 20-45 ns per block and frame, so the walk wins from about 26 lanes per step.
    const bool can_window = c->have_fw_map && fastme_window_smem(m, c->p.nref_frames) <= 200 * 1024;
    const bool can_tables = c->have_map && me_can_map(g.bs, Rm);
    const bool window = c->fastme_direct == 3 || (c->fastme_direct == 0 && g.bs == 16 && (step_lanes >= 26 || !can_tables));
    if (window && can_window) {
        CK(launch_fastme_window(&c->fw_map, m, nl, c->p.nref_frames, c->ref_pool, g.plane_bytes, g.pitch, c->d_cmp + L0, st));
        return BVC_OK;
    }
    if (c->fastme_direct != 1 && can_tables) {
        const size_t n1 = 2 * (size_t)Rm + 1;
        const size_t stride = (n1 * n1 + 7) / 8 * 8;   // 16-byte multiples: the walk stages a block's table with cp.async
        const size_t per_lane = (size_t)c->p.nref_frames * m.nphase * g.nblk * stride;
        const size_t need = per_lane * (size_t)c->max_lanes * sizeof(uint16_t) + 256;
        if (need > c->sad_map.cap) {
            if (c->sad_map.p) CK(cudaFree(c->sad_map.p));
            c->sad_map.p = nullptr; c->sad_map.cap = 0;
            CK(cudaMalloc(&c->sad_map.p, need));
            c->sad_map.cap = need;
        }
        MeArgs mm = m;
        mm.R = Rm;
        mm.Rh = Rm * m.sc;
        mm.sad_map = reinterpret_cast<uint16_t*>(c->sad_map.p) + L0 * per_lane;
        mm.max_refs = c->p.nref_frames;
        mm.map_stride = (int)stride;
        CK(cudaMemsetAsync(mm.sad_map, 0xFF, (size_t)nl * per_lane * sizeof(uint16_t), st));
        CK(launch_me_fullsearch(&c->ref_map, mm, nl, c->ref_pool, g.plane_bytes, g.pitch, st));
        if (c->fastme_direct == 2) {
            CK(launch_fastme_walk(mm, nl, c->ref_pool, g.plane_bytes, g.pitch, c->d_cmp + L0, st));
            c->launches += 1;
            return BVC_OK;
        }
        char* scratch = nullptr;
        int rc = scratch_for(fastme_table_bytes(c->max_lanes, g.nblk), fastme_table_bytes((int)L0, g.nblk), &scratch);
        if (rc != BVC_OK) return rc;
        CK(launch_fastme_table(mm, nl, c->ref_pool, g.plane_bytes, g.pitch, scratch, c->d_cmp + L0, st));
        c->launches += 3;
        return BVC_OK;
    }
    CK(launch_fastme(m, nl, c->ref_pool, g.plane_bytes, g.pitch, c->d_cmp + L0, st));
    return BVC_OK;
}

// Enqueue the kernels of lanes [l0, l0+nl) of one step: the motion search on `st_me`, everything after it on
// `st_post` (the same stream for the frame-level calls).  Per-lane scratch arrays are indexed by the lane
// inside a launch, so a lane group simply gets base pointers advanced by l0 lanes.
static int enqueue_step(bvc_ctx* c, const StepPlan& sp, bool frame_api, cudaStream_t st_me, cudaStream_t st_post, int l0, int nl,
                        cudaEvent_t ev_me_done, cudaStream_t st_pack = nullptr, cudaEvent_t ev_tq_done = nullptr,
                        cudaEvent_t ev_pack_done = nullptr) {
    const Geom& g = c->g;
    const size_t nb = (size_t)g.nblk, L0 = (size_t)l0;
    const bool side_pack = st_pack && st_pack != st_post;
    int4* const mv_set = c->d_mv + (size_t)(sp.k & 1) * c->max_lanes * nb;
    // the transform of this step rewrites the per-block strings the previous step's assembly reads
    if (side_pack) CK(cudaStreamWaitEvent(st_post, ev_pack_done, 0));
    TqArgs t{};
    t.cur_base = c->in_pool; t.cur_plane_bytes = g.plane_bytes; t.cur_pitch = g.pitch;
    t.ref_base = c->ref_pool; t.ref_plane_bytes = g.plane_bytes; t.ref_pitch = g.pitch;
    t.lanes = c->d_fr_lanes + sp.desc_off + L0;
    t.mv = mv_set + L0 * nb; t.modes = c->d_modes + L0 * nb; t.isad = c->d_isad + L0 * nb; t.qp_rows = c->d_qp_rows + L0 * g.bh;
    t.levels = (frame_api || sp.intra) ? c->d_levels + L0 * (size_t)g.W * g.H : nullptr;
    t.resid_mc = frame_api ? c->d_resid_mc : nullptr;
    t.resid_nomc = frame_api ? c->d_resid_nomc : nullptr;
    t.blk_bits = c->d_blk_bits + L0 * nb * c->blk_words; t.blk_nbits = c->d_blk_nbits + L0 * nb; t.blk_words = c->blk_words;
    t.W = g.W; t.H = g.H; t.bs = g.bs; t.bw = g.bw; t.bh = g.bh; t.nblk = g.nblk;
    t.frac = c->p.frac_me; t.multi_ref = c->p.nref_frames > 1; t.ticket = c->d_ticket + L0;
    t.top_mail = c->d_top_mail + L0 * nb * g.bs;
    if (sp.intra) { c->epoch = (c->epoch % 0xFFFFFEu) + 1; t.epoch = c->epoch; }   // one epoch per I frame (all its rows, also row by row)
    t.row_begin = 0; t.row_count = g.bh;
    t.quad = c->iquad;
    t.cta_cap = (!frame_api && st_post != st_me) ? c->tq_cta_cap : 0;   // beside the other group's search: see tq_pframe_kernel
    // rate control (RCflag 1) on the clip path: the transform runs block row by block row, and the launch that
    // accounts a row's bits also picks the next row's QP -- the whole chain stays on the device, all lanes in lock step
    const bool rc_rows = !frame_api && c->rc.n > 0;
    RcArgs rc = c->rc;
    rc.remaining = c->d_rc_remaining + L0;
    rc.qp_rows = c->d_qp_rows + L0 * g.bh;
    PackArgs pk{};
    pk.mv = t.mv; pk.modes = t.modes; pk.qp_rows = t.qp_rows;
    pk.blk_bits = t.blk_bits; pk.blk_nbits = t.blk_nbits; pk.blk_words = c->blk_words;
    pk.coef_off = c->d_coef_off + L0 * (nb + 1);
    pk.pred_off = c->d_pred_off + L0 * nb; pk.tiles = c->pack_tiles_n;
    pk.tile_tot = c->d_tile_tot + L0 * (size_t)(c->pack_tiles_n + 1) * 2; pk.tile_base = c->d_tile_base + L0 * (size_t)(c->pack_tiles_n + 1) * 2;
    pk.lanes = t.lanes;
    pk.coef_stream = c->d_coef_stream; pk.pred_stream = c->d_pred_stream;
    pk.frame_bits = c->d_frame_bits; pk.row_bits = c->d_row_bits + L0 * g.bh;
    pk.coef_cap_words = c->coef_cap_words; pk.pred_cap_words = c->pred_cap_words; pk.slot_overflow = c->d_overflow + 1;
    pk.bw = g.bw; pk.bh = g.bh; pk.nblk = g.nblk; pk.base_qp = c->p.qp;
    pk.intra = sp.intra; pk.with_ref = c->p.nref_frames > 1;
    auto transform_rows = [&](bool intra) -> int {
        CK(launch_rc_begin(rc, nl, g.bh, st_post));
        for (int row = 0; row < g.bh; row++) {
            t.row_begin = row; t.row_count = 1;
            if (intra) CK(launch_tq_iframe(t, nl, st_post));
            else CK(launch_tq_pframe(t, nl, st_post));
            CK(launch_row_bits_rc(pk, rc, nl, row, nullptr, st_post));
        }
        t.row_begin = 0; t.row_count = g.bh;
        c->launches += 1 + (intra ? 3 : 2) * g.bh;
        return BVC_OK;
    };
    if (sp.intra) {
        const int e0 = tick(c, st_post);
        if (rc_rows) { int rcr = transform_rows(true); if (rcr != BVC_OK) return rcr; }
        else if (side_pack) { CK(launch_tq_iframe(t, nl, st_post, false)); c->launches += 1; }   // entropy coding follows on st_pack
        else { CK(launch_tq_iframe(t, nl, st_post)); c->launches += 2; }
        span(c, BVC_K_TQ_I, e0, tick(c, st_post));
    } else {
        MeArgs m{};
        m.cur_base = c->in_pool; m.cur_plane_bytes = g.plane_bytes; m.cur_pitch = g.pitch;
        m.lanes = c->d_me_lanes + sp.desc_off + L0;
        m.out = mv_set + L0 * nb;
        m.W = g.W; m.H = g.H; m.bs = g.bs; m.bw = g.bw; m.bh = g.bh; m.nblk = g.nblk;
        m.sc = c->p.frac_me ? 2 : 1;
        m.nphase = c->p.frac_me ? 4 : 1;
        m.R = c->p.search_range;
        m.Rh = c->p.search_range * m.sc;
        m.uniform_nref = sp.nref;
        m.tail_split = c->tail_split && st_post == st_me;   // with lane groups the other group's kernels fill the tail
        m.tall_mode = c->me_tall;
        const int e0 = tick(c, st_me);
        if (c->p.fast_me) {
            int rcf = launch_fastme_any(c, m, nl, L0, st_me, sp.nl);
            if (rcf != BVC_OK) return rcf;
        } else {
            CK(launch_me_fullsearch(c->have_map ? &c->ref_map : nullptr, m, nl, c->ref_pool, g.plane_bytes, g.pitch, st_me,
                                    c->have_tall ? &c->ref_map_tall : nullptr));
        }
        const int e1 = tick(c, st_me);
        span(c, BVC_K_ME, e0, e1);
        int e1p = e1;
        if (st_post != st_me) {
            CK(cudaEventRecord(ev_me_done, st_me));
            CK(cudaStreamWaitEvent(st_post, ev_me_done, 0));
            e1p = tick(c, st_post);
        }
        if (rc_rows) { int rcr = transform_rows(false); if (rcr != BVC_OK) return rcr; c->launches += 1; }
        else { CK(launch_tq_pframe(t, nl, st_post)); c->launches += 2; }
        span(c, BVC_K_TQ_P, e1p, tick(c, st_post));
    }
    if (side_pack) {
        CK(cudaEventRecord(ev_tq_done, st_post));
        CK(cudaStreamWaitEvent(st_pack, ev_tq_done, 0));
        if (sp.intra && !rc_rows) { CK(launch_tq_ientropy(t, nl, st_pack)); c->launches += 1; }
        CK(launch_pack(pk, nl, st_pack));
        CK(cudaEventRecord(ev_pack_done, st_pack));
        c->launches += c->pack_tiles_n > 1 ? 3 : 2;
        return BVC_OK;
    }
    const int ep = tick(c, st_post);
    CK(launch_pack(pk, nl, st_post));
    span(c, BVC_K_PACK, ep, tick(c, st_post));
    c->launches += c->pack_tiles_n > 1 ? 3 : 2;
    return BVC_OK;
}
static int enqueue_step(bvc_ctx* c, const StepPlan& sp, bool frame_api) {
    return enqueue_step(c, sp, frame_api, c->st, c->st, 0, sp.nl, nullptr);
}

// half-pel phase planes (K2).  Descriptors (phase-0 plane pointers and the three destination planes
// behind them) are uploaded ahead of time; the launch itself is asynchronous.
static int upload_halfpel_desc(bvc_ctx* c, const std::vector<int>& planes, size_t desc_off) {
    const size_t n = planes.size();
    if (!n) return BVC_OK;
    std::vector<const uint8_t*> src(n);
    std::vector<uint8_t*> dst(n);
    for (size_t i = 0; i < n; i++) { src[i] = plane_ptr(c, planes[i]); dst[i] = plane_ptr(c, planes[i] + 1); }
    CK(cudaMemcpy(c->d_hp_src + desc_off, src.data(), n * sizeof(void*), cudaMemcpyHostToDevice));
    CK(cudaMemcpy(c->d_hp_dst + desc_off, dst.data(), n * sizeof(void*), cudaMemcpyHostToDevice));
    return BVC_OK;
}
static int enqueue_halfpel(bvc_ctx* c, size_t desc_off, int n, cudaStream_t st = nullptr) {
    if (!c->p.frac_me || n <= 0) return BVC_OK;
    if (!st) st = c->st;
    const Geom& g = c->g;
    const int e0 = tick(c, st);
    CK(launch_halfpel(c->d_hp_src + desc_off, c->d_hp_dst + desc_off, n, g.W, g.H, g.pitch, g.plane_bytes, st));
    span(c, BVC_K_HALFPEL, e0, tick(c, st));
    c->launches += 1;
    return BVC_OK;
}

// ---------------------------------------------------------------------------------------------
// frame-level API
// Upload the current frame and its reference window into lane 0, build the half-pel planes, write the lane
// descriptors and per-row QPs.  Leaves everything enqueued on c->st.
static int frame_prepare(bvc_ctx* c, const uint8_t* cur, const uint8_t* const* refs, int nref_avail, const int32_t* qp_rows,
                         bool intra, FrameLane* fl_out) {
    const Geom& g = c->g;
    CK(cudaSetDevice(c->device));
    if (!cur) return fail(c, BVC_ERR_INVALID, "cur is null");
    if (!intra && (nref_avail < 1 || nref_avail > c->p.nref_frames || !refs))
        return fail(c, BVC_ERR_INVALID, "nref_avail must be 1..nref_frames");
    int rc;
    if ((rc = ensure_in_pool(c, 1)) != BVC_OK) return rc;
    if ((rc = ensure_lane_desc(c, 1)) != BVC_OK) return rc;
    if ((rc = ensure_streams(c, 1)) != BVC_OK) return rc;
    c->resident_frames = 0;   // plane 0 of the input pool is overwritten: a clip uploaded earlier is no longer resident
    if ((rc = upload_plane(c, c->in_pool, cur)) != BVC_OK) return rc;
    MeLane ml{};
    FrameLane fl{};
    ml.cur_plane = 0; fl.cur_plane = 0; fl.slot = 0;
    ml.nref = fl.nref = intra ? 0 : nref_avail;
    std::vector<int> hp;
    for (int k = 0; k < ml.nref; k++) {
        ml.ref_plane[k] = fl.ref_plane[k] = ring_plane(c, 0, k);
        if ((rc = upload_plane(c, plane_ptr(c, ml.ref_plane[k]), refs[k])) != BVC_OK) return rc;
        hp.push_back(ml.ref_plane[k]);
    }
    fl.out_plane = ring_plane(c, 0, ml.nref);
    CK(cudaMemcpyAsync(c->d_me_lanes, &ml, sizeof ml, cudaMemcpyHostToDevice, c->st));
    CK(cudaMemcpyAsync(c->d_fr_lanes, &fl, sizeof fl, cudaMemcpyHostToDevice, c->st));
    std::vector<int32_t> q(g.bh, c->p.qp);
    if (qp_rows) q.assign(qp_rows, qp_rows + g.bh);
    CK(cudaMemcpyAsync(c->d_qp_rows, q.data(), g.bh * 4, cudaMemcpyHostToDevice, c->st));
    CK(cudaStreamSynchronize(c->st));
    if (c->p.frac_me) {
        if ((rc = upload_halfpel_desc(c, hp, 0)) != BVC_OK) return rc;
        if ((rc = enqueue_halfpel(c, 0, (int)hp.size())) != BVC_OK) return rc;
    }
    *fl_out = fl;
    return BVC_OK;
}

static int launch_me_lane0(bvc_ctx* c) {
    const Geom& g = c->g;
    MeArgs m{};
    m.cur_base = c->in_pool; m.cur_plane_bytes = g.plane_bytes; m.cur_pitch = g.pitch;
    m.lanes = c->d_me_lanes; m.out = c->d_mv;
    m.W = g.W; m.H = g.H; m.bs = g.bs; m.bw = g.bw; m.bh = g.bh; m.nblk = g.nblk;
    m.sc = c->p.frac_me ? 2 : 1; m.nphase = c->p.frac_me ? 4 : 1; m.R = c->p.search_range; m.Rh = m.R * m.sc; m.tail_split = c->tail_split;
    if (c->p.fast_me) { int rcf = launch_fastme_any(c, m, 1, 0, c->st, 1); if (rcf != BVC_OK) return rcf; }
    else CK(launch_me_fullsearch(c->have_map ? &c->ref_map : nullptr, m, 1, c->ref_pool, g.plane_bytes, g.pitch, c->st));
    c->launches += 1;
    return BVC_OK;
}

// Download the outputs of the frame in lane 0 (after ME, or after ME + transform + pack).
static int frame_collect(bvc_ctx* c, bvc_frame_out* out, bool intra, int nref_avail, const FrameLane& fl, bool me_only,
                         int32_t* mv_out, int32_t* sad_out, int64_t* cmp_out) {
    const Geom& g = c->g;
    int rc;
    std::vector<int4> hmv;
    std::vector<int32_t> hsad(g.nblk);
    long long hcmp = 0;
    if (!intra) {
        hmv.resize(g.nblk);
        CK(cudaMemcpyAsync(hmv.data(), c->d_mv, (size_t)g.nblk * sizeof(int4), cudaMemcpyDeviceToHost, c->st));
        if (c->p.fast_me) CK(cudaMemcpyAsync(&hcmp, c->d_cmp, sizeof hcmp, cudaMemcpyDeviceToHost, c->st));
    } else {
        CK(cudaMemcpyAsync(hsad.data(), c->d_isad, (size_t)g.nblk * 4, cudaMemcpyDeviceToHost, c->st));
    }
    long long fb[2] = {0, 0};
    if (!me_only) CK(cudaMemcpyAsync(fb, c->d_frame_bits, sizeof fb, cudaMemcpyDeviceToHost, c->st));
    CK(cudaStreamSynchronize(c->st));
    if (!intra) {
        for (int b = 0; b < g.nblk; b++) hsad[b] = hmv[b].w;
        int32_t* mvd = me_only ? mv_out : (out ? out->mv : nullptr);
        if (mvd) for (int b = 0; b < g.nblk; b++) { mvd[3 * b] = hmv[b].x; mvd[3 * b + 1] = hmv[b].y; mvd[3 * b + 2] = hmv[b].z; }
    }
    int64_t cmp = 0;
    if (intra) cmp = 2LL * g.nblk;  // params.py:62
    else if (c->p.fast_me) cmp = hcmp;
    else {
        const int64_t n1 = 2LL * c->p.search_range * (c->p.frac_me ? 2 : 1) + 1;
        cmp = (int64_t)g.nblk * nref_avail * n1 * n1;  // nominal count, block_predictor.py:91
    }
    if (me_only) {
        if (sad_out) memcpy(sad_out, hsad.data(), (size_t)g.nblk * 4);
        if (cmp_out) *cmp_out = cmp;
        return BVC_OK;
    }
    if (!out) return BVC_OK;
    if (out->sad) memcpy(out->sad, hsad.data(), (size_t)g.nblk * 4);
    // avg_mae: sum of per-block MAE in raster order / number of blocks (PFrame.py:67,88; IFrame.py:51,76)
    double s = 0.0;
    for (int b = 0; b < g.nblk; b++) s += (double)hsad[b] / (double)(g.bs * g.bs);
    out->avg_mae = s / (double)g.nblk;
    out->mae_comparisons = cmp;
    out->pred_nbits = fb[0];
    out->coef_nbits = fb[1];
    const size_t pb = (size_t)((fb[0] + 7) / 8), cb = (size_t)((fb[1] + 7) / 8);
    if ((out->pred_bytes && pb > out->pred_cap) || (out->coef_bytes && cb > out->coef_cap))
        return fail(c, BVC_ERR_NOMEM, "output bit buffer too small");
    if (out->pred_bytes && pb) CK(cudaMemcpyAsync(out->pred_bytes, c->d_pred_stream, pb, cudaMemcpyDeviceToHost, c->st));
    if (out->coef_bytes && cb) CK(cudaMemcpyAsync(out->coef_bytes, c->d_coef_stream, cb, cudaMemcpyDeviceToHost, c->st));
    if (out->recon && (rc = download_plane(c, out->recon, plane_ptr(c, fl.out_plane))) != BVC_OK) return rc;
    if (out->levels) CK(cudaMemcpyAsync(out->levels, c->d_levels, (size_t)g.W * g.H * 2, cudaMemcpyDeviceToHost, c->st));
    if (out->resid_mc) CK(cudaMemcpyAsync(out->resid_mc, c->d_resid_mc, (size_t)g.W * g.H, cudaMemcpyDeviceToHost, c->st));
    if (out->resid_nomc && !intra) CK(cudaMemcpyAsync(out->resid_nomc, c->d_resid_nomc, (size_t)g.W * g.H, cudaMemcpyDeviceToHost, c->st));
    if (out->modes && intra) CK(cudaMemcpyAsync(out->modes, c->d_modes, (size_t)g.nblk * 4, cudaMemcpyDeviceToHost, c->st));
    std::vector<long long> rb(g.bh);
    CK(cudaMemcpyAsync(rb.data(), c->d_row_bits, (size_t)g.bh * 8, cudaMemcpyDeviceToHost, c->st));
    CK(cudaStreamSynchronize(c->st));
    if (out->bits_per_row) for (int r = 0; r < g.bh; r++) out->bits_per_row[r] = rb[r];
    // restore the base-QP rows for later clip calls
    std::vector<int32_t> qb((size_t)c->max_lanes * g.bh, c->p.qp);
    CK(cudaMemcpy(c->d_qp_rows, qb.data(), qb.size() * 4, cudaMemcpyHostToDevice));
    return BVC_OK;
}

static int frame_common(bvc_ctx* c, const uint8_t* cur, const uint8_t* const* refs, int nref_avail, const int32_t* qp_rows,
                        bvc_frame_out* out, bool intra, bool me_only, int32_t* mv_out, int32_t* sad_out, int64_t* cmp_out) {
    FrameLane fl{};
    int rc;
    c->row_open = false;
    if ((rc = frame_prepare(c, cur, refs, nref_avail, qp_rows, intra, &fl)) != BVC_OK) return rc;
    StepPlan sp;
    sp.nl = 1; sp.intra = intra; sp.desc_off = 0; sp.nref = intra ? 0 : nref_avail;
    if (me_only) {
        if ((rc = launch_me_lane0(c)) != BVC_OK) return rc;
    } else {
        c->ev_used = 0; c->spans.clear();
        if ((rc = enqueue_step(c, sp, true)) != BVC_OK) return rc;
    }
    return frame_collect(c, out, intra, nref_avail, fl, me_only, mv_out, sad_out, cmp_out);
}

extern "C" int bvc_encode_iframe(bvc_ctx* c, const uint8_t* cur, const int32_t* qp_rows, bvc_frame_out* out) {
    if (!c) return BVC_ERR_INVALID;
    return frame_common(c, cur, nullptr, 0, qp_rows, out, true, false, nullptr, nullptr, nullptr);
}
extern "C" int bvc_encode_pframe(bvc_ctx* c, const uint8_t* cur, const uint8_t* const* refs, int nref_avail,
                                 const int32_t* qp_rows, bvc_frame_out* out) {
    if (!c) return BVC_ERR_INVALID;
    return frame_common(c, cur, refs, nref_avail, qp_rows, out, false, false, nullptr, nullptr, nullptr);
}
extern "C" int bvc_me_search(bvc_ctx* c, const uint8_t* cur, const uint8_t* const* refs, int nref_avail, int32_t* mv,
                             int32_t* sad, int64_t* comparisons) {
    if (!c) return BVC_ERR_INVALID;
    return frame_common(c, cur, refs, nref_avail, nullptr, nullptr, false, true, mv, sad, comparisons);
}

// ---- row-by-row encoding: the rate-control feedback loop (RCflag = 1) -------------------------------
// Frame.get_rc_qp (Frame.py:168-188) picks the QP of block row k from the bits rows < k actually
// consumed (PFrame.py:53-83, IFrame.py:38-70), so rows are encoded one launch at a time and their bit
// count is read back before the next row.  Motion estimation does not depend on the QP and runs once,
// for the whole frame, in bvc_frame_begin.
static void fill_row_args(bvc_ctx* c, TqArgs& t, PackArgs& pk, bool intra) {
    const Geom& g = c->g;
    t = TqArgs{};
    t.cur_base = c->in_pool; t.cur_plane_bytes = g.plane_bytes; t.cur_pitch = g.pitch;
    t.ref_base = c->ref_pool; t.ref_plane_bytes = g.plane_bytes; t.ref_pitch = g.pitch;
    t.lanes = c->d_fr_lanes;
    t.mv = c->d_mv; t.modes = c->d_modes; t.isad = c->d_isad; t.qp_rows = c->d_qp_rows;
    t.levels = c->d_levels; t.resid_mc = c->d_resid_mc; t.resid_nomc = c->d_resid_nomc;
    t.blk_bits = c->d_blk_bits; t.blk_nbits = c->d_blk_nbits; t.blk_words = c->blk_words;
    t.W = g.W; t.H = g.H; t.bs = g.bs; t.bw = g.bw; t.bh = g.bh; t.nblk = g.nblk;
    t.frac = c->p.frac_me; t.multi_ref = c->p.nref_frames > 1; t.ticket = c->d_ticket;
    t.top_mail = c->d_top_mail; t.epoch = c->epoch; t.quad = c->iquad;
    pk = PackArgs{};
    pk.mv = c->d_mv; pk.modes = c->d_modes; pk.qp_rows = c->d_qp_rows;
    pk.blk_bits = c->d_blk_bits; pk.blk_nbits = c->d_blk_nbits; pk.blk_words = c->blk_words;
    pk.coef_off = c->d_coef_off; pk.lanes = c->d_fr_lanes;
    pk.pred_off = c->d_pred_off; pk.tiles = c->pack_tiles_n; pk.tile_tot = c->d_tile_tot; pk.tile_base = c->d_tile_base;
    pk.coef_stream = c->d_coef_stream; pk.pred_stream = c->d_pred_stream;
    pk.frame_bits = c->d_frame_bits; pk.row_bits = c->d_row_bits;
    pk.coef_cap_words = c->coef_cap_words; pk.pred_cap_words = c->pred_cap_words; pk.slot_overflow = c->d_overflow + 1;
    pk.bw = g.bw; pk.bh = g.bh; pk.nblk = g.nblk; pk.base_qp = c->p.qp;
    pk.intra = intra; pk.with_ref = c->p.nref_frames > 1;
}

extern "C" int bvc_frame_begin(bvc_ctx* c, const uint8_t* cur, const uint8_t* const* refs, int nref_avail, int intra) {
    if (!c) return BVC_ERR_INVALID;
    int rc;
    c->row_open = false;
    if ((rc = frame_prepare(c, cur, refs, nref_avail, nullptr, intra != 0, &c->row_fl)) != BVC_OK) return rc;
    if (!intra && (rc = launch_me_lane0(c)) != BVC_OK) return rc;
    c->row_open = true;
    c->row_intra = intra != 0;
    if (intra) c->epoch = (c->epoch % 0xFFFFFEu) + 1;   // all rows of this frame share one mailbox epoch
    c->row_nref = nref_avail;
    c->row_next = 0;
    return BVC_OK;
}

extern "C" int bvc_frame_encode_row(bvc_ctx* c, int row, int qp, int64_t* row_bits) {
    if (!c) return BVC_ERR_INVALID;
    if (!c->row_open) return fail(c, BVC_ERR_INVALID, "bvc_frame_begin has not been called");
    if (row != c->row_next || row >= c->g.bh) return fail(c, BVC_ERR_INVALID, "rows must be encoded in order");
    int lg = 0; while ((1 << lg) < c->g.bs) lg++;
    if (qp < 0 || qp > lg + 7) return fail(c, BVC_ERR_INVALID, "qp > log2(block_size) + 7");
    CK(cudaSetDevice(c->device));
    const int32_t q32 = qp;
    CK(cudaMemcpyAsync(c->d_qp_rows + row, &q32, 4, cudaMemcpyHostToDevice, c->st));
    TqArgs t; PackArgs pk;
    fill_row_args(c, t, pk, c->row_intra);
    t.row_begin = row; t.row_count = 1;
    if (c->row_intra) CK(launch_tq_iframe(t, 1, c->st));
    else CK(launch_tq_pframe(t, 1, c->st));
    CK(launch_row_bits(pk, 1, row, c->d_rowbits, c->st));
    c->launches += 2;
    long long bits = 0;
    CK(cudaMemcpyAsync(&bits, c->d_rowbits, sizeof bits, cudaMemcpyDeviceToHost, c->st));
    CK(cudaStreamSynchronize(c->st));
    if (row_bits) *row_bits = bits;
    c->row_next = row + 1;
    return BVC_OK;
}

extern "C" int bvc_frame_end(bvc_ctx* c, bvc_frame_out* out) {
    if (!c) return BVC_ERR_INVALID;
    if (!c->row_open || c->row_next != c->g.bh) return fail(c, BVC_ERR_INVALID, "not every block row has been encoded");
    CK(cudaSetDevice(c->device));
    TqArgs t; PackArgs pk;
    fill_row_args(c, t, pk, c->row_intra);
    CK(launch_pack(pk, 1, c->st));
    c->launches += c->pack_tiles_n > 1 ? 3 : 2;
    c->row_open = false;
    return frame_collect(c, out, c->row_intra, c->row_nref, c->row_fl, false, nullptr, nullptr, nullptr);
}

extern "C" int bvc_interp_halfpel(bvc_ctx* c, const uint8_t* ref, uint8_t* out2x) {
    if (!c || !ref || !out2x) return BVC_ERR_INVALID;
    if (!c->p.frac_me) return fail(c, BVC_ERR_INVALID, "context was created without frac_me");
    const Geom& g = c->g;
    CK(cudaSetDevice(c->device));
    int rc;
    if ((rc = ensure_lane_desc(c, 1)) != BVC_OK) return rc;
    const int pl = ring_plane(c, 0, 0);
    if ((rc = upload_plane(c, plane_ptr(c, pl), ref)) != BVC_OK) return rc;
    std::vector<int> v{pl};
    if ((rc = upload_halfpel_desc(c, v, 0)) != BVC_OK) return rc;
    if ((rc = enqueue_halfpel(c, 0, 1)) != BVC_OK) return rc;
    struct DevTmp { void* p = nullptr; ~DevTmp() { cudaFree(p); } } tmp;   // freed on every exit path
    CK(cudaMalloc(&tmp.p, (size_t)4 * g.W * g.H));
    CK(launch_halfpel_interleave(plane_ptr(c, pl), g.W, g.H, g.pitch, g.plane_bytes, static_cast<uint8_t*>(tmp.p), c->st));
    c->launches += 1;
    CK(cudaMemcpyAsync(out2x, tmp.p, (size_t)4 * g.W * g.H, cudaMemcpyDeviceToHost, c->st));
    CK(cudaStreamSynchronize(c->st));
    return BVC_OK;
}

extern "C" int bvc_dct_quant_recon(int device, const int16_t* residual, const int16_t* pred, int nblocks, int bs, int qp,
                                   int16_t* level, uint8_t* recon, double* idct, double* coef) {
    bvc_ctx tmp;
    bvc_ctx* c = &tmp;
    if (!(bs == 4 || bs == 8 || bs == 16)) return fail(nullptr, BVC_ERR_UNSUPPORTED, "block size must be 4, 8 or 16");
    if (!residual || !pred || !level || !recon || nblocks < 1) return fail(nullptr, BVC_ERR_INVALID, "bad arguments");
    auto run = [&]() -> int {
        CK(cudaSetDevice(device));
        const size_t n = (size_t)nblocks * bs * bs;
        int16_t *dr = nullptr, *dp = nullptr, *dl = nullptr;
        uint8_t* dc = nullptr;
        double *di = nullptr, *dco = nullptr;
        struct Free6 { void **a, **b, **c, **d, **e, **f; ~Free6() { cudaFree(*a); cudaFree(*b); cudaFree(*c); cudaFree(*d); cudaFree(*e); cudaFree(*f); } }
            guard{(void**)&dr, (void**)&dp, (void**)&dl, (void**)&dc, (void**)&di, (void**)&dco};   // freed on every exit path
        CK(dalloc(&dr, n)); CK(dalloc(&dp, n)); CK(dalloc(&dl, n)); CK(dalloc(&dc, n));
        if (idct) CK(dalloc(&di, n));
        if (coef) CK(dalloc(&dco, n));
        CK(cudaMemcpy(dr, residual, n * 2, cudaMemcpyHostToDevice));
        CK(cudaMemcpy(dp, pred, n * 2, cudaMemcpyHostToDevice));
        CK(launch_tq_blocks(dr, dp, nblocks, bs, qp, dl, dc, di, dco, 0));
        CK(cudaDeviceSynchronize());
        CK(cudaMemcpy(level, dl, n * 2, cudaMemcpyDeviceToHost));
        CK(cudaMemcpy(recon, dc, n, cudaMemcpyDeviceToHost));
        if (idct) CK(cudaMemcpy(idct, di, n * 8, cudaMemcpyDeviceToHost));
        if (coef) CK(cudaMemcpy(coef, dco, n * 8, cudaMemcpyDeviceToHost));
        return BVC_OK;
    };
    int rc = run();
    if (rc != BVC_OK) g_create_error = tmp.err;
    return rc;
}


// ---------------------------------------------------------------------------------------------
// decoder: decode_video (decoder.py:26-87)
template <typename T>
static int dbuf(bvc_ctx* c, bvc_ctx::DBuf& b, size_t n, T** out) {
    const size_t bytes = n * sizeof(T) + 256;   // slack: the tokenizer stages aligned words a little past a stream's end
    if (bytes > b.cap) {
        if (b.p) CK(cudaFree(b.p));
        b.p = nullptr; b.cap = 0;
        CK(cudaMalloc(&b.p, bytes));
        b.cap = bytes;
    }
    *out = reinterpret_cast<T*>(b.p);
    return BVC_OK;
}

// Events created for the duration of one call; destroyed on every exit path.
struct EventBag {
    std::vector<cudaEvent_t> ev;
    ~EventBag() { for (auto e : ev) if (e) cudaEventDestroy(e); }
    cudaError_t make(cudaEvent_t* out, unsigned flags = cudaEventDefault) {
        cudaError_t r = cudaEventCreateWithFlags(out, flags);
        if (r == cudaSuccess) ev.push_back(*out);
        return r;
    }
};

struct DecRecord { int intra; size_t pred_off, pred_len, coef_off, coef_len; };

// init_refs / n_init: reference window the first frame sees (deque order, oldest first).  n_init < 0: the decoder's
// own start-up window, one plane filled with 128 (decoder.py:34-38).
// pred_only: stop after the prediction data (Frame.entropy_decode_prediction_data on its own).
static int decode_impl(bvc_ctx* c, const uint8_t* data, size_t len, int max_frames, const uint8_t* const* init_refs, int n_init,
                       uint8_t* frames_out, int* nframes_out, int16_t* levels_out, int32_t* pred_out, int32_t* qp_out,
                       uint8_t* kinds_out, bool pred_only = false) {
    const Geom& g = c->g;
    CK(cudaSetDevice(c->device));
    if (!data || !nframes_out || max_frames < 0) return fail(c, BVC_ERR_INVALID, "bad arguments");
    *nframes_out = 0;
    // ---- container records (decoder.py:46-69) ----
    std::vector<DecRecord> recs;
    for (size_t o = 0; o < len && (int)recs.size() < max_frames;) {
        if (o + 3 > len) return fail(c, BVC_ERR_INVALID, "truncated frame record");
        DecRecord r;
        r.intra = data[o] == 1;   // PredictionMode.INTRA_FRAME.value; anything else is decoded as a P frame
        r.pred_len = ((size_t)data[o + 1] << 8) | data[o + 2];
        r.pred_off = o + 3;
        if (r.pred_off + r.pred_len + 3 > len) return fail(c, BVC_ERR_INVALID, "truncated frame record");
        const uint8_t* q = data + r.pred_off + r.pred_len;
        r.coef_len = ((size_t)q[0] << 16) | ((size_t)q[1] << 8) | q[2];
        r.coef_off = r.pred_off + r.pred_len + 3;
        if (r.coef_off + r.coef_len > len) return fail(c, BVC_ERR_INVALID, "truncated frame record");
        o = r.coef_off + r.coef_len;
        recs.push_back(r);
    }
    const int n = (int)recs.size();
    if (n == 0) return BVC_OK;
    const size_t nb = (size_t)g.nblk;
    int rc;

    // ---- GOP lanes: a GOP starts at every I frame (the window is cleared, decoder.py:55-58) ----
    std::vector<uint8_t> intra(n);
    for (int f = 0; f < n; f++) intra[f] = (uint8_t)recs[f].intra;
    struct Gop { int first, count, virt0; };
    std::vector<Gop> gops;
    for (int f = 0; f < n; f++) {
        if (intra[f] || gops.empty()) gops.push_back({f, 0, 0});
        gops.back().count++;
    }
    const int nin = n_init < 0 ? 1 : n_init;
    if (!intra[0]) gops[0].virt0 = nin;     // leading P frames see the start-up window
    if (!intra[0] && nin < 1) return fail(c, BVC_ERR_INVALID, "P frame without a reference frame");
    if (nin > c->p.nref_frames) return fail(c, BVC_ERR_INVALID, "more initial references than nref_frames");
    const int G = c->max_lanes;
    std::vector<FrameLane> frl;
    struct DStep { size_t off_i, n_i, off_p, n_p; std::vector<int> frames, outplanes; long long chunk_begin, nchunks, bits; size_t list_off; };
    std::vector<DStep> steps;
    for (size_t g0 = 0; g0 < gops.size(); g0 += G) {
        const size_t g1 = std::min(gops.size(), g0 + G);
        int maxlen = 0;
        for (size_t gi = g0; gi < g1; gi++) maxlen = std::max(maxlen, gops[gi].count);
        for (int k = 0; k < maxlen; k++) {
            DStep st{};
            std::vector<FrameLane> li, lp;
            std::vector<int> fi, fp, oi, op;
            for (size_t gi = g0; gi < g1; gi++) {
                if (k >= gops[gi].count) continue;
                const int lane = (int)(gi - g0), f = gops[gi].first + k, vk = k + gops[gi].virt0;
                FrameLane fl{};
                fl.cur_plane = 0; fl.slot = f;
                fl.out_plane = ring_plane(c, lane, vk % c->slots);
                const int nav = intra[f] ? 0 : std::min(vk, c->p.nref_frames);
                fl.nref = nav;
                for (int j = 0; j < nav; j++) fl.ref_plane[j] = ring_plane(c, lane, (vk - nav + j) % c->slots);
                (intra[f] ? li : lp).push_back(fl);
                (intra[f] ? fi : fp).push_back(f);
                (intra[f] ? oi : op).push_back(fl.out_plane);
            }
            st.off_i = frl.size(); st.n_i = li.size();
            frl.insert(frl.end(), li.begin(), li.end());
            st.off_p = frl.size(); st.n_p = lp.size();
            frl.insert(frl.end(), lp.begin(), lp.end());
            st.frames = fi; st.frames.insert(st.frames.end(), fp.begin(), fp.end());
            st.outplanes = oi; st.outplanes.insert(st.outplanes.end(), op.begin(), op.end());
            steps.push_back(std::move(st));
        }
    }

    // ---- streams and chunk map.  Stream 2f + k = frame f's prediction data (k = 0) / coefficients (k = 1); the chunk tables
    // are laid out step by step, so the streams a decode step needs are one contiguous chunk range and can be tokenized
    // on their own, ahead of the step (the tokenizer runs on its own stream, a few steps in front of the rebuild kernels).
    const int CB = eg_chunk_bits();
    std::vector<EgStream> streams(2 * (size_t)n);
    std::vector<int> step_frames, step_streams;   // concatenated per-step lists (frames; stream ids)
    long long nchunks = 0, slab_syms = 0;
    for (auto& st : steps) {
        st.chunk_begin = nchunks; st.bits = 0; st.list_off = step_frames.size();
        for (int f : st.frames) {
            step_frames.push_back(f);
            for (int k = 0; k < 2; k++) {
                EgStream& s = streams[2 * (size_t)f + k];
                s.byte0 = (long long)(k ? recs[f].coef_off : recs[f].pred_off);
                s.nbits = 8LL * (long long)(k ? recs[f].coef_len : recs[f].pred_len);
                s.chunk0 = nchunks;
                s.sym0 = 0; s.nsym = 0; s.neob = 0; s.frame = f; s.kind = k;
                nchunks += (s.nbits + CB - 1) / CB;
                st.bits += s.nbits;
                step_streams.push_back(2 * f + k);
            }
        }
        st.nchunks = nchunks - st.chunk_begin;
        slab_syms = std::max(slab_syms, st.bits);   // a symbol is at least one bit
    }
    // symbols live in a ring of per-step slabs: a step's symbols are only read by that step's kernels
    constexpr int SLABS = 4;
    slab_syms = (slab_syms + 15) & ~15LL;
    uint8_t *d_in, *d_exit, *d_neob, *d_entry, *d_intra, *d_frame_ok;
    uint16_t* d_nsym;
    EgStream* d_streams;
    int *d_chunk_stream, *d_symbase, *d_eobbase, *d_blk_start, *d_step_frames, *d_step_streams;
    int4* d_mv;
    int32_t *d_modes, *d_qp;
    long long* d_sym0;
    int16_t *d_syms, *d_levels = nullptr;
    FrameLane* d_lanes;
    if ((rc = dbuf(c, c->dec_in, len, &d_in)) || (rc = dbuf(c, c->dec_streams, streams.size(), &d_streams)) ||
        (rc = dbuf(c, c->dec_chunk_stream, (size_t)nchunks, &d_chunk_stream)) || (rc = dbuf(c, c->dec_exit, (size_t)nchunks * 32, &d_exit)) ||
        (rc = dbuf(c, c->dec_nsym, (size_t)nchunks * 32, &d_nsym)) || (rc = dbuf(c, c->dec_neob, (size_t)nchunks * 32, &d_neob)) ||
        (rc = dbuf(c, c->dec_entry, (size_t)nchunks, &d_entry)) || (rc = dbuf(c, c->dec_symbase, (size_t)nchunks, &d_symbase)) ||
        (rc = dbuf(c, c->dec_eobbase, (size_t)nchunks, &d_eobbase)) || (rc = dbuf(c, c->dec_intra, (size_t)n, &d_intra)) ||
        (rc = dbuf(c, c->dec_mv, (size_t)n * nb, &d_mv)) || (rc = dbuf(c, c->dec_modes, (size_t)n * nb, &d_modes)) ||
        (rc = dbuf(c, c->dec_qp, (size_t)n * g.bh, &d_qp)) || (rc = dbuf(c, c->dec_blk_start, (size_t)n * (nb + 1), &d_blk_start)) ||
        (rc = dbuf(c, c->dec_sym0, (size_t)n, &d_sym0)) ||
        (rc = dbuf(c, c->dec_frame_ok, (size_t)n, &d_frame_ok)) || (rc = dbuf(c, c->dec_step_frames, step_frames.size(), &d_step_frames)) ||
        (rc = dbuf(c, c->dec_step_streams, step_streams.size(), &d_step_streams)) ||
        (rc = dbuf(c, c->dec_syms, (size_t)slab_syms * SLABS + 8, &d_syms)) || (rc = dbuf(c, c->dec_lanes, frl.size(), &d_lanes)))
        return rc;
    if (levels_out && (rc = dbuf(c, c->dec_levels, (size_t)n * g.W * g.H, &d_levels))) return rc;
    // the small tables go through one pinned staging block (truly asynchronous copies), in front of the container itself
    {
        const size_t b_streams = streams.size() * sizeof(EgStream), b_intra = ((size_t)n + 15) & ~(size_t)15,
                     b_sf = step_frames.size() * sizeof(int), b_ss = step_streams.size() * sizeof(int), b_frl = frl.size() * sizeof(FrameLane);
        if ((rc = ensure_pinned(c, &c->h_desc, &c->h_desc_cap, b_streams + b_intra + b_sf + b_ss + b_frl + 64)) != BVC_OK) return rc;
        uint8_t* h = static_cast<uint8_t*>(c->h_desc);
        memcpy(h, streams.data(), b_streams);
        CK(cudaMemcpyAsync(d_streams, h, b_streams, cudaMemcpyHostToDevice, c->st));
        h += b_streams;
        memcpy(h, intra.data(), (size_t)n);
        CK(cudaMemcpyAsync(d_intra, h, (size_t)n, cudaMemcpyHostToDevice, c->st));
        h += b_intra;
        memcpy(h, step_frames.data(), b_sf);
        CK(cudaMemcpyAsync(d_step_frames, h, b_sf, cudaMemcpyHostToDevice, c->st));
        h += b_sf;
        memcpy(h, step_streams.data(), b_ss);
        CK(cudaMemcpyAsync(d_step_streams, h, b_ss, cudaMemcpyHostToDevice, c->st));
        h += b_ss;
        memcpy(h, frl.data(), b_frl);
        CK(cudaMemcpyAsync(d_lanes, h, b_frl, cudaMemcpyHostToDevice, c->st));
    }
    CK(launch_eg_chunk_map(d_streams, (int)streams.size(), d_chunk_stream, c->st));
    CK(cudaMemsetAsync(d_frame_ok, 0, (size_t)n, c->st));
    CK(cudaMemsetAsync(c->d_overflow, 0, sizeof(int), c->st));   // reused as the decoder's error flag
    c->launches += 1;
    // The container goes up on the copy stream in two phases: first the records of the frames the first step rebuilds (the
    // I frame of every GOP lane: a few hundred KB each), then everything else -- the first step's tokenizing and wavefront
    // then run under the upload of the other 90+ % (1.4 ms for the 76 MB headline container).
    EventBag bag;
    cudaEvent_t ev_in0, ev_in1;
    CK(bag.make(&ev_in0, cudaEventDisableTiming));
    CK(bag.make(&ev_in1, cudaEventDisableTiming));
    {
        CK(cudaMemsetAsync(d_in + len, 0, 256, c->st_h2d));
        std::vector<std::pair<size_t, size_t>> first;   // [begin, end) of the first step's records, ascending
        if (!steps.empty())
            for (int f : steps[0].frames) first.emplace_back(recs[f].pred_off - 3, recs[f].coef_off + recs[f].coef_len);
        std::sort(first.begin(), first.end());
        for (auto& r : first) CK(cudaMemcpyAsync(d_in + r.first, data + r.first, r.second - r.first, cudaMemcpyHostToDevice, c->st_h2d));
        CK(cudaEventRecord(ev_in0, c->st_h2d));
        size_t at = 0;
        for (auto& r : first) {
            if (r.first > at) CK(cudaMemcpyAsync(d_in + at, data + at, r.first - at, cudaMemcpyHostToDevice, c->st_h2d));
            at = std::max(at, r.second);
        }
        if (len > at) CK(cudaMemcpyAsync(d_in + at, data + at, len - at, cudaMemcpyHostToDevice, c->st_h2d));
        CK(cudaEventRecord(ev_in1, c->st_h2d));
    }
    if (c->p.frac_me) {
        if ((rc = ensure_lane_desc(c, frl.size() + (size_t)nin)) != BVC_OK) return rc;
        for (auto& st : steps)
            if ((rc = upload_halfpel_desc(c, st.outplanes, st.off_i)) != BVC_OK) return rc;
    }
    // start-up window of lane 0
    if (gops[0].virt0) {
        std::vector<int> hp;
        for (int j = 0; j < nin; j++) {
            const int pl = ring_plane(c, 0, j);
            if (n_init < 0) CK(launch_fill_plane(plane_ptr(c, pl), g.plane_bytes, 128, c->st));
            else if ((rc = upload_plane(c, plane_ptr(c, pl), init_refs[j])) != BVC_OK) return rc;
            hp.push_back(pl);
        }
        if (c->p.frac_me) {
            if ((rc = upload_halfpel_desc(c, hp, frl.size())) != BVC_OK) return rc;
            if ((rc = enqueue_halfpel(c, frl.size(), nin)) != BVC_OK) return rc;
        }
    }
    cudaEvent_t ev_up;
    CK(bag.make(&ev_up, cudaEventDisableTiming));
    CK(cudaEventRecord(ev_up, c->st));
    // ---- D1-D4 per step on the tokenizer stream: speculative walk, chain, symbol offsets (on the device: no host round
    // trip), symbols + block starts, prediction data ----
    cudaStream_t st_tok = c->st_grp[0];
    CK(cudaStreamWaitEvent(st_tok, ev_up, 0));
    CK(cudaStreamWaitEvent(st_tok, ev_in0, 0));
    std::vector<cudaEvent_t> ev_tok(steps.size(), nullptr), ev_dec(steps.size(), nullptr);
    // The speculative walk and the chain pass of a stream do not touch the symbol ring, and the chain is one serial walk
    // per stream (0.1-0.3 ms whatever the number of streams), so they run for batches of steps -- 1, 2, 4, 8, ... steps: the
    // first step's streams are ready after one short batch, the later batches stay far ahead of the rebuild kernels.
    size_t chained = 0;   // steps [0, chained) have had their chain pass enqueued
    auto enqueue_tokenize = [&](size_t si) -> int {
        if (si >= steps.size()) return BVC_OK;
        if (si >= chained) {
            if (chained == 1) CK(cudaStreamWaitEvent(st_tok, ev_in1, 0));   // everything after the first step needs the whole container
            const size_t s0 = chained, s1 = std::min(steps.size(), s0 + std::max<size_t>(1, s0));   // batch sizes 1, 1, 2, 4, 8, ...
            long long nch = 0;
            size_t nfr = 0;
            for (size_t k = s0; k < s1; k++) { nch += steps[k].nchunks; nfr += steps[k].frames.size(); }
            CK(launch_eg_tokenize_spec(d_in, d_streams, d_step_streams + 2 * steps[s0].list_off, (int)(2 * nfr), d_chunk_stream, steps[s0].chunk_begin,
                                       nch, d_exit, d_nsym, d_neob, d_entry, d_symbase, d_eobbase, c->d_overflow, st_tok));
            c->launches += 2;
            chained = s1;
        }
        const DStep& st = steps[si];
        const int nf = (int)st.frames.size();
        if (si >= (size_t)SLABS && ev_dec[si - SLABS]) CK(cudaStreamWaitEvent(st_tok, ev_dec[si - SLABS], 0));   // the slab's previous tenant has been rebuilt
        CK(launch_eg_offsets(d_streams, d_step_streams + 2 * st.list_off, 2 * nf, (long long)(si % SLABS) * slab_syms, d_sym0, d_frame_ok, g.nblk,
                             pred_only ? 1 : 0, c->d_overflow, st_tok));
        CK(launch_eg_tokenize_emit(d_in, d_streams, d_chunk_stream, st.chunk_begin, st.nchunks, d_entry, d_symbase, d_eobbase, d_syms, d_blk_start,
                                   g.nblk, st_tok));
        CK(launch_pred_decode(d_streams, d_syms, d_intra, d_step_frames + st.list_off, nf, d_mv, d_modes, d_qp, g.bw, g.bh, c->p.qp,
                              c->p.nref_frames > 1, c->d_overflow, st_tok));
        c->launches += 3;
        CK(bag.make(&ev_tok[si], cudaEventDisableTiming));
        CK(cudaEventRecord(ev_tok[si], st_tok));
        return BVC_OK;
    };
    for (size_t si = 0; si < (size_t)SLABS - 1; si++)
        if ((rc = enqueue_tokenize(si)) != BVC_OK) return rc;
    int err = 0;
    // ---- D5/D6 step by step ----
    DecArgs a{};
    a.ref_base = c->ref_pool; a.ref_plane_bytes = g.plane_bytes; a.ref_pitch = g.pitch;
    a.mv_all = d_mv; a.modes_all = d_modes; a.qp_all = d_qp; a.syms = d_syms; a.coef_sym0 = d_sym0; a.blk_start = d_blk_start;
    a.levels_out = d_levels; a.ticket = c->d_ticket; a.err_flag = c->d_overflow;
    a.W = g.W; a.H = g.H; a.bs = g.bs; a.bw = g.bw; a.bh = g.bh; a.nblk = g.nblk; a.frac = c->p.frac_me;
    a.frame_ok = d_frame_ok; a.top_mail = c->d_top_mail;
    // Decoded planes go back on their own stream so that the download of step s overlaps the kernels of step s+1 (the
    // 1.25 GB of planes of the headline clip are the decoder's bound).  A plane of the reconstruction ring is rewritten
    // `slots` frames later: the kernels that rewrite it wait for its pending download.
    // The rebuild kernels run on a high-priority stream: the tokenizer's batches (low priority, many CTAs) would otherwise
    // hold the GPU while a step waits -- the block scheduler does not place a second kernel beside one that still has CTAs
    // to dispatch, but it does hand freed slots to the higher priority first.
    cudaStream_t sd = c->st_post[0];
    CK(cudaStreamWaitEvent(sd, ev_up, 0));
    std::vector<cudaEvent_t> ev_copied(frames_out ? steps.size() : 0);
    std::vector<int> pending_copy(c->ref_planes, -1);   // plane -> step whose download still reads it
    for (size_t si = 0; si < steps.size(); si++) {
        auto& st = steps[si];
        if ((rc = enqueue_tokenize(si + SLABS - 1)) != BVC_OK) return rc;   // the tokenizer stays SLABS - 1 steps ahead
        if (pred_only) continue;                                            // prediction data only: nothing to rebuild
        CK(cudaStreamWaitEvent(sd, ev_tok[si], 0));
        if (frames_out) {
            int wait_for = -1;
            for (int pl : st.outplanes) wait_for = std::max(wait_for, pending_copy[pl]);
            if (wait_for >= 0) CK(cudaStreamWaitEvent(sd, ev_copied[wait_for], 0));
        }
        if (st.n_i) {
            c->epoch = (c->epoch % 0xFFFFFEu) + 1;   // one mailbox epoch per I step
            a.epoch = c->epoch;
            a.lanes = d_lanes + st.off_i;
            CK(launch_dec_iframe(a, (int)st.n_i, sd));
            c->launches += 1;
        }
        if (st.n_p) {
            a.lanes = d_lanes + st.off_p;
            CK(launch_dec_pframe(a, (int)st.n_p, sd));
            c->launches += 1;
        }
        if ((rc = enqueue_halfpel(c, st.off_i, (int)(st.n_i + st.n_p), sd)) != BVC_OK) return rc;
        CK(bag.make(&ev_dec[si], cudaEventDisableTiming));
        CK(cudaEventRecord(ev_dec[si], sd));
        if (frames_out) {
            CK(bag.make(&ev_copied[si], cudaEventDisableTiming));
            CK(cudaStreamWaitEvent(c->st_d2h, ev_dec[si], 0));
            // the planes of a step are equally spaced on both sides when the step's frames are (frame k of consecutive
            // GOPs, lanes side by side in the reconstruction ring): one strided copy, "row" = one plane
            const size_t nfr = st.frames.size(), fbytes = (size_t)g.W * g.H;
            bool strided = g.pitch == g.W && nfr > 2;
            long long hstride = 0, dstride = 0;
            if (strided) {
                hstride = (long long)st.frames[1] - st.frames[0];
                dstride = (long long)st.outplanes[1] - st.outplanes[0];
                for (size_t l = 2; l < nfr && strided; l++)
                    strided = (long long)st.frames[l] - st.frames[l - 1] == hstride && (long long)st.outplanes[l] - st.outplanes[l - 1] == dstride;
                strided = strided && hstride > 0 && dstride > 0 && (unsigned long long)hstride * fbytes <= 0x7fffffffull &&
                          (unsigned long long)dstride * g.plane_bytes <= 0x7fffffffull;
            }
            if (strided) {
                CK(cudaMemcpy2DAsync(frames_out + (size_t)st.frames[0] * fbytes, (size_t)hstride * fbytes, plane_ptr(c, st.outplanes[0]),
                                     (size_t)dstride * g.plane_bytes, fbytes, nfr, cudaMemcpyDeviceToHost, c->st_d2h));
            } else {
                for (size_t l = 0; l < nfr; l++)
                    if ((rc = download_plane(c, frames_out + (size_t)st.frames[l] * fbytes, plane_ptr(c, st.outplanes[l]), c->st_d2h)) != BVC_OK)
                        return rc;
            }
            for (size_t l = 0; l < nfr; l++) pending_copy[st.outplanes[l]] = (int)si;
            CK(cudaEventRecord(ev_copied[si], c->st_d2h));
        }
    }
    if (!pred_only && !steps.empty()) CK(cudaStreamWaitEvent(c->st, ev_dec[steps.size() - 1], 0));
    if (frames_out && !steps.empty() && !pred_only) CK(cudaStreamWaitEvent(c->st, ev_copied[steps.size() - 1], 0));
    if (!steps.empty()) CK(cudaStreamWaitEvent(c->st, ev_tok[steps.size() - 1], 0));
    CK(cudaStreamWaitEvent(c->st, ev_in1, 0));   // a one-step clip never waited for the second phase of the upload
    CK(cudaMemcpyAsync(&err, c->d_overflow, sizeof err, cudaMemcpyDeviceToHost, c->st));
    if (levels_out) CK(cudaMemcpyAsync(levels_out, d_levels, (size_t)n * g.W * g.H * sizeof(int16_t), cudaMemcpyDeviceToHost, c->st));
    std::vector<int4> hmv;
    std::vector<int32_t> hmodes;
    if (pred_out) {
        hmv.resize((size_t)n * nb); hmodes.resize((size_t)n * nb);
        CK(cudaMemcpyAsync(hmv.data(), d_mv, hmv.size() * sizeof(int4), cudaMemcpyDeviceToHost, c->st));
        CK(cudaMemcpyAsync(hmodes.data(), d_modes, hmodes.size() * sizeof(int32_t), cudaMemcpyDeviceToHost, c->st));
    }
    if (qp_out) CK(cudaMemcpyAsync(qp_out, d_qp, (size_t)n * g.bh * sizeof(int32_t), cudaMemcpyDeviceToHost, c->st));
    CK(cudaStreamSynchronize(c->st));
    if (err) return fail(c, BVC_ERR_INVALID, "malformed stream (exp-Golomb code cut short or too long, a coefficient stream without one EOB-terminated run "
                               "per block, missing prediction symbols, bad intra mode or motion vector out of range)");
    if (pred_out)
        for (int f = 0; f < n; f++)
            for (size_t b = 0; b < nb; b++) {
                int32_t* o = pred_out + ((size_t)f * nb + b) * 3;
                if (intra[f]) { o[0] = hmodes[(size_t)f * nb + b]; o[1] = 0; o[2] = 0; }
                else { const int4 m = hmv[(size_t)f * nb + b]; o[0] = m.x; o[1] = m.y; o[2] = m.z; }
            }
    if (kinds_out) memcpy(kinds_out, intra.data(), (size_t)n);
    *nframes_out = n;
    return BVC_OK;
}

extern "C" int bvc_decode_clip(bvc_ctx* c, const uint8_t* data, size_t len, int max_frames, uint8_t* frames_out, int* nframes_out,
                               int16_t* levels_out, int32_t* pred_out, int32_t* qp_rows_out, uint8_t* kinds_out) {
    if (!c) return BVC_ERR_INVALID;
    return decode_impl(c, data, len, max_frames, nullptr, -1, frames_out, nframes_out, levels_out, pred_out, qp_rows_out, kinds_out);
}

extern "C" int bvc_decode_frame(bvc_ctx* c, int intra, const uint8_t* pred, size_t pred_len, const uint8_t* coef, size_t coef_len,
                                const uint8_t* const* refs, int nref_avail, uint8_t* recon, int16_t* levels, int32_t* pred_out,
                                int32_t* qp_rows_out) {
    if (!c) return BVC_ERR_INVALID;
    if (pred_len > 0xFFFF || coef_len > 0xFFFFFF || (pred_len && !pred) || (coef_len && !coef)) return fail(c, BVC_ERR_INVALID, "bad payload");
    const bool pred_only = coef == nullptr;
    if (pred_only) { coef_len = 0; refs = nullptr; nref_avail = intra ? 0 : 1; recon = nullptr; levels = nullptr; }
    else if (!intra && (nref_avail < 1 || !refs)) return fail(c, BVC_ERR_INVALID, "nref_avail must be 1..nref_frames");
    std::vector<uint8_t> rec(6 + pred_len + coef_len);
    rec[0] = intra ? 1 : 0;
    rec[1] = (uint8_t)(pred_len >> 8); rec[2] = (uint8_t)pred_len;
    if (pred_len) memcpy(&rec[3], pred, pred_len);
    rec[3 + pred_len] = (uint8_t)(coef_len >> 16); rec[4 + pred_len] = (uint8_t)(coef_len >> 8); rec[5 + pred_len] = (uint8_t)coef_len;
    if (coef_len) memcpy(&rec[6 + pred_len], coef, coef_len);
    int n = 0;
    return decode_impl(c, rec.data(), rec.size(), 1, refs, intra ? 0 : (pred_only ? -1 : nref_avail), recon, &n, levels, pred_out,
                       qp_rows_out, nullptr, pred_only);
}

// ---------------------------------------------------------------------------------------------
// clip-level API
extern "C" int bvc_clip_upload(bvc_ctx* c, const uint8_t* frames, int nframes) {
    if (!c || !frames || nframes < 1) return BVC_ERR_INVALID;
    CK(cudaSetDevice(c->device));
    int rc;
    if ((rc = ensure_in_pool(c, (size_t)nframes)) != BVC_OK) return rc;
    const Geom& g = c->g;
    if (g.pitch == g.W && g.plane_bytes == (size_t)g.W * g.H) {
        CK(cudaMemcpyAsync(c->in_pool, frames, (size_t)nframes * g.plane_bytes, cudaMemcpyHostToDevice, c->st));
    } else {
        for (int f = 0; f < nframes; f++)
            CK(cudaMemcpy2DAsync(c->in_pool + (size_t)f * g.plane_bytes, g.pitch, frames + (size_t)f * g.W * g.H, g.W, g.W, g.H,
                                 cudaMemcpyHostToDevice, c->st));
    }
    CK(cudaStreamSynchronize(c->st));
    c->resident_frames = nframes;
    return BVC_OK;
}

// Input stage (assign1/ex2.py:14-46 read_y_component + common.pad_frame common.py:22-32): the luma planes of an I420
// (YUV 4:2:0 planar) file image go straight to HBM -- a strided copy skips the chroma planes, no host-side repacking --
// and frames whose size is not a multiple of the block size are padded bottom / right with 128 on the device.
extern "C" int bvc_clip_upload_i420(bvc_ctx* c, const uint8_t* yuv, int src_w, int src_h, int nframes) {
    if (!c || !yuv || nframes < 1) return BVC_ERR_INVALID;
    const Geom& g = c->g;
    if (src_w < 1 || src_h < 1 || src_w > g.W || src_h > g.H || g.W - src_w >= g.bs || g.H - src_h >= g.bs)
        return fail(c, BVC_ERR_INVALID, "context size must be the source size rounded up to the block size");
    CK(cudaSetDevice(c->device));
    int rc;
    if ((rc = ensure_in_pool(c, (size_t)nframes)) != BVC_OK) return rc;
    const size_t ysz = (size_t)src_w * src_h, fsz = ysz + 2 * ((size_t)(src_w / 2) * (src_h / 2));
    if (src_w == g.W && src_h == g.H && g.pitch == g.W) {
        // one 2-D copy: "row" f = the luma plane of frame f, source pitch = a whole I420 frame
        CK(cudaMemcpy2DAsync(c->in_pool, g.plane_bytes, yuv, fsz, ysz, (size_t)nframes, cudaMemcpyHostToDevice, c->st));
    } else {
        if (src_w != g.W || src_h != g.H) CK(launch_fill_plane(c->in_pool, (size_t)nframes * g.plane_bytes, 128, c->st));
        for (int f = 0; f < nframes; f++)
            CK(cudaMemcpy2DAsync(c->in_pool + (size_t)f * g.plane_bytes, g.pitch, yuv + (size_t)f * fsz, src_w, src_w, src_h,
                                 cudaMemcpyHostToDevice, c->st));
    }
    CK(cudaStreamSynchronize(c->st));
    c->resident_frames = nframes;
    return BVC_OK;
}

// Clip encoder.  GOPs are encoded max_lanes at a time ("waves" of max_lanes x I_Period frames, frame k of every GOP of the
// wave per step).  Nothing in here scales with the length of the clip except tiny per-frame descriptors:
//   * input planes live in a ring of IN_RING_STEPS steps x max_lanes planes, refilled three steps ahead on the copy
//     stream (a slot is reused once every lane group has finished the step that read it);
//   * the per-frame bit streams live in one wave's worth of slots; at the end of a wave its container fragment is laid
//     out on the device into one of two staging buffers and goes to the caller's buffer on the download stream while
//     the next wave is already running -- the host only waits for a wave's size after it has enqueued the wave after it.
// out == nullptr (keep_on_device): the fragments are appended to c->d_container instead (bvc_encode_clip_device).
static constexpr int IN_RING_STEPS = 6;

static int encode_clip_impl(bvc_ctx* c, const uint8_t* host_frames, int nframes, uint8_t* out, size_t out_cap, size_t* out_len,
                            uint8_t* recon, bool keep_on_device = false) {
    const Geom& g = c->g;
    CK(cudaSetDevice(c->device));
    c->container_len = 0;
    if (nframes < 1 || (!out && !keep_on_device) || !out_len) return fail(c, BVC_ERR_INVALID, "bad arguments");
    const int IP = c->p.i_period, G = c->max_lanes, D = IN_RING_STEPS;
    const int ngop = (nframes + IP - 1) / IP;
    const int nwaves = (ngop + G - 1) / G;
    const size_t wave_frames = (size_t)std::min((long long)nframes, (long long)G * IP);
    int rc;
    if ((rc = ensure_in_pool(c, host_frames ? (size_t)D * G : (size_t)nframes)) != BVC_OK) return rc;
    if (host_frames) c->resident_frames = 0;   // the pool is about to be overwritten with this call's frames
    if ((rc = ensure_streams(c, wave_frames)) != BVC_OK) return rc;
    // staging for one wave's container fragment (x2: the download of wave w overlaps the assembly of wave w+1)
    // sized for 1 bit per pixel (the headline workload codes 0.5) unless an earlier call found that too small
    const size_t frag_worst = wave_frames * (6 + 4 * (c->coef_cap_words + c->pred_cap_words));
    // the staging of a wave is cut into one region per container part (lane group, see below); no region needs to hold
    // more than the caller's buffer does
    const int NG = std::max(1, std::min(c->ngroups, G));
    const size_t frag_total = std::min(frag_worst, std::max(c->frag_min, wave_frames * ((size_t)g.W * g.H / 8)));
    const size_t region = std::min((out_cap + 255) & ~(size_t)255, (frag_total / NG + ((size_t)1 << 20) + 256) & ~(size_t)255);
    if ((rc = ensure_fragments(c, region * NG, nwaves > 1 ? 2 : 1)) != BVC_OK) return rc;
    if (keep_on_device && (rc = ensure_container(c, out_cap)) != BVC_OK) return rc;
    const int per = (G + NG - 1) / NG;
    // A wave's container is laid out in NG parts, one per lane group (lanes are GOPs, a group's GOPs are consecutive in the
    // stream): a part is assembled on its group's assembly stream right behind the group's last frame and sent off as soon
    // as its size is known, so the first group's part travels while the last group still searches.
    const size_t nparts = (size_t)nwaves * NG;
    if ((rc = ensure_pinned(c, &c->h_totals, &c->h_totals_cap, (nparts + 1) * 16)) != BVC_OK) return rc;
    long long* h_total = static_cast<long long*>(c->h_totals);          // [nwaves][NG] part bytes
    int* h_over = reinterpret_cast<int*>(h_total + nparts);             // [nwaves][NG] overflow flags (2 ints per part)

    // ---- plan: one step per (wave, k); lane l of a step = GOP (wave*G + l) ----
    std::vector<StepPlan> steps;
    std::vector<MeLane> mel;
    std::vector<FrameLane> frl;
    std::vector<std::vector<int>> step_frames;   // clip frame index of every lane of a step
    std::vector<std::vector<int>> step_outplane;
    std::vector<int> step_wave;
    std::vector<int> wave_nframes(nwaves, 0);
    for (int w = 0; w < nwaves; w++) {
        const int g0 = w * G, g1 = std::min(ngop, g0 + G);
        const int wave_first = g0 * IP;
        wave_nframes[w] = std::min(nframes, g1 * IP) - wave_first;
        for (int k = 0; k < IP; k++) {
            StepPlan sp;
            sp.intra = (k == 0);
            sp.nref = std::min(k, c->p.nref_frames);
            sp.k = k;
            sp.desc_off = mel.size();
            const int sidx = (int)steps.size();          // global step index: the input ring slot is sidx % D
            std::vector<int> fr, op;
            for (int gi = g0; gi < g1; gi++) {
                const int f = gi * IP + k;
                if (f >= nframes) continue;
                const int lane = gi - g0;
                MeLane ml{};
                FrameLane fl{};
                ml.cur_plane = fl.cur_plane = host_frames ? (sidx % D) * G + lane : f;
                fl.slot = f - wave_first;
                const int nav = std::min(k, c->p.nref_frames);  // deque(maxlen=nRef), cleared at the I frame
                ml.nref = fl.nref = nav;
                for (int j = 0; j < nav; j++) {
                    const int src_k = k - nav + j;              // oldest first (encoder.py:33,154)
                    ml.ref_plane[j] = fl.ref_plane[j] = ring_plane(c, lane, src_k % c->slots);
                }
                fl.out_plane = ring_plane(c, lane, k % c->slots);
                mel.push_back(ml);
                frl.push_back(fl);
                fr.push_back(f);
                op.push_back(fl.out_plane);
            }
            sp.nl = (int)fr.size();
            if (sp.nl == 0) continue;
            steps.push_back(sp);
            step_frames.push_back(fr);
            step_outplane.push_back(op);
            step_wave.push_back(w);
        }
    }
    const size_t nsteps = steps.size();
    if ((rc = ensure_lane_desc(c, mel.size())) != BVC_OK) return rc;
    if ((rc = ensure_pinned(c, &c->h_desc, &c->h_desc_cap, mel.size() * (sizeof(MeLane) + sizeof(FrameLane)) + 64)) != BVC_OK) return rc;
    memcpy(c->h_desc, mel.data(), mel.size() * sizeof(MeLane));
    memcpy((uint8_t*)c->h_desc + mel.size() * sizeof(MeLane), frl.data(), frl.size() * sizeof(FrameLane));
    CK(cudaMemcpyAsync(c->d_me_lanes, c->h_desc, mel.size() * sizeof(MeLane), cudaMemcpyHostToDevice, c->st));
    CK(cudaMemcpyAsync(c->d_fr_lanes, (uint8_t*)c->h_desc + mel.size() * sizeof(MeLane), frl.size() * sizeof(FrameLane),
                       cudaMemcpyHostToDevice, c->st));
    if (c->p.frac_me) {
        for (size_t s = 0; s < nsteps; s++)
            if ((rc = upload_halfpel_desc(c, step_outplane[s], steps[s].desc_off)) != BVC_OK) return rc;
    }
    CK(cudaMemsetAsync(c->d_overflow, 0, 2 * sizeof(int), c->st));

    c->ev_used = 0;
    c->spans.clear();
    EventBag bag;
    cudaEvent_t ev_clip0, ev_clip1;
    CK(bag.make(&ev_clip0));
    CK(bag.make(&ev_clip1));
    CK(cudaEventRecord(ev_clip0, c->st));

    // "step s is finished by group gi" (its search and transform have read the input planes of ring slot s % D)
    std::vector<cudaEvent_t> ev_step((size_t)D * NG, nullptr);
    for (auto& e : ev_step) CK(bag.make(&e, cudaEventDisableTiming));
    // ---- input: uploaded step by step on its own stream so the copies overlap compute ----
    std::vector<cudaEvent_t> ev_h2d(host_frames ? nsteps * NG : 0, nullptr);
    auto enqueue_upload = [&](size_t s) -> int {
        if (!host_frames || s >= nsteps) return BVC_OK;
        const std::vector<int>& fr = step_frames[s];
        const size_t fbytes = (size_t)g.W * g.H, spitch = (size_t)IP * fbytes;
        // one copy per lane group, in group order: a group starts on a step as soon as ITS planes are there
        for (int gi = 0; gi < NG; gi++) {
            const int l0 = gi * per, nl = std::min((int)fr.size(), l0 + per) - l0;
            if (nl <= 0) continue;
            if (s >= (size_t)D) CK(cudaStreamWaitEvent(c->st_h2d, ev_step[(s % D) * NG + gi], 0));   // the slot's previous tenant: step s - D
            uint8_t* dst = c->in_pool + ((size_t)(s % D) * G + l0) * g.plane_bytes;   // lanes of a step sit side by side in the ring
            if (g.pitch == g.W && nl > 1 && spitch <= 0x7fffffffull && g.plane_bytes <= 0x7fffffffull) {
                // the frames of a step are IP apart (frame k of consecutive GOPs): one strided copy, "row" = one plane
                CK(cudaMemcpy2DAsync(dst, g.plane_bytes, host_frames + (size_t)fr[l0] * fbytes, spitch, fbytes, nl,
                                     cudaMemcpyHostToDevice, c->st_h2d));
            } else {
                for (int l = 0; l < nl; l++)
                    CK(cudaMemcpy2DAsync(dst + l * g.plane_bytes, g.pitch, host_frames + (size_t)fr[l0 + l] * fbytes, g.W, g.W, g.H,
                                         cudaMemcpyHostToDevice, c->st_h2d));
            }
            CK(bag.make(&ev_h2d[s * NG + gi], cudaEventDisableTiming));
            CK(cudaEventRecord(ev_h2d[s * NG + gi], c->st_h2d));
        }
        return BVC_OK;
    };
    if (host_frames) {
        CK(cudaStreamWaitEvent(c->st_h2d, ev_clip0, 0));
        for (size_t s = 0; s < 3; s++)
            if ((rc = enqueue_upload(s)) != BVC_OK) return rc;
    }

    // Lane groups: group gi owns lanes [gi*per, (gi+1)*per) of every step and three streams.  Inside a group
    // the order is ME(k) -> [TQ, half-pel, recon download](k) -> ME(k+1), stream assembly on the side; across groups
    // there is no dependency, so the post-ME kernels of one group run while the other groups search.
    // per-kernel event pairs only make sense when the kernels of a step run back to back on one stream; with lane groups
    // they would just be ~500 extra event records per clip
    struct TimingGuard { bvc_ctx* c; bool saved; ~TimingGuard() { c->timing = saved; } } timing_guard{c, c->timing};
    c->timing = c->timing && NG == 1;
    for (int gi = 0; gi < NG; gi++) {
        CK(cudaStreamWaitEvent(c->st_grp[gi], ev_clip0, 0));
        CK(cudaStreamWaitEvent(c->st_post[gi], ev_clip0, 0));
        CK(cudaStreamWaitEvent(c->st_pack[gi], ev_clip0, 0));
        CK(cudaEventRecord(c->ev_pack[gi], c->st_pack[gi]));   // "no assembly pending" for the first step of this call
    }
    CK(cudaStreamWaitEvent(c->st_d2h, ev_clip0, 0));

    // ---- end of a wave: container parts (encoder.py:104-121) on the device, sizes to the host ----
    std::vector<cudaEvent_t> ev_total(nparts, nullptr), ev_frag_free(nparts, nullptr);
    auto part_stream = [&](int gi) -> cudaStream_t { return NG == 1 ? c->st : c->st_pack[gi]; };
    auto part_frames = [&](int w, int gi) -> int {
        const long long n = std::min((long long)wave_nframes[w] - (long long)gi * per * IP, (long long)per * IP);
        return n > 0 ? (int)n : 0;
    };
    size_t out_off = 0;
    int flushed = 0;       // waves whose parts have been sent on their way
    auto finish_part = [&](int w, int gi) -> int {
        const int n = part_frames(w, gi);
        const size_t pi = (size_t)w * NG + gi;
        h_total[pi] = 0; h_over[2 * pi] = h_over[2 * pi + 1] = 0;
        if (n == 0) return BVC_OK;
        cudaStream_t ps = part_stream(gi);   // behind the group's last stream assembly; the next wave's assembly follows on the same stream
        if (w >= 2 && ev_frag_free[pi - 2 * NG]) CK(cudaStreamWaitEvent(ps, ev_frag_free[pi - 2 * NG], 0));   // staging buffer w & 1 has been copied out
        const size_t first = (size_t)gi * per * IP;   // first frame slot of the part (a multiple of I_Period)
        ContainerArgs ca{};
        ca.frame_bits = c->d_frame_bits + 2 * first;
        ca.coef_stream = c->d_coef_stream + first * c->coef_cap_words; ca.pred_stream = c->d_pred_stream + first * c->pred_cap_words;
        ca.coef_cap_words = c->coef_cap_words; ca.pred_cap_words = c->pred_cap_words;
        ca.frame_off = c->d_frame_off + first + gi; ca.overflow = c->d_overflow;
        ca.out = c->d_frag[w & 1] + (size_t)gi * region; ca.out_cap = (long long)region;
        ca.nframes = n; ca.i_period = IP;
        const int ec0 = tick(c, ps);
        CK(launch_container(ca, ps));
        span(c, BVC_K_PACK, ec0, tick(c, ps));
        c->launches += 2;
        CK(cudaMemcpyAsync(&h_total[pi], ca.frame_off + n, sizeof(long long), cudaMemcpyDeviceToHost, ps));
        CK(cudaMemcpyAsync(&h_over[2 * pi], c->d_overflow, 2 * sizeof(int), cudaMemcpyDeviceToHost, ps));
        CK(bag.make(&ev_total[pi], cudaEventDisableTiming));
        CK(cudaEventRecord(ev_total[pi], ps));
        return BVC_OK;
    };
    // host side of a finished part: wait for its size, then send it to its place
    auto flush_part = [&](int w, int gi) -> int {
        const size_t pi = (size_t)w * NG + gi;
        if (!ev_total[pi]) return BVC_OK;   // no frames in this part
        CK(cudaEventSynchronize(ev_total[pi]));
        if (h_over[2 * pi]) return fail(c, BVC_ERR_OVERFLOW, "payload length does not fit the container field");
        if (h_over[2 * pi + 1]) return fail(c, BVC_ERR_NOMEM, "a frame's bit stream does not fit its device slot: raise it with bvc_set_stream_slot_bytes");
        const size_t total = (size_t)h_total[pi];
        if (out_off + total > out_cap) {
            // report what the whole clip needs, as far as it is known: at least this much
            *out_len = std::max(out_off + total, out_cap + 1);
            return fail(c, BVC_ERR_NOMEM, "output buffer too small (*out_len = a lower bound of the bytes needed)");
        }
        if (total > region) {   // the part did not fit its staging region: remember, the caller repeats the call
            c->frag_min = std::max(c->frag_min, (size_t)NG * (total + total / 4));
            *out_len = 0;
            return fail(c, BVC_ERR_NOMEM, "container staging buffer too small for this content: it has been enlarged, repeat the call");
        }
        CK(cudaStreamWaitEvent(c->st_d2h, ev_total[pi], 0));
        const uint8_t* src = c->d_frag[w & 1] + (size_t)gi * region;
        if (keep_on_device) CK(cudaMemcpyAsync(c->d_container + out_off, src, total, cudaMemcpyDeviceToDevice, c->st_d2h));
        else CK(cudaMemcpyAsync(out + out_off, src, total, cudaMemcpyDeviceToHost, c->st_d2h));
        CK(bag.make(&ev_frag_free[pi], cudaEventDisableTiming));
        CK(cudaEventRecord(ev_frag_free[pi], c->st_d2h));
        out_off += total;
        return BVC_OK;
    };

    // ---- the clip is enqueued wave after wave; the host is always at least one whole wave ahead of the GPU ----
    for (size_t s = 0; s < nsteps; s++) {
        const int w = step_wave[s];
        const bool wave_start = s == 0 || step_wave[s - 1] != w;
        if (wave_start && w >= 2) {   // wave w-2 has long finished: send its parts off (this is the only host wait, a wave behind)
            for (int gi = 0; gi < NG; gi++)
                if ((rc = flush_part(w - 2, gi)) != BVC_OK) return rc;
            flushed = w - 1;
        }
        if (host_frames)
            if ((rc = enqueue_upload(s + 3)) != BVC_OK) return rc;
        for (int gi = 0; gi < NG; gi++) {
            const int l0 = gi * per, nl = std::min(steps[s].nl, l0 + per) - l0;
            cudaStream_t sm = NG == 1 ? c->st : c->st_grp[gi], spst = NG == 1 ? c->st : c->st_post[gi];
            if (nl > 0) {
                if (host_frames) {
                    CK(cudaStreamWaitEvent(sm, ev_h2d[s * NG + gi], 0));
                    if (spst != sm) CK(cudaStreamWaitEvent(spst, ev_h2d[s * NG + gi], 0));
                }
                if ((rc = enqueue_step(c, steps[s], false, sm, spst, l0, nl, c->ev_me[gi], NG == 1 ? nullptr : c->st_pack[gi], c->ev_tq[gi],
                                       c->ev_pack[gi])) != BVC_OK) return rc;
                // phase planes of the new reconstructions (build_pre_interpolated_buffer, encoder.py:155)
                if ((rc = enqueue_halfpel(c, steps[s].desc_off + l0, nl, spst)) != BVC_OK) return rc;
                if (recon) {
                    for (int l = l0; l < l0 + nl; l++)
                        if ((rc = download_plane(c, recon + (size_t)step_frames[s][l] * g.W * g.H, plane_ptr(c, step_outplane[s][l]), spst)) != BVC_OK)
                            return rc;
                }
                if (spst != sm) {   // the next motion search of this group needs this step's reconstructions
                    CK(cudaEventRecord(c->ev_post[gi], spst));
                    CK(cudaStreamWaitEvent(sm, c->ev_post[gi], 0));
                }
            }
            if (host_frames) CK(cudaEventRecord(ev_step[(s % D) * NG + gi], spst));   // this group is done with ring slot s % D
        }
        if (s + 1 == nsteps || step_wave[s + 1] != w)
            for (int gi = 0; gi < NG; gi++)
                if ((rc = finish_part(w, gi)) != BVC_OK) return rc;
    }
    for (int w = flushed; w < nwaves; w++)
        for (int gi = 0; gi < NG; gi++)
            if ((rc = flush_part(w, gi)) != BVC_OK) return rc;
    if (NG > 1) {   // everything the groups still have in flight (phase planes, reconstruction downloads) before the clip counts as done
        for (int gi = 0; gi < NG; gi++) {
            CK(cudaEventRecord(c->ev_post[gi], c->st_post[gi]));
            CK(cudaStreamWaitEvent(c->st, c->ev_post[gi], 0));
            CK(cudaEventRecord(c->ev_me[gi], c->st_grp[gi]));
            CK(cudaStreamWaitEvent(c->st, c->ev_me[gi], 0));
            CK(cudaEventRecord(c->ev_pack[gi], c->st_pack[gi]));
            CK(cudaStreamWaitEvent(c->st, c->ev_pack[gi], 0));
        }
    }
    for (int gi = 0; gi < NG; gi++)
        if (ev_frag_free[(size_t)(nwaves - 1) * NG + gi]) CK(cudaStreamWaitEvent(c->st, ev_frag_free[(size_t)(nwaves - 1) * NG + gi], 0));
    CK(cudaEventRecord(ev_clip1, c->st));
    CK(cudaStreamSynchronize(c->st));
    CK(cudaStreamSynchronize(c->st_d2h));
    *out_len = out_off;
    c->container_len = out_off;

    // ---- instrumentation ----
    for (int i = 0; i < BVC_NUM_KERNEL_CLASSES; i++) { c->last_ms[i] = 0; c->last_launches[i] = 0; }
    for (const auto& sp : c->spans) {
        float ms = 0;
        if (cudaEventElapsedTime(&ms, c->ev_pool[sp.e0], c->ev_pool[sp.e1]) == cudaSuccess) { c->last_ms[sp.cls] += ms; c->last_launches[sp.cls]++; }
    }
    {
        float ms = 0;
        cudaEventElapsedTime(&ms, ev_clip0, ev_clip1);
        c->last_clip_ms = ms;
    }
    return BVC_OK;
}

extern "C" int bvc_encode_clip(bvc_ctx* c, const uint8_t* frames, int nframes, uint8_t* out, size_t out_cap, size_t* out_len,
                               uint8_t* recon) {
    if (!c || !frames) return BVC_ERR_INVALID;
    return encode_clip_impl(c, frames, nframes, out, out_cap, out_len, recon);
}
extern "C" int bvc_encode_clip_resident(bvc_ctx* c, int nframes, uint8_t* out, size_t out_cap, size_t* out_len, uint8_t* recon) {
    if (!c) return BVC_ERR_INVALID;
    if (nframes > c->resident_frames) return fail(c, BVC_ERR_INVALID, "clip not resident: call bvc_clip_upload first");
    return encode_clip_impl(c, nullptr, nframes, out, out_cap, out_len, recon);
}

extern "C" int bvc_encode_clip_device(bvc_ctx* c, const uint8_t* frames, int nframes, size_t cap_hint, size_t* out_len) {
    if (!c) return BVC_ERR_INVALID;
    if (!frames && nframes > c->resident_frames) return fail(c, BVC_ERR_INVALID, "clip not resident: call bvc_clip_upload first");
    const size_t cap = cap_hint ? cap_hint : (size_t)nframes * c->g.W * c->g.H / 2 + ((size_t)1 << 20);
    return encode_clip_impl(c, frames, nframes, nullptr, cap, out_len, nullptr, true);
}
extern "C" int bvc_container_download(bvc_ctx* c, uint8_t* dst, size_t offset, size_t len) {
    if (!c || (!dst && len)) return BVC_ERR_INVALID;
    if (offset > c->container_len || len > c->container_len - offset) return fail(c, BVC_ERR_INVALID, "range outside the container of the last clip call");
    CK(cudaSetDevice(c->device));
    if (len) CK(cudaMemcpyAsync(dst, c->d_container + offset, len, cudaMemcpyDeviceToHost, c->st));
    CK(cudaStreamSynchronize(c->st));
    return BVC_OK;
}
extern "C" int bvc_host_register(void* ptr, size_t bytes) {
    if (!ptr || !bytes) return BVC_ERR_INVALID;
    if (cudaHostRegister(ptr, bytes, cudaHostRegisterPortable) != cudaSuccess) { cudaGetLastError(); return fail(nullptr, BVC_ERR_CUDA, "cudaHostRegister failed"); }
    return BVC_OK;
}
extern "C" int bvc_host_unregister(void* ptr) {
    if (!ptr) return BVC_ERR_INVALID;
    if (cudaHostUnregister(ptr) != cudaSuccess) { cudaGetLastError(); return fail(nullptr, BVC_ERR_CUDA, "cudaHostUnregister failed"); }
    return BVC_OK;
}

// Rate control on the clip path.  rc_flag 1 = RCflag 1 of the reference (encoder/Frame.py:168-188: the QP of every block
// row follows from the frame's bit budget and the bits the rows before it consumed); the feedback loop runs on the
// device.  rc_flag 0 turns it off.  RCflag 2 / 3 couple consecutive frames and GOPs (two passes, scene changes,
// encoder.py:85-98) and stay on the frame-level calls.
extern "C" int bvc_set_rate_control(bvc_ctx* c, int rc_flag, double frame_bit_budget, int n, const int32_t* qps, const int64_t* row_bits) {
    if (!c) return BVC_ERR_INVALID;
    CK(cudaSetDevice(c->device));
    if (rc_flag == 0) {
        c->rc = RcArgs{};
        std::vector<int32_t> q((size_t)c->max_lanes * c->g.bh, c->p.qp);   // back to the base QP on every row
        CK(cudaMemcpyAsync(c->d_qp_rows, q.data(), q.size() * 4, cudaMemcpyHostToDevice, c->st));
        CK(cudaStreamSynchronize(c->st));
        return BVC_OK;
    }
    if (rc_flag != 1) return fail(c, BVC_ERR_UNSUPPORTED, "only RCflag 1 runs on the clip path (2 / 3 couple GOPs: use the frame-level calls)");
    if (n < 1 || n > 16 || !qps || !row_bits) return fail(c, BVC_ERR_INVALID, "rate-control table must hold 1..16 (qp, bits per row) entries");
    int lg = 0; while ((1 << lg) < c->g.bs) lg++;
    RcArgs rc{};
    rc.n = n;
    for (int i = 0; i < n; i++) {
        if (qps[i] < 0 || qps[i] > lg + 7 || (i && qps[i] <= qps[i - 1])) return fail(c, BVC_ERR_INVALID, "rate-control QPs must ascend within 0..log2(block_size)+7");
        rc.qp[i] = qps[i];
        rc.bits[i] = row_bits[i];
    }
    rc.frame_budget = frame_bit_budget;
    c->rc = rc;
    return BVC_OK;
}

// Bytes reserved per frame for its coefficient bit stream on the device (0 = default: 6 bits per pixel, at least 1 MB;
// values above the worst case are clamped to it).  Frees the slots; they are reallocated by the next clip call.
extern "C" int bvc_set_stream_slot_bytes(bvc_ctx* c, size_t bytes) {
    if (!c) return BVC_ERR_INVALID;
    CK(cudaSetDevice(c->device));
    CK(cudaStreamSynchronize(c->st));
    c->slot_bytes = bytes;
    c->coef_cap_words = default_coef_cap_words(c);
    cudaFree(c->d_coef_stream); cudaFree(c->d_pred_stream); cudaFree(c->d_frame_bits); cudaFree(c->d_frame_off);
    c->d_coef_stream = c->d_pred_stream = nullptr; c->d_frame_bits = c->d_frame_off = nullptr;
    c->stream_slots = 0;
    return BVC_OK;
}
