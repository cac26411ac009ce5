/*
 * bvc_oracle.c -- CPU restatement of the encoder hot path of dheri/basic_video_codec.
 * TEST INFRASTRUCTURE ONLY (see bvc_oracle.h).  `file:line` = reference repository paths.
 */
#include "bvc_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define BVO_MAX_BS 32
#define BVO_EOB 8190 /* Frame.EOB_MARKER, Frame.py:23 */

/* ------------------------------------------------------------------------------------------ */
/* half-pel plane: block_predictor.py:145-177.  ceil((a+b)/2) = (a+b+1)>>1, ceil(s/4)=(s+3)>>2.  */
/* The last row and column of the 2x plane are never written by the reference (stay 0).         */
void bvo_halfpel_plane(const uint8_t *ref, int W, int H, uint8_t *out)
{
    const int W2 = 2 * W;
    memset(out, 0, (size_t)4 * W * H);
    for (int y = 0; y < H; y++) {
        const uint8_t *r0 = ref + (size_t)y * W;
        const uint8_t *r1 = (y + 1 < H) ? r0 + W : NULL;
        uint8_t *o0 = out + (size_t)(2 * y) * W2;
        uint8_t *o1 = o0 + W2;
        for (int x = 0; x < W; x++) {
            int a = r0[x];
            o0[2 * x] = (uint8_t)a;
            if (x + 1 < W) o0[2 * x + 1] = (uint8_t)((a + r0[x + 1] + 1) >> 1);
            if (r1) {
                o1[2 * x] = (uint8_t)((a + r1[x] + 1) >> 1);
                if (x + 1 < W) o1[2 * x + 1] = (uint8_t)((a + r0[x + 1] + r1[x] + r1[x + 1] + 3) >> 2);
            }
        }
    }
}

/* ------------------------------------------------------------------------------------------ */
/* SAD == i*i*mae (common.py:43-45): int16 cur minus uint8 ref, |.|, mean -> same ordering.      */
static inline int32_t sad_int(const uint8_t *cur, int cstride, const uint8_t *ref, int rstride, int bs)
{
    int32_t s = 0;
    for (int y = 0; y < bs; y++) {
        const uint8_t *c = cur + (size_t)y * cstride;
        const uint8_t *r = ref + (size_t)y * rstride;
        for (int x = 0; x < bs; x++) s += abs((int)c[x] - (int)r[x]);
    }
    return s;
}
/* half-pel: ref sampled with step 2 in both directions, block_predictor.py:103-111 */
static inline int32_t sad_step2(const uint8_t *cur, int cstride, const uint8_t *ref, int rstride, int bs)
{
    int32_t s = 0;
    for (int y = 0; y < bs; y++) {
        const uint8_t *c = cur + (size_t)y * cstride;
        const uint8_t *r = ref + (size_t)(2 * y) * rstride;
        for (int x = 0; x < bs; x++) s += abs((int)c[x] - (int)r[2 * x]);
    }
    return s;
}

/* is_out_of_range + get_ref_block_at_mv, block_predictor.py:93-143.
 * Returns -1 when the candidate leaves the plane (the reference raises and the caller skips). */
static inline int32_t cand_sad(const uint8_t *cur, int W, int H, int ox, int oy, int bs,
                               const uint8_t *plane, int frac, int mx, int my)
{
    const uint8_t *c = cur + (size_t)oy * W + ox;
    if (!frac) {
        int x = ox + mx, y = oy + my;
        if (x < 0 || y < 0 || x + bs > W || y + bs > H) return -1;
        return sad_int(c, W, plane + (size_t)y * W + x, W, bs);
    } else {
        int W2 = 2 * W, H2 = 2 * H;
        int x = 2 * ox + mx, y = 2 * oy + my;
        if (x < 0 || y < 0 || x + 2 * bs > W2 || y + 2 * bs > H2) return -1;
        return sad_step2(c, W, plane + (size_t)y * W2 + x, W2, bs);
    }
}

int32_t bvo_full_search_block(const uint8_t *cur, int W, int H, int ox, int oy, int bs,
                              const uint8_t *const *planes, int nref, int range, int frac,
                              int32_t mv[3], int64_t *valid_out)
{
    /* block_predictor.py:65-66 : range doubles in half-pel units */
    const int R = frac ? 2 * range : range;
    int32_t best = INT32_MAX; /* min_mae = inf, :72 */
    int bx = 0, by = 0, bk = 0; /* best_mv = [0,0,0], :73 */
    int64_t valid = 0;
    for (int k = 0; k < nref; k++) {                 /* :76 */
        for (int my = -R; my <= R; my++) {            /* :78 */
            for (int mx = -R; mx <= R; mx++) {        /* :79 */
                int32_t s = cand_sad(cur, W, H, ox, oy, bs, planes[k], frac, mx, my);
                if (s < 0) continue;                 /* :81-84 */
                valid++;
                /* :88 -- lower error, or equal error and smaller |mvx|+|mvy| */
                if (s < best || (s == best && abs(mx) + abs(my) < abs(bx) + abs(by))) {
                    best = s; bx = mx; by = my; bk = k;
                }
            }
        }
    }
    mv[0] = bx; mv[1] = by; mv[2] = bk;
    if (valid_out) *valid_out = valid;
    return best;
}

int32_t bvo_fast_me_block(const uint8_t *cur, int W, int H, int ox, int oy, int bs,
                          const uint8_t *const *planes, int nref, int frac,
                          int mvp_x, int mvp_y, int32_t mv[3], int64_t *comparisons)
{
    for (;;) {
        /* key order inside one reference: origin, pmv_origin, top, right, bottom, left (:23-28) */
        const int px[6] = {0, mvp_x, mvp_x, mvp_x + 1, mvp_x, mvp_x - 1};
        const int py[6] = {0, mvp_y, mvp_y - 1, mvp_y, mvp_y + 1, mvp_y};
        int32_t best = INT32_MAX; /* min_mae = inf, :15 */
        int best_p = 0;
        /* Late-binding closures (:20-47): in iteration k every key registered so far (refs 0..k)
         * is evaluated against reference k, keys of ref 0 first, strict '<'.  The winning key is
         * therefore always a ref-0 key; its *value* may come from any reference. */
        for (int k = 0; k < nref; k++)
            for (int kk = 0; kk <= k; kk++)
                for (int p = 0; p < 6; p++) {
                    int32_t s = cand_sad(cur, W, H, ox, oy, bs, planes[k], frac, px[p], py[p]);
                    if (s < 0) continue;             /* exception swallowed, :43-46 */
                    if (comparisons) (*comparisons)++; /* :40 */
                    if (s < best) { best = s; best_p = p; } /* first strict minimum, :41 */
                }
        mv[0] = px[best_p]; mv[1] = py[best_p]; mv[2] = 0; /* reported ref index is always 0 */
        if (best_p <= 1) return best;                 /* "origin" in key, :50-51 */
        if (abs(mv[0]) >= 16 || abs(mv[1]) >= 16) return best; /* :55-56 */
        mvp_x = mv[0]; mvp_y = mv[1];                 /* recurse with mvp = best_mv, :58 */
    }
}

int64_t bvo_me_frame(const bvo_config *cfg, const uint8_t *cur, const uint8_t *const *planes,
                     int nref_avail, int32_t *mv, int32_t *sad)
{
    const int W = cfg->width, H = cfg->height, bs = cfg->block;
    const int bw = W / bs, bh = H / bs;
    int64_t total = 0;
    if (!cfg->fastme) {
        const int R = cfg->frac ? 2 * cfg->range : cfg->range;
        for (int b = 0; b < bw * bh; b++) {
            int ox = (b % bw) * bs, oy = (b / bw) * bs;
            sad[b] = bvo_full_search_block(cur, W, H, ox, oy, bs, planes, nref_avail, cfg->range,
                                           cfg->frac, mv + 3 * b, NULL);
            /* nominal count, block_predictor.py:91 */
            total += (int64_t)nref_avail * (2 * R + 1) * (2 * R + 1);
        }
    } else {
        int mvp_x = 0, mvp_y = 0; /* mv_field = {(0,0):[0,0]}, PFrame.py:34,44,105 */
        for (int b = 0; b < bw * bh; b++) {
            int ox = (b % bw) * bs, oy = (b / bw) * bs;
            int64_t cmp = 0;
            sad[b] = bvo_fast_me_block(cur, W, H, ox, oy, bs, planes, nref_avail, cfg->frac,
                                       mvp_x, mvp_y, mv + 3 * b, &cmp);
            total += cmp;
            mvp_x = mv[3 * b]; mvp_y = mv[3 * b + 1]; /* previous block in raster order */
        }
    }
    return total;
}

/* ------------------------------------------------------------------------------------------ */
/* Defined fp64 DCT.  Tables are derived here independently of the product's generated header   */
/* (long-double libm, rounded once to double); tests check both agree bit-for-bit.               */
typedef struct dct_tab {
    int ready;
    double ct[BVO_MAX_BS * BVO_MAX_BS]; /* Ct[u][x] */
    double w[BVO_MAX_BS * BVO_MAX_BS];  /* W[u][v]  */
} dct_tab;
static dct_tab g_tab[6]; /* log2(bs) = 1..5 */

static int ilog2(int v) { int l = 0; while ((1 << l) < v) l++; return l; }

static const dct_tab *get_tab(int bs)
{
    dct_tab *t = &g_tab[ilog2(bs)];
    if (t->ready) return t;
#pragma omp critical(bvo_tab)
    {
        if (!t->ready) {
            const long double PI_L = 3.14159265358979323846264338327950288419716939937510L;
            for (int u = 0; u < bs; u++)
                for (int x = 0; x < bs; x++) {
                    double v;
                    if (u == 0) v = 1.0;
                    else if (2 * u == bs) v = ((x & 3) == 0 || (x & 3) == 3) ? 1.0 : -1.0;
                    else {
                        int k = ((2 * x + 1) * u) % (4 * bs);
                        v = (double)cosl(PI_L * (long double)k / (long double)(2 * bs));
                    }
                    t->ct[u * bs + x] = v;
                }
            const double w_ss = (double)(1.0L / bs);
            const double w_sn = (double)(sqrtl(2.0L) / bs);
            const double w_nn = (double)(2.0L / bs);
            for (int u = 0; u < bs; u++)
                for (int v = 0; v < bs; v++) {
                    int su = (u == 0 || 2 * u == bs), sv = (v == 0 || 2 * v == bs);
                    t->w[u * bs + v] = (su && sv) ? w_ss : (su || sv) ? w_sn : w_nn;
                }
            t->ready = 1;
        }
    }
    return t;
}
const double *bvo_dct_ct(int bs) { return get_tab(bs)->ct; }
const double *bvo_dct_w(int bs) { return get_tab(bs)->w; }

/* dct.py:21-32 */
int bvo_q_shift(int bs, int qp, int u, int v)
{
    if (u + v < bs - 1) return qp;
    if (u + v == bs - 1) return qp + 1;
    return qp + 2;
}

/* Forward: apply_dct_2d dct.py:9-12 transforms columns first, then rows.  Both passes use the
 * even/odd symmetry Ct[u][N-1-x] = (-1)^u Ct[u][x]:  with s[x] = a[x] + a[N-1-x], d[x] = a[x] - a[N-1-x],
 *     out[u] = sum_{x < N/2} Ct[u][x] * (u even ? s[x] : d[x])        (fma chain from 0.0, x ascending)
 * pass 1 (columns):  T[u][x]  from the integer residual (s, d exact);
 * pass 2 (rows):     Yu[u][v] from T (s, d rounded once);   coef = Yu * W   (one rounding).         */
static void fold_fwd(const double *a, int n, const double *ct, double *out)
{
    const int h = n / 2;
    double s[BVO_MAX_BS / 2], d[BVO_MAX_BS / 2];
    for (int x = 0; x < h; x++) { s[x] = a[x] + a[n - 1 - x]; d[x] = a[x] - a[n - 1 - x]; }
    for (int u = 0; u < n; u++) {
        const double *in = (u & 1) ? d : s;
        double acc = 0.0;
        for (int x = 0; x < h; x++) acc = fma(ct[u * n + x], in[x], acc);
        out[u] = acc;
    }
}
/* Inverse 1-D: E[y] = sum_{u even} Ct[u][y] V[u], O[y] = sum_{u odd} Ct[u][y] V[u] (fma chains, u
 * ascending), out[y] = E + O, out[N-1-y] = E - O for y < N/2.                                      */
static void fold_inv(const double *v, int n, const double *ct, double *out)
{
    const int h = n / 2;
    for (int y = 0; y < h; y++) {
        double e = 0.0, o = 0.0;
        for (int u = 0; u < n; u += 2) e = fma(ct[u * n + y], v[u], e);
        for (int u = 1; u < n; u += 2) o = fma(ct[u * n + y], v[u], o);
        out[y] = e + o;
        out[n - 1 - y] = e - o;
    }
}

void bvo_fdct(const int16_t *res, int bs, double *coef)
{
    const dct_tab *t = get_tab(bs);
    double T[BVO_MAX_BS * BVO_MAX_BS], a[BVO_MAX_BS], o[BVO_MAX_BS];
    for (int x = 0; x < bs; x++) {
        for (int y = 0; y < bs; y++) a[y] = (double)res[y * bs + x];
        fold_fwd(a, bs, t->ct, o);
        for (int u = 0; u < bs; u++) T[u * bs + x] = o[u];
    }
    for (int u = 0; u < bs; u++) {
        fold_fwd(T + u * bs, bs, t->ct, o);
        for (int v = 0; v < bs; v++) coef[u * bs + v] = o[v] * t->w[u * bs + v];
    }
}

/* Inverse: V = coef * W (one rounding); pass 1 over u for every column v; pass 2 over v for every row. */
void bvo_idct(const double *coef, int bs, double *out)
{
    const dct_tab *t = get_tab(bs);
    double Rm[BVO_MAX_BS * BVO_MAX_BS], a[BVO_MAX_BS], o[BVO_MAX_BS];
    for (int v = 0; v < bs; v++) {
        for (int u = 0; u < bs; u++) a[u] = coef[u * bs + v] * t->w[u * bs + v];
        fold_inv(a, bs, t->ct, o);
        for (int y = 0; y < bs; y++) Rm[y * bs + v] = o[y];
    }
    for (int y = 0; y < bs; y++) fold_inv(Rm + y * bs, bs, t->ct, out + y * bs);
}

/* Frame.py:190-202 */
void bvo_transform_block(const int16_t *res, const int16_t *pred, int bs, int qp,
                         int16_t *level, uint8_t *recon, double *idct, double *coef_out)
{
    double coef[BVO_MAX_BS * BVO_MAX_BS], resc[BVO_MAX_BS * BVO_MAX_BS], id[BVO_MAX_BS * BVO_MAX_BS];
    bvo_fdct(res, bs, coef);
    for (int u = 0; u < bs; u++)
        for (int v = 0; v < bs; v++) {
            int s = bvo_q_shift(bs, qp, u, v);
            /* quantize_block dct.py:35-37: np.round = half-to-even == rint in the default mode;
             * dividing by 2^s is exact. */
            double q = rint(ldexp(coef[u * bs + v], -s));
            level[u * bs + v] = (int16_t)q;
            /* rescale_block dct.py:40-42 (exact) */
            resc[u * bs + v] = ldexp(q, s);
        }
    bvo_idct(resc, bs, id);
    for (int i = 0; i < bs * bs; i++) {
        /* Frame.py:200-201: round(idct + pred) -> int16 -> clip 0..255 */
        double r = rint(id[i] + (double)pred[i]);
        int v = (int)(int16_t)r;
        recon[i] = (uint8_t)(v < 0 ? 0 : v > 255 ? 255 : v);
    }
    if (idct) memcpy(idct, id, sizeof(double) * bs * bs);
    if (coef_out) memcpy(coef_out, coef, sizeof(double) * bs * bs);
}

/* ------------------------------------------------------------------------------------------ */
/* bit container: bitarray semantics (MSB first; tobytes() zero-pads the last byte)              */
void bvo_bits_init(bvo_bits *b) { b->data = NULL; b->nbits = 0; b->cap_bytes = 0; }
void bvo_bits_free(bvo_bits *b) { free(b->data); bvo_bits_init(b); }
static void bits_reserve(bvo_bits *b, size_t extra_bits)
{
    size_t need = (b->nbits + extra_bits + 7) / 8 + 8;
    if (need > b->cap_bytes) {
        size_t nc = b->cap_bytes ? b->cap_bytes * 2 : 4096;
        while (nc < need) nc *= 2;
        b->data = (uint8_t *)realloc(b->data, nc);
        memset(b->data + b->cap_bytes, 0, nc - b->cap_bytes);
        b->cap_bytes = nc;
    }
}
static void put_bits(bvo_bits *b, uint64_t val, int n) /* n <= 57, MSB of the n-bit field first */
{
    bits_reserve(b, (size_t)n);
    for (int i = n - 1; i >= 0; i--) {
        if ((val >> i) & 1u) b->data[b->nbits >> 3] |= (uint8_t)(0x80u >> (b->nbits & 7));
        b->nbits++;
    }
}
/* exp_golomb_encode, entropy_encoder.py:8-29 */
static inline uint32_t eg_map(int32_t v) { return v <= 0 ? (uint32_t)(-2 * (int64_t)v) : (uint32_t)(2 * (int64_t)v - 1); }
int bvo_eg_len(int32_t v)
{
    uint32_t e = eg_map(v) + 1;
    int nb = 32 - __builtin_clz(e);
    return 2 * nb - 1;
}
void bvo_put_eg(bvo_bits *b, int32_t v)
{
    uint32_t e = eg_map(v) + 1;
    int nb = 32 - __builtin_clz(e);
    put_bits(b, e, 2 * nb - 1); /* nb-1 zeros then e in nb bits */
}

/* zigzag_order, entropy_encoder.py:115-135 */
void bvo_zigzag(const int16_t *blk, int stride, int bs, int16_t *out)
{
    int n = 0;
    for (int s = 0; s < 2 * bs - 1; s++) {
        if ((s & 1) == 0) {
            for (int i = 0; i <= s; i++)
                if (i < bs && s - i < bs) out[n++] = blk[i * stride + (s - i)];
        } else {
            for (int i = 0; i <= s; i++)
                if (i < bs && s - i < bs) out[n++] = blk[(s - i) * stride + i];
        }
    }
}

/* rle_encode, entropy_encoder.py:65-88 */
int bvo_rle(const int16_t *zz, int n, int32_t *out)
{
    int m = 0, i = 0;
    while (i < n) {
        if (zz[i] == 0) {
            int c = 0;
            while (i < n && zz[i] == 0) { c++; i++; }
            out[m++] = (i < n) ? c : 0;
        } else {
            int st = i, c = 0;
            while (i < n && zz[i] != 0) { c++; i++; }
            out[m++] = -c;
            for (int j = st; j < i; j++) out[m++] = zz[j];
        }
    }
    return m;
}

/* entropy_encode_dct_coffs_row, Frame.py:61-75 */
static void code_coef_row(const bvo_config *cfg, const int16_t *levels, int row, bvo_bits *bits)
{
    const int W = cfg->width, bs = cfg->block;
    int16_t zz[BVO_MAX_BS * BVO_MAX_BS];
    int32_t rl[2 * BVO_MAX_BS * BVO_MAX_BS + 1];
    for (int x = 0; x < W; x += bs) {
        bvo_zigzag(levels + (size_t)row * bs * W + x, W, bs, zz);
        int m = bvo_rle(zz, bs * bs, rl);
        for (int j = 0; j < m; j++) bvo_put_eg(bits, rl[j]);
        bvo_put_eg(bits, BVO_EOB);
    }
}

/* ------------------------------------------------------------------------------------------ */
static void gather_block_i16(const uint8_t *p, int stride, int bs, int16_t *out)
{
    for (int y = 0; y < bs; y++)
        for (int x = 0; x < bs; x++) out[y * bs + x] = p[(size_t)y * stride + x];
}

void bvo_encode_pframe(const bvo_config *cfg, const uint8_t *cur, const uint8_t *const *refs,
                       const uint8_t *const *hp_refs, int nref_avail, const int32_t *qp_rows,
                       bvo_frame_out *out)
{
    const int W = cfg->width, H = cfg->height, bs = cfg->block;
    const int bw = W / bs, bh = H / bs, nblk = bw * bh;
    const uint8_t *const *planes = cfg->frac ? hp_refs : refs;

    out->mae_comparisons = bvo_me_frame(cfg, cur, planes, nref_avail, out->mv, out->sad);

    int16_t c16[BVO_MAX_BS * BVO_MAX_BS], p16[BVO_MAX_BS * BVO_MAX_BS], res[BVO_MAX_BS * BVO_MAX_BS];
    int16_t lev[BVO_MAX_BS * BVO_MAX_BS];
    uint8_t rec[BVO_MAX_BS * BVO_MAX_BS];
    double id[BVO_MAX_BS * BVO_MAX_BS];
    double mae_sum = 0.0;
    int32_t prev_mv[3] = {0, 0, 0}; /* PFrame.py:140 ; chained across rows (:144) */
    size_t pred_len = 0, coef_len = 0;

    int64_t prev_row_bits = 0;
    for (int by = 0; by < bh; by++) {
        const int qp = out->qp_cb ? out->qp_cb(out->qp_user, by, prev_row_bits) : (qp_rows ? qp_rows[by] : cfg->qp);
        if (out->qp_used) out->qp_used[by] = qp;
        for (int bx = 0; bx < bw; bx++) {
            const int b = by * bw + bx, ox = bx * bs, oy = by * bs;
            const int32_t *mv = out->mv + 3 * b;
            gather_block_i16(cur + (size_t)oy * W + ox, W, bs, c16);
            /* find_mv_predicted_block PFrame.py:230-244: refs[mv[2]] if len(refs) > 1 else refs[0] */
            const int k = (nref_avail > 1) ? mv[2] : 0;
            if (!cfg->frac) {
                gather_block_i16(refs[k] + (size_t)(oy + mv[1]) * W + (ox + mv[0]), W, bs, p16);
            } else {
                const uint8_t *hp = hp_refs[k] + (size_t)(2 * oy + mv[1]) * (2 * W) + (2 * ox + mv[0]);
                for (int y = 0; y < bs; y++)
                    for (int x = 0; x < bs; x++) p16[y * bs + x] = hp[(size_t)(2 * y) * (2 * W) + 2 * x];
            }
            for (int i = 0; i < bs * bs; i++) res[i] = (int16_t)(c16[i] - p16[i]); /* PFrame.py:248 */
            bvo_transform_block(res, p16, bs, qp, lev, rec, id, NULL);
            for (int y = 0; y < bs; y++)
                for (int x = 0; x < bs; x++) {
                    size_t o = (size_t)(oy + y) * W + ox + x;
                    out->recon[o] = rec[y * bs + x];
                    out->levels[o] = lev[y * bs + x];
                    /* PFrame.py:39,63: float64 idct residual stored into an int8 plane (C cast) */
                    if (out->resid_mc) out->resid_mc[o] = (int8_t)(int32_t)id[y * bs + x];
                    /* PFrame.py:40,64,103,116: int16(cur) - int16(refs[0]) stored into int8 */
                    if (out->resid_nomc) out->resid_nomc[o] = (int8_t)(cur[o] - refs[0][o]);
                }
            mae_sum += (double)out->sad[b] / (double)(bs * bs); /* PFrame.py:67 */
        }
        /* entropy_encode_prediction_data_row PFrame.py:136-163 */
        bvo_put_eg(&out->pred_bits, qp - cfg->qp); /* always relative to the base qp (:42,72) */
        for (int bx = 0; bx < bw; bx++) {
            const int32_t *mv = out->mv + 3 * (by * bw + bx);
            bvo_put_eg(&out->pred_bits, mv[0] - prev_mv[0]);
            bvo_put_eg(&out->pred_bits, mv[1] - prev_mv[1]);
            if (cfg->nref > 1) bvo_put_eg(&out->pred_bits, mv[2] - prev_mv[2]);
            prev_mv[0] = mv[0]; prev_mv[1] = mv[1]; prev_mv[2] = mv[2];
        }
        code_coef_row(cfg, out->levels, by, &out->coef_bits);
        prev_row_bits = (int64_t)(out->coef_bits.nbits - coef_len) + (int64_t)(out->pred_bits.nbits - pred_len);
        if (out->bits_per_row) out->bits_per_row[by] = prev_row_bits;
        pred_len = out->pred_bits.nbits; coef_len = out->coef_bits.nbits;
    }
    out->avg_mae = mae_sum / (double)nblk; /* PFrame.py:88 */
}

void bvo_encode_iframe(const bvo_config *cfg, const uint8_t *cur, const int32_t *qp_rows,
                       bvo_frame_out *out)
{
    const int W = cfg->width, H = cfg->height, bs = cfg->block;
    const int bw = W / bs, bh = H / bs, nblk = bw * bh;
    int16_t c16[BVO_MAX_BS * BVO_MAX_BS], p16[BVO_MAX_BS * BVO_MAX_BS], res[BVO_MAX_BS * BVO_MAX_BS];
    int16_t lev[BVO_MAX_BS * BVO_MAX_BS];
    uint8_t rec[BVO_MAX_BS * BVO_MAX_BS];
    double mae_sum = 0.0;
    size_t pred_len = 0, coef_len = 0;
    uint8_t *recon = out->recon;
    memset(recon, 0, (size_t)W * H);
    out->mae_comparisons = 0;

    int64_t prev_row_bits = 0;
    for (int by = 0; by < bh; by++) {
        const int qp = out->qp_cb ? out->qp_cb(out->qp_user, by, prev_row_bits) : (qp_rows ? qp_rows[by] : cfg->qp);
        if (out->qp_used) out->qp_used[by] = qp;
        for (int bx = 0; bx < bw; bx++) {
            const int b = by * bw + bx, ox = bx * bs, oy = by * bs;
            const uint8_t *c = cur + (size_t)oy * W + ox;
            /* intra_predict_block IFrame.py:184-195.  Mode 0 ("horizontal"): np.tile(left,(i,1))
             * -> pred[r][c] = recon[oy+c][ox-1]; mode 1: tile(top).T -> pred[r][c] = recon[oy-1][ox+r]
             * (:198-213).  In-frame predictors are uint8 so cur - pred wraps mod 256 (:189-190);
             * border predictors are int64 128 -> true absolute difference. */
            int32_t sh = 0, sv = 0;
            for (int r = 0; r < bs; r++)
                for (int cc = 0; cc < bs; cc++) {
                    int cv = c[(size_t)r * W + cc];
                    if (ox > 0) sh += (uint8_t)(cv - recon[(size_t)(oy + cc) * W + ox - 1]);
                    else sh += abs(cv - 128);
                    if (oy > 0) sv += (uint8_t)(cv - recon[(size_t)(oy - 1) * W + ox + r]);
                    else sv += abs(cv - 128);
                }
            const int mode = (sh < sv) ? 0 : 1; /* :192-195, tie -> vertical */
            out->modes[b] = mode;
            out->sad[b] = mode == 0 ? sh : sv;
            for (int r = 0; r < bs; r++)
                for (int cc = 0; cc < bs; cc++) {
                    int pv;
                    if (mode == 0) pv = ox > 0 ? recon[(size_t)(oy + cc) * W + ox - 1] : 128;
                    else pv = oy > 0 ? recon[(size_t)(oy - 1) * W + ox + r] : 128;
                    p16[r * bs + cc] = (int16_t)pv;
                }
            gather_block_i16(c, W, bs, c16);
            for (int i = 0; i < bs * bs; i++) res[i] = (int16_t)(c16[i] - p16[i]); /* IFrame.py:222 */
            bvo_transform_block(res, p16, bs, qp, lev, rec, NULL, NULL);
            for (int y = 0; y < bs; y++)
                for (int x = 0; x < bs; x++) {
                    size_t o = (size_t)(oy + y) * W + ox + x;
                    recon[o] = rec[y * bs + x];
                    out->levels[o] = lev[y * bs + x];
                    /* IFrame.py:30,57-58: int16 residual stored into a uint8 plane */
                    if (out->resid_mc) out->resid_mc[o] = (int8_t)(uint8_t)res[y * bs + x];
                }
            mae_sum += (double)out->sad[b] / (double)(bs * bs);
            out->mae_comparisons += 2; /* params.py:62 */
        }
        /* entropy_encode_prediction_data_row IFrame.py:116-130 */
        bvo_put_eg(&out->pred_bits, qp - cfg->qp);
        for (int bx = 0; bx < bw; bx++) bvo_put_eg(&out->pred_bits, out->modes[by * bw + bx]);
        code_coef_row(cfg, out->levels, by, &out->coef_bits);
        prev_row_bits = (int64_t)(out->coef_bits.nbits - coef_len) + (int64_t)(out->pred_bits.nbits - pred_len);
        if (out->bits_per_row) out->bits_per_row[by] = prev_row_bits;
        pred_len = out->pred_bits.nbits; coef_len = out->coef_bits.nbits;
    }
    out->avg_mae = mae_sum / (double)nblk;
}

/* ------------------------------------------------------------------------------------------ */
typedef struct bytebuf { uint8_t *p; size_t n, cap; } bytebuf;
static void bb_put(bytebuf *b, const void *src, size_t n)
{
    if (b->n + n > b->cap) {
        size_t nc = b->cap ? b->cap * 2 : 65536;
        while (nc < b->n + n) nc *= 2;
        b->p = (uint8_t *)realloc(b->p, nc);
        b->cap = nc;
    }
    memcpy(b->p + b->n, src, n);
    b->n += n;
}

/* one GOP (I then P frames), appended to `bb`.  encoder.py:75-121,154-155,174-186 */
static int encode_gop(const bvo_config *cfg, const uint8_t *frames, int n, bytebuf *bb, uint8_t *recon_out)
{
    const int W = cfg->width, H = cfg->height, bs = cfg->block;
    const size_t fsz = (size_t)W * H;
    const int nblk = (W / bs) * (H / bs);
    const int nref = cfg->nref;
    uint8_t **refs = (uint8_t **)calloc((size_t)nref, sizeof(uint8_t *));
    uint8_t **hps = (uint8_t **)calloc((size_t)nref, sizeof(uint8_t *));
    for (int k = 0; k < nref; k++) {
        refs[k] = (uint8_t *)malloc(fsz);
        hps[k] = cfg->frac ? (uint8_t *)malloc(4 * fsz) : NULL;
    }
    int navail = 0;
    bvo_frame_out fo;
    memset(&fo, 0, sizeof(fo));
    fo.recon = (uint8_t *)malloc(fsz);
    fo.levels = (int16_t *)malloc(fsz * sizeof(int16_t));
    fo.mv = (int32_t *)malloc(sizeof(int32_t) * 3 * nblk);
    fo.sad = (int32_t *)malloc(sizeof(int32_t) * nblk);
    fo.modes = (int32_t *)malloc(sizeof(int32_t) * nblk);
    int rc = 0;
    for (int f = 0; f < n; f++) {
        const uint8_t *cur = frames + (size_t)f * fsz;
        bvo_bits_init(&fo.pred_bits);
        bvo_bits_init(&fo.coef_bits);
        uint8_t mode;
        if (f == 0) { /* I frame: reference window cleared, encoder.py:175-178 */
            navail = 0;
            bvo_encode_iframe(cfg, cur, NULL, &fo);
            mode = 1;
        } else {
            bvo_encode_pframe(cfg, cur, (const uint8_t *const *)refs, (const uint8_t *const *)hps, navail, NULL, &fo);
            mode = 0;
        }
        /* container, encoder.py:104-121 */
        size_t pb = (fo.pred_bits.nbits + 7) / 8, cb = (fo.coef_bits.nbits + 7) / 8;
        if (pb > 0xFFFF || cb > 0xFFFFFF) rc = -2; /* OverflowError in the reference */
        uint8_t hdr[3] = {mode, (uint8_t)(pb >> 8), (uint8_t)pb};
        bb_put(bb, hdr, 3);
        bb_put(bb, fo.pred_bits.data, pb);
        uint8_t hdr2[3] = {(uint8_t)(cb >> 16), (uint8_t)(cb >> 8), (uint8_t)cb};
        bb_put(bb, hdr2, 3);
        bb_put(bb, fo.coef_bits.data, cb);
        bvo_bits_free(&fo.pred_bits);
        bvo_bits_free(&fo.coef_bits);
        if (recon_out) memcpy(recon_out + (size_t)f * fsz, fo.recon, fsz);
        /* deque(maxlen=nRef).append : index 0 = oldest, encoder.py:33,154-155 */
        if (navail == nref) {
            uint8_t *t = refs[0], *th = hps[0];
            for (int k = 0; k + 1 < nref; k++) { refs[k] = refs[k + 1]; hps[k] = hps[k + 1]; }
            refs[nref - 1] = t; hps[nref - 1] = th;
            navail--;
        }
        memcpy(refs[navail], fo.recon, fsz);
        if (cfg->frac) bvo_halfpel_plane(fo.recon, W, H, hps[navail]);
        navail++;
    }
    for (int k = 0; k < nref; k++) { free(refs[k]); free(hps[k]); }
    free(refs); free(hps);
    free(fo.recon); free(fo.levels); free(fo.mv); free(fo.sad); free(fo.modes);
    return rc;
}

int bvo_encode_clip(const bvo_config *cfg, const uint8_t *frames, int nframes, int first_index,
                    uint8_t **out, size_t *out_len, uint8_t *recon_out, int nthreads)
{
    const size_t fsz = (size_t)cfg->width * cfg->height;
    const int ip = cfg->i_period;
    /* GOP boundaries: frame idx (1-based) with (idx-1) % I_Period == 0 is an I frame.  A clip
     * that does not start on a boundary is not self-contained (its references are missing). */
    if ((first_index - 1) % ip != 0) return -1;
    int ngop = (nframes + ip - 1) / ip;
    bytebuf *bbs = (bytebuf *)calloc((size_t)ngop, sizeof(bytebuf));
    int rc = 0;
#ifdef _OPENMP
    if (nthreads < 1) nthreads = 1;
#pragma omp parallel for schedule(dynamic, 1) num_threads(nthreads) reduction(min : rc)
#endif
    for (int g = 0; g < ngop; g++) {
        int f0 = g * ip, n = nframes - f0 < ip ? nframes - f0 : ip;
        int r = encode_gop(cfg, frames + (size_t)f0 * fsz, n, &bbs[g], recon_out ? recon_out + (size_t)f0 * fsz : NULL);
        if (r < rc) rc = r;
    }
    size_t total = 0;
    for (int g = 0; g < ngop; g++) total += bbs[g].n;
    uint8_t *o = (uint8_t *)malloc(total ? total : 1);
    size_t off = 0;
    for (int g = 0; g < ngop; g++) { memcpy(o + off, bbs[g].p, bbs[g].n); off += bbs[g].n; free(bbs[g].p); }
    free(bbs);
    *out = o; *out_len = total;
    (void)nthreads;
    return rc;
}

/* ------------------------------------------------------------------------------------------ */
/* Decoder: decode_video (decoder.py:26-87).  Restated so that the GPU decoder has a CPU  */
/* checker; pinned by running the reference's own decode_video on the golden streams                       */
/* (oracle/gen_golden_decode.py, tests/golden/decode_ref.json).                                            */

typedef struct { const uint8_t *d; size_t nbits, pos; } bitrd;
static inline int rd_bit(const bitrd *r, size_t i) { return (r->d[i >> 3] >> (7 - (i & 7))) & 1; }
/* exp_golomb_decode, entropy_encoder.py:32-62.  Returns 1 = symbol, 0 = end of stream (fewer than 8 zero
 * bits of byte padding left, :40-44), -1 = "Not enough bits" (ValueError, :45-46). */
static int rd_eg(bitrd *r, int32_t *out)
{
    size_t m = 0, left = r->nbits - r->pos;
    while (m < left && !rd_bit(r, r->pos + m)) m++;
    if (m >= left) return left < 8 ? 0 : -1;
    if (r->pos + 2 * m + 1 > r->nbits || m > 30) return -1; /* the reference would index past the end (IndexError) */
    uint32_t value = 1;
    for (size_t i = 1; i <= m; i++) value = (value << 1) | (uint32_t)rd_bit(r, r->pos + m + i);
    value -= 1;
    *out = (value % 2 == 0) ? -(int32_t)(value / 2) : (int32_t)((value + 1) / 2); /* :57 */
    r->pos += 2 * m + 1;
    return 1;
}

/* Frame.entropy_decode_dct_coffs, Frame.py:81-110: symbols -> blocks at every EOB marker -> rle_decode
 * (entropy_encoder.py:91-112) -> pad to bs*bs -> inverse_zigzag_order (:138-160) -> merge_blocks (raster). */
static int decode_coefs(const bvo_config *cfg, const uint8_t *data, size_t nbytes, int16_t *levels)
{
    const int W = cfg->width, H = cfg->height, bs = cfg->block, bw = W / bs, nblk = bw * (H / bs), n = bs * bs;
    bitrd r = {data, nbytes * 8, 0};
    int32_t *sym = (int32_t *)malloc(sizeof(int32_t) * (2 * (size_t)n + 8));
    int16_t zz[BVO_MAX_BS * BVO_MAX_BS];
    int nsym = 0, b = 0, rc = 0;
    memset(levels, 0, sizeof(int16_t) * (size_t)W * H);
    for (;;) {
        int32_t v;
        int k = rd_eg(&r, &v);
        if (k == 0) break;
        if (k < 0) { rc = -1; break; }
        if (v != BVO_EOB) {
            if (nsym >= 2 * n + 8) { rc = -1; break; }
            sym[nsym++] = v;
            continue;
        }
        if (b >= nblk) { rc = -1; break; }
        /* rle_decode */
        int len = 0;
        memset(zz, 0, sizeof zz);
        for (int i = 0; i < nsym; i++) {
            int32_t c = sym[i];
            if (c == 0) break;
            if (c > 0) { len += c; continue; }            /* run of zeros */
            for (int j = 0; j < -c && i + 1 + j < nsym; j++, len++)
                if (len < n) zz[len] = (int16_t)sym[i + 1 + j];
            i += -c;
        }
        nsym = 0;
        /* inverse zig-zag into block b (raster order) */
        const int ox = (b % bw) * bs, oy = (b / bw) * bs;
        int idx = 0;
        for (int sd = 0; sd < 2 * bs - 1; sd++)
            for (int i = 0; i <= sd; i++) {
                if (i >= bs || sd - i >= bs) continue;
                const int rr = (sd % 2 == 0) ? i : sd - i, cc = (sd % 2 == 0) ? sd - i : i;
                levels[(size_t)(oy + rr) * W + ox + cc] = zz[idx++];
            }
        b++;
    }
    free(sym);
    if (rc == 0 && b != nblk) rc = -1;
    return rc;
}

/* PFrame / IFrame.entropy_decode_prediction_data (PFrame.py:166-228, IFrame.py:132-166): per block row the QP
 * difference to the base QP, then per block the MV difference to the previous block in raster order
 * (chained from (0,0,0)) or the intra mode.  pred: nblk*3 (mvx,mvy,ref | mode,0,0). */
static int decode_pred(const bvo_config *cfg, const uint8_t *data, size_t nbytes, int intra, int32_t *pred, int32_t *qp_rows)
{
    const int bs = cfg->block, bw = cfg->width / bs, bh = cfg->height / bs, nblk = bw * bh;
    bitrd r = {data, nbytes * 8, 0};
    int32_t prev[3] = {0, 0, 0}, v;
    for (int b = 0; b < nblk; b++) {
        if (b % bw == 0) {
            if (rd_eg(&r, &v) != 1) return -1;
            qp_rows[b / bw] = cfg->qp + v;
        }
        if (intra) {
            if (rd_eg(&r, &v) != 1) return -1;
            pred[3 * b] = v; pred[3 * b + 1] = 0; pred[3 * b + 2] = 0;
            if (v != 0 && v != 1) return -1;             /* find_intra_predict_block raises ValueError, IFrame.py:175-182 */
        } else {
            for (int k = 0; k < (cfg->nref > 1 ? 3 : 2); k++) {
                if (rd_eg(&r, &v) != 1) return -1;
                prev[k] += v;
            }
            pred[3 * b] = prev[0]; pred[3 * b + 1] = prev[1]; pred[3 * b + 2] = prev[2];
        }
    }
    return 0;
}

static void dequant_idct(const int16_t *lev, int stride, int bs, int qp, double *id)
{
    double resc[BVO_MAX_BS * BVO_MAX_BS];
    for (int u = 0; u < bs; u++)
        for (int v = 0; v < bs; v++) resc[u * bs + v] = ldexp((double)lev[(size_t)u * stride + v], bvo_q_shift(bs, qp, u, v));
    bvo_idct(resc, bs, id);
}

int bvo_decode_clip(const bvo_config *cfg, const uint8_t *data, size_t len, int max_frames, uint8_t *frames_out,
                    int *nframes_out, int16_t *levels_out, int32_t *pred_out, int32_t *qp_out, uint8_t *kinds_out)
{
    const int W = cfg->width, H = cfg->height, bs = cfg->block, bw = W / bs, bh = H / bs, nblk = bw * bh;
    const size_t P = (size_t)W * H;
    const int nref = cfg->nref;
    /* reference window, decoder.py:34-38: starts with one 128-filled frame; an I frame clears it (:57-58) */
    uint8_t **refs = (uint8_t **)calloc((size_t)nref, sizeof *refs), **hps = (uint8_t **)calloc((size_t)nref, sizeof *hps);
    for (int k = 0; k < nref; k++) { refs[k] = (uint8_t *)malloc(P); hps[k] = cfg->frac ? (uint8_t *)malloc(4 * P) : NULL; }
    int navail = 1;
    memset(refs[0], 128, P);
    if (cfg->frac) bvo_halfpel_plane(refs[0], W, H, hps[0]);
    int16_t *levels = (int16_t *)malloc(sizeof(int16_t) * P);
    int32_t *pred = (int32_t *)malloc(sizeof(int32_t) * 3 * (size_t)nblk), *qps = (int32_t *)malloc(sizeof(int32_t) * (size_t)bh);
    uint8_t *cur = (uint8_t *)malloc(P);
    double id[BVO_MAX_BS * BVO_MAX_BS];
    size_t o = 0;
    int n = 0, rc = 0;
    while (n < max_frames && o < len) {
        if (o + 3 > len) { rc = -1; break; }
        const int mode = data[o];
        const size_t pl = ((size_t)data[o + 1] << 8) | data[o + 2];
        if (o + 3 + pl + 3 > len) { rc = -1; break; }
        const uint8_t *pd = data + o + 3;
        const size_t cl = ((size_t)pd[pl] << 16) | ((size_t)pd[pl + 1] << 8) | pd[pl + 2];
        const uint8_t *cd = pd + pl + 3;
        if (o + 6 + pl + cl > len) { rc = -1; break; }
        o += 6 + pl + cl;
        const int intra = (mode == 1); /* PredictionMode.INTRA_FRAME.value, decoder.py:55 */
        if (intra) navail = 0;
        if (decode_pred(cfg, pd, pl, intra, pred, qps) || decode_coefs(cfg, cd, cl, levels)) { rc = -1; break; }
        memset(cur, 0, P);
        for (int by = 0; by < bh && rc == 0; by++)
            for (int bx = 0; bx < bw; bx++) {
                const int b = by * bw + bx, ox = bx * bs, oy = by * bs;
                dequant_idct(levels + (size_t)oy * W + ox, W, bs, qps[by], id);
                for (int y = 0; y < bs; y++)
                    for (int x = 0; x < bs; x++) {
                        int pv;
                        if (intra) {   /* find_intra_predict_block IFrame.py:175-213 on the frame being rebuilt */
                            if (pred[3 * b] == 0) pv = ox > 0 ? cur[(size_t)(oy + x) * W + ox - 1] : 128;
                            else pv = oy > 0 ? cur[(size_t)(oy - 1) * W + ox + y] : 128;
                        } else {       /* find_mv_predicted_block PFrame.py:230-244 */
                            const int k = navail > 1 ? pred[3 * b + 2] : 0;
                            const int mx = pred[3 * b], my = pred[3 * b + 1];
                            if (k < 0 || k >= navail) { rc = -1; break; }
                            if (!cfg->frac) {
                                if (ox + mx < 0 || oy + my < 0 || ox + mx + bs > W || oy + my + bs > H) { rc = -1; break; }
                                pv = refs[k][(size_t)(oy + my + y) * W + ox + mx + x];
                            } else {
                                if (2 * ox + mx < 0 || 2 * oy + my < 0 || 2 * ox + mx + 2 * bs > 2 * W || 2 * oy + my + 2 * bs > 2 * H) { rc = -1; break; }
                                pv = hps[k][(size_t)(2 * oy + my + 2 * y) * (2 * W) + 2 * ox + mx + 2 * x];
                            }
                        }
                        /* PFrame.py:305-308 / IFrame.py:107-108: round(idct + pred) -> int16 -> clip -> uint8 */
                        const double rr = rint(id[y * bs + x] + (double)pv);
                        const int v = (int)(int16_t)rr;
                        cur[(size_t)(oy + y) * W + ox + x] = (uint8_t)(v < 0 ? 0 : v > 255 ? 255 : v);
                    }
            }
        if (rc) break;
        memcpy(frames_out + (size_t)n * P, cur, P);
        if (levels_out) memcpy(levels_out + (size_t)n * P, levels, sizeof(int16_t) * P);
        if (pred_out) memcpy(pred_out + (size_t)n * 3 * nblk, pred, sizeof(int32_t) * 3 * (size_t)nblk);
        if (qp_out) memcpy(qp_out + (size_t)n * bh, qps, sizeof(int32_t) * (size_t)bh);
        if (kinds_out) kinds_out[n] = (uint8_t)intra;
        /* reference_frames.append(decoded_frame): deque(maxlen=nRefFrames), decoder.py:84-85 */
        if (navail == nref) {
            uint8_t *t = refs[0], *h = hps[0];
            for (int k = 1; k < nref; k++) { refs[k - 1] = refs[k]; hps[k - 1] = hps[k]; }
            refs[nref - 1] = t; hps[nref - 1] = h;
            navail--;
        }
        memcpy(refs[navail], cur, P);
        if (cfg->frac) bvo_halfpel_plane(cur, W, H, hps[navail]);
        navail++;
        n++;
    }
    *nframes_out = n;
    for (int k = 0; k < nref; k++) { free(refs[k]); free(hps[k]); }
    free(refs); free(hps); free(levels); free(pred); free(qps); free(cur);
    return rc;
}

void bvo_free(void *p) { free(p); }
