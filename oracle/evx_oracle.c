/* oracle/evx_oracle.c -- TEST INFRASTRUCTURE ONLY (see evx_oracle.h).
 *
 * CPU restatement of the EVX-1 pixel pipeline and slice serialiser.  Every
 * function cites the reference file:line whose arithmetic it restates.  The
 * arithmetic contract is integer only: int16 storage,
 * int32 intermediates, C truncating division, sign-dependent rounding.
 */
#include "evx_oracle.h"

#include <stdlib.h>
#include <string.h>

#define MB 16
#define SAD_CAP 8192          /* EVX_MOTION_SAD_THRESHOLD, motion.cpp:19 */
#define SEARCH_RADIUS 16      /* motion.cpp:24 */
#define INT32_BIG 0x7FFFFFFF

enum { T_INTRA = 1, T_MOTION = 2, T_COPY = 4 };   /* types.h:68-71 */

typedef struct { int16_t *p[3]; } planes_t;

struct evxo_ctx
{
    int vw, vh;             /* visible size */
    int w, h;               /* 16-aligned plane size (evx1enc.cpp:79-80) */
    int mbw, mbh;
    int R, linear, deblocking;
    planes_t src, coef, ring[8];
    evxo_block_desc *table;
    uint64_t n_fullpel, n_subpel;
};

/* ------------------------------------------------------------------ helpers */

static inline int32_t iabs32(int32_t v) { return v == (int32_t) 0x80000000 ? INT32_BIG : (v < 0 ? -v : v); }   /* math.h:205-211 */
static inline int16_t iabs16(int16_t v) { return v == (int16_t) -32768 ? 32767 : (int16_t) (v < 0 ? -v : v); }    /* math.h:197-203 */

/* math.h:228-236 */
static inline int32_t rdiv(int32_t n, int32_t d)
{
    if (((uint32_t) n ^ (uint32_t) d) & 0x80000000u) return (n - d / 2) / d;
    return (n + d / 2) / d;
}

/* math.h:69-138: LUT floor-log2 with log2(0) == 0 */
static inline int ilog2(uint32_t v) { int r = 0; while (v >>= 1) r++; return r; }

static inline int16_t clip16(int16_t v, int16_t lo, int16_t hi) { return v < lo ? lo : (v > hi ? hi : v); }  /* math.h:213-216 */

static inline int isign16(int16_t v) { return v == 0 ? 0 : ((v & 0x8000) ? -1 : 1); }   /* math.h:149-154 */

static int alloc_planes(planes_t *pl, int w, int h)
{
    pl->p[0] = (int16_t *) calloc((size_t) w * h, 2);               /* image.cpp:89 zero fill */
    pl->p[1] = (int16_t *) calloc((size_t) (w / 2) * (h / 2), 2);
    pl->p[2] = (int16_t *) calloc((size_t) (w / 2) * (h / 2), 2);
    return pl->p[0] && pl->p[1] && pl->p[2];
}

static void free_planes(planes_t *pl) { for (int i = 0; i < 3; ++i) free(pl->p[i]); }

static void zero_planes(planes_t *pl, int w, int h)
{
    memset(pl->p[0], 0, (size_t) w * h * 2);
    memset(pl->p[1], 0, (size_t) (w / 2) * (h / 2) * 2);
    memset(pl->p[2], 0, (size_t) (w / 2) * (h / 2) * 2);
}

/* a 16x16 + 2x 8x8 view, macroblock.h:68-88 (chroma origin is (x>>1, y>>1)) */
typedef struct { int16_t *y, *u, *v; int stride; } view_t;

static view_t view_at(const planes_t *pl, int w, int x, int y)
{
    view_t b;
    b.y = pl->p[0] + (size_t) y * w + x;
    b.u = pl->p[1] + (size_t) (y >> 1) * (w >> 1) + (x >> 1);
    b.v = pl->p[2] + (size_t) (y >> 1) * (w >> 1) + (x >> 1);
    b.stride = w;
    return b;
}

/* ------------------------------------------------------------------ context */

evxo_ctx *evxo_create(int width, int height, int ref_count, int linear_quant, int deblocking)
{
    if (width <= 0 || height <= 0 || ref_count < 1 || ref_count > 8) return NULL;
    evxo_ctx *c = (evxo_ctx *) calloc(1, sizeof(*c));
    if (!c) return NULL;
    c->vw = width; c->vh = height;
    c->w = (width + 15) & ~15; c->h = (height + 15) & ~15;
    c->mbw = c->w / MB; c->mbh = c->h / MB;
    c->R = ref_count; c->linear = linear_quant; c->deblocking = deblocking;
    int ok = alloc_planes(&c->src, c->w, c->h) && alloc_planes(&c->coef, c->w, c->h);
    for (int i = 0; i < c->R; ++i) ok = ok && alloc_planes(&c->ring[i], c->w, c->h);
    c->table = (evxo_block_desc *) calloc((size_t) c->mbw * c->mbh, sizeof(evxo_block_desc));
    if (!ok || !c->table) { evxo_destroy(c); return NULL; }
    return c;
}

void evxo_destroy(evxo_ctx *c)
{
    if (!c) return;
    free_planes(&c->src); free_planes(&c->coef);
    for (int i = 0; i < 8; ++i) free_planes(&c->ring[i]);
    free(c->table);
    free(c);
}

void evxo_reset(evxo_ctx *c)
{
    zero_planes(&c->src, c->w, c->h); zero_planes(&c->coef, c->w, c->h);
    for (int i = 0; i < c->R; ++i) zero_planes(&c->ring[i], c->w, c->h);
    memset(c->table, 0, (size_t) c->mbw * c->mbh * sizeof(evxo_block_desc));
    c->n_fullpel = c->n_subpel = 0;
}

int evxo_aligned_width(const evxo_ctx *c) { return c->w; }
int evxo_aligned_height(const evxo_ctx *c) { return c->h; }
int evxo_block_count(const evxo_ctx *c) { return c->mbw * c->mbh; }
evxo_block_desc *evxo_block_table(evxo_ctx *c) { return c->table; }

int16_t *evxo_plane(evxo_ctx *c, int which, int slot, int comp)
{
    if (comp < 0 || comp > 2) return NULL;
    if (which == 0) return c->src.p[comp];
    if (which == 1) return c->coef.p[comp];
    if (which == 2) return c->ring[((slot % c->R) + c->R) % c->R].p[comp];
    return NULL;
}

void evxo_get_counters(const evxo_ctx *c, uint64_t *fullpel, uint64_t *subpel) { *fullpel = c->n_fullpel; *subpel = c->n_subpel; }
void evxo_reset_counters(evxo_ctx *c) { c->n_fullpel = c->n_subpel = 0; }

/* common.cpp:192-195 */
static inline int ring_slot(const evxo_ctx *c, uint32_t index, int offset) { return (int) ((index + (uint32_t) c->R - (uint32_t) offset) % (uint32_t) c->R); }

/* ------------------------------------------------------------------ colour */

/* convert.cpp:11-14, 30-73, 95-160.  y uses >>8, chroma uses C '/' (toward zero);
 * chroma accumulates four samples in an int16 then (sum+2)>>2.  Rows/cols beyond
 * the visible size are never written (they stay 0). */
void evxo_convert_in(evxo_ctx *c, const uint8_t *rgb)
{
    int cw = c->w >> 1;
    for (int j = 0; j < c->vh; j += 2)
    for (int i = 0; i < c->vw; i += 2)
    {
        int16_t su = 0, sv = 0;
        for (int dy = 0; dy < 2; ++dy)
        for (int dx = 0; dx < 2; ++dx)
        {
            const uint8_t *px = rgb + ((size_t) (j + dy) * c->vw + (i + dx)) * 3;
            int r = px[0], g = px[1], b = px[2];
            c->src.p[0][(size_t) (j + dy) * c->w + i + dx] = (int16_t) (((77 * r + 150 * g + 29 * b + 128) >> 8) + 16);
            su = (int16_t) (su + ((-43 * r - 85 * g + 128 * b + 128) / 256 + 128));
            sv = (int16_t) (sv + ((128 * r - 107 * g - 21 * b + 128) / 256 + 128));
        }
        c->src.p[1][(size_t) (j >> 1) * cw + (i >> 1)] = (int16_t) ((su + 2) >> 2);
        c->src.p[2][(size_t) (j >> 1) * cw + (i >> 1)] = (int16_t) ((sv + 2) >> 2);
    }
}

/* math.h:218-221: saturate() funnels its int32 argument through an int16 parameter */
static inline uint8_t sat8(int32_t v) { return (uint8_t) clip16((int16_t) v, 0, 255); }

/* convert.cpp:16-19, 75-93, 162-223 */
void evxo_convert_out(evxo_ctx *c, uint32_t index, uint8_t *rgb)
{
    const planes_t *pl = &c->ring[ring_slot(c, index, 0)];
    int cw = c->w >> 1;
    for (int j = 0; j < c->vh; ++j)
    for (int i = 0; i < c->vw; ++i)
    {
        int32_t y = pl->p[0][(size_t) j * c->w + i];
        int32_t u = pl->p[1][(size_t) (j >> 1) * cw + (i >> 1)];
        int32_t v = pl->p[2][(size_t) (j >> 1) * cw + (i >> 1)];
        uint8_t *o = rgb + ((size_t) j * c->vw + i) * 3;
        o[0] = sat8((256 * (y - 16) + 358 * (v - 128) + 128) >> 8);
        o[1] = sat8((256 * (y - 16) - 88 * (u - 128) - 182 * (v - 128) + 128) >> 8);
        o[2] = sat8((256 * (y - 16) + 452 * (u - 128) + 128) >> 8);
    }
}

/* ------------------------------------------------------------------ block metrics */

/* analysis.h:42-55 */
static int32_t block_sad(const view_t *a, const view_t *b)
{
    int32_t s = 0;
    for (int j = 0; j < MB; ++j)
    for (int i = 0; i < MB; ++i) s += iabs32(a->y[j * a->stride + i] - b->y[j * b->stride + i]);
    return s;
}

/* analysis.h:57-68: the int16 overload of abs is the one selected here */
static int32_t block_sad_self(const view_t *a)
{
    int32_t s = 0;
    for (int j = 0; j < MB; ++j)
    for (int i = 0; i < MB; ++i) s += iabs16(a->y[j * a->stride + i]);
    return s;
}

/* analysis.h:103-125: max |d| over luma AND both chroma blocks */
static int32_t block_mad(const view_t *a, const view_t *b)
{
    int32_t m = 0;
    for (int j = 0; j < MB; ++j)
    for (int i = 0; i < MB; ++i) { int32_t t = iabs32(a->y[j * a->stride + i] - b->y[j * b->stride + i]); if (t > m) m = t; }
    int as = a->stride >> 1, bs = b->stride >> 1;
    for (int j = 0; j < 8; ++j)
    for (int i = 0; i < 8; ++i)
    {
        int32_t tu = iabs32(a->u[j * as + i] - b->u[j * bs + i]);
        int32_t tv = iabs32(a->v[j * as + i] - b->v[j * bs + i]);
        if (tu > m) m = tu;
        if (tv > m) m = tv;
    }
    return m;
}

/* macroblock.h:203-241.  evx_round_out(n,a)/k: add a away from zero, truncate. */
static inline int16_t lerp_half(int32_t a, int32_t b) { int32_t t = a + b; return (int16_t) ((t < 0 ? t - 1 : t + 1) / 2); }
static inline int16_t lerp_quarter(int32_t a, int32_t b) { int32_t t = 3 * a + b; return (int16_t) ((t < 0 ? t - 2 : t + 2) / 4); }

/* scratch macroblock (the reference's motion_cache / transform_cache, common.h:108-109) */
typedef struct { int16_t y[256], u[64], v[64]; } mbuf_t;

static view_t view_of(mbuf_t *m) { view_t b; b.y = m->y; b.u = m->u; b.v = m->v; b.stride = MB; return b; }

static void lerp_block(const view_t *a, const view_t *b, int quarter, mbuf_t *out)
{
    for (int j = 0; j < MB; ++j)
    for (int i = 0; i < MB; ++i)
    {
        int32_t pa = a->y[j * a->stride + i], pb = b->y[j * b->stride + i];
        out->y[j * MB + i] = quarter ? lerp_quarter(pa, pb) : lerp_half(pa, pb);
    }
    int as = a->stride >> 1, bs = b->stride >> 1;
    for (int j = 0; j < 8; ++j)
    for (int i = 0; i < 8; ++i)
    {
        out->u[j * 8 + i] = quarter ? lerp_quarter(a->u[j * as + i], b->u[j * bs + i]) : lerp_half(a->u[j * as + i], b->u[j * bs + i]);
        out->v[j * 8 + i] = quarter ? lerp_quarter(a->v[j * as + i], b->v[j * bs + i]) : lerp_half(a->v[j * as + i], b->v[j * bs + i]);
    }
}

/* ------------------------------------------------------------------ motion search */

typedef struct
{
    int best_x, best_y;            /* int16 in the reference; frame coordinates always fit */
    int32_t best_sad, best_mad, best_ssd;
    int sp_index, sp_amount, sp_enabled;
} select_t;

typedef struct { const planes_t *pred; int thr; int px, py; } sparams_t;

/* motion.cpp:61-84 */
static int frac_index(int i, int j)
{
    i++; j++;
    if (j == 0) return i;
    if (j == 1) return i == 0 ? 3 : 4;
    return i + 5;
}

/* motion.cpp:86-109 */
static void frac_direction(int idx, int *dx, int *dy)
{
    if (idx <= 2) { *dy = -1; *dx = idx - 1; }
    else if (idx == 3) { *dx = -1; *dy = 0; }
    else if (idx == 4) { *dx = 1; *dy = 0; }
    else { *dy = 1; *dx = idx - 6; }
}

/* motion.cpp:111-149 -- note the precedence of the second rule:
 * sad<best || (sad==best && ssd<best_ssd && sad<8192) || mad<thr */
static void eval_fullpel(evxo_ctx *c, int x, int y, const sparams_t *sp, const view_t *src, select_t *s)
{
    view_t t = view_at(sp->pred, c->w, x, y);
    int32_t sad = block_sad(src, &t);
    int32_t ssd = (x - sp->px) * (x - sp->px) + (y - sp->py) * (y - sp->py);
    int32_t mad = block_mad(src, &t);
    c->n_fullpel++;
    int take;
    if (s->best_mad < sp->thr)
        take = mad < s->best_mad || (mad == s->best_mad && ssd < s->best_ssd);
    else
        take = sad < s->best_sad || (sad == s->best_sad && ssd < s->best_ssd && (uint32_t) sad < SAD_CAP) || mad < sp->thr;
    if (take) { s->best_x = x; s->best_y = y; s->best_sad = sad; s->best_ssd = ssd; s->best_mad = mad; }
}

/* motion.cpp:151-223: half first, then quarter; neither moves best_x/best_y */
static void eval_subpel(evxo_ctx *c, int tx, int ty, int i, int j, const sparams_t *sp, const view_t *src,
                        const view_t *best, select_t *s)
{
    view_t t = view_at(sp->pred, c->w, tx, ty);
    mbuf_t tmp;
    view_t tv = view_of(&tmp);
    for (int quarter = 0; quarter < 2; ++quarter)
    {
        lerp_block(best, &t, quarter, &tmp);
        int32_t sad = block_sad(src, &tv);
        int32_t mad = block_mad(src, &tv);
        c->n_subpel++;
        int take;
        if (s->best_mad < sp->thr) take = mad < s->best_mad;
        else take = (sad < s->best_sad && (uint32_t) sad < SAD_CAP) || mad < sp->thr;
        if (take)
        {
            s->sp_enabled = 1; s->sp_amount = quarter; s->sp_index = frac_index(i, j);
            s->best_sad = sad; s->best_mad = mad;
        }
    }
}

/* motion.cpp:225-275: one 3x3 (or 3x3-of-a-rectangle) round around the best of the
 * START of the round; intra rounds additionally require y<=py-16 || x<=px-16 */
static void search_round(evxo_ctx *c, int left, int top, int right, int bottom, int step, int intra,
                         const sparams_t *sp, const view_t *src, select_t *s)
{
    int bx = s->best_x, by = s->best_y;
    for (int j = top; j <= bottom; j += step)
    for (int i = left; i <= right; i += step)
    {
        int x = bx + i, y = by + j;
        if (intra && y > sp->py - MB && x > sp->px - MB) continue;
        if (x < 0 || x > c->w - MB || y < 0 || y > c->h - MB) continue;
        eval_fullpel(c, x, y, sp, src, s);
    }
}

/* motion.cpp:277-352 */
static void search_subpel(evxo_ctx *c, int intra, const sparams_t *sp, const view_t *src, select_t *s)
{
    view_t best = view_at(sp->pred, c->w, s->best_x, s->best_y);
    s->sp_index = 0; s->sp_amount = 0; s->sp_enabled = 0;
    for (int j = -1; j <= 1; ++j)
    for (int i = -1; i <= 1; ++i)
    {
        int tx = s->best_x + i, ty = s->best_y + j;
        if (i == 0 && j == 0) continue;
        if (intra && ty > sp->py - MB && tx > sp->px - MB) continue;
        if (tx < 0 || tx > c->w - MB || ty < 0 || ty > c->h - MB) continue;
        eval_subpel(c, tx, ty, i, j, sp, src, &best, s);
    }
}

static void fill_desc(evxo_block_desc *d, int intra, int target, const select_t *s, const sparams_t *sp)
{
    /* clear_block_desc zeroes the first 8 bytes only (common.cpp:67-73); q_index and
     * variance keep whatever the caller's storage held.  Here they are left untouched. */
    int type = intra ? T_INTRA : 0;
    if (s->best_x != sp->px || s->best_y != sp->py || s->sp_enabled) type |= T_MOTION;
    if (s->best_mad < sp->thr) type |= T_COPY;
    d->block_type = type;
    d->prediction_target = (uint8_t) target;
    d->motion_x = (int16_t) (s->best_x - sp->px);
    d->motion_y = (int16_t) (s->best_y - sp->py);
    d->sp_pred = (uint8_t) s->sp_enabled;
    d->sp_amount = (uint8_t) s->sp_amount;
    d->sp_index = (uint8_t) s->sp_index;
}

/* motion.cpp:354-419 */
static int32_t intra_prediction(evxo_ctx *c, uint32_t index, int quality, const view_t *src, int px, int py, evxo_block_desc *out)
{
    select_t s = { px, py, block_sad_self(src), INT32_BIG, INT32_BIG, 0, 0, 0 };
    sparams_t sp = { &c->ring[ring_slot(c, index, 0)], (quality >> 2) + 1, px, py };
    search_round(c, -SEARCH_RADIUS, -(SEARCH_RADIUS << 1), SEARCH_RADIUS, 0, SEARCH_RADIUS, 1, &sp, src, &s);
    for (int i = SEARCH_RADIUS >> 1; i > 0; i >>= 1) search_round(c, -i, -i, i, i, i, 1, &sp, src, &s);
    search_subpel(c, 1, &sp, src, &s);
    fill_desc(out, 1, 0, &s, &sp);
    return s.best_sad;
}

/* motion.cpp:421-494 */
static int32_t inter_prediction(evxo_ctx *c, uint32_t index, int quality, const view_t *src, int px, int py, int offset, evxo_block_desc *out)
{
    select_t s = { px, py, INT32_BIG, INT32_BIG, INT32_BIG, 0, 0, 0 };
    sparams_t sp = { &c->ring[ring_slot(c, index, offset)], (quality >> 2) + 1, px, py };
    view_t t = view_at(sp.pred, c->w, px, py);
    s.best_sad = block_sad(src, &t);
    s.best_mad = block_mad(src, &t);
    c->n_fullpel++;
    if (s.best_mad >= sp.thr)
    {
        for (int i = SEARCH_RADIUS; i > 0; i >>= 1) search_round(c, -i, -i, i, i, i, 0, &sp, src, &s);
        search_subpel(c, 0, &sp, src, &s);
    }
    fill_desc(out, 0, offset, &s, &sp);
    return s.best_sad;
}

int32_t evxo_inter_prediction(evxo_ctx *c, uint32_t index, int quality, int px, int py, int offset, evxo_block_desc *out)
{
    view_t src = view_at(&c->src, c->w, px, py);
    return inter_prediction(c, index, quality, &src, px, py, offset, out);
}

int32_t evxo_intra_prediction(evxo_ctx *c, uint32_t index, int quality, int px, int py, evxo_block_desc *out)
{
    view_t src = view_at(&c->src, c->w, px, py);
    return intra_prediction(c, index, quality, &src, px, py, out);
}

/* encode.cpp:17-67 */
static int32_t classify(evxo_ctx *c, int frame_type, uint32_t index, int quality, const view_t *src, int px, int py, evxo_block_desc *out)
{
    evxo_block_desc best = *out;      /* q_index / variance: stale storage, never meaningful for copy blocks */
    int32_t best_sad = intra_prediction(c, index, quality, src, px, py, &best);
    if (frame_type == 1)
    {
        for (int offset = 1; offset < c->R; ++offset)
        {
            evxo_block_desc cand = best;
            int32_t sad = inter_prediction(c, index, quality, src, px, py, offset, &cand);
            int cc = (cand.block_type & T_COPY) != 0, bc = (best.block_type & T_COPY) != 0;
            if (cc != bc) { if (cc) { best = cand; best_sad = sad; } }
            else if (sad < best_sad) { best = cand; best_sad = sad; }
        }
    }
    *out = best;
    return best_sad;
}

/* ------------------------------------------------------------------ transform */

/* xftables.h:57-67: LUT[j*8+i] = round(128*cos((2i+1) j pi/16)), rebuilt from the eight
 * distinct magnitudes by the cosine's symmetries */
static int16_t g_lut[64];
static int g_lut_ready = 0;

static void build_lut(void)
{
    static const int16_t c16[9] = { 128, 126, 118, 106, 91, 71, 49, 25, 0 };   /* 128*cos(k*pi/16), k=0..8 */
    for (int j = 0; j < 8; ++j)
    for (int i = 0; i < 8; ++i)
    {
        int k = ((2 * i + 1) * j) & 31;
        int16_t v;
        if (k <= 8) v = c16[k];
        else if (k <= 16) v = (int16_t) -c16[16 - k];
        else if (k <= 24) v = (int16_t) -c16[k - 16];
        else v = c16[32 - k];
        g_lut[j * 8 + i] = v;
    }
    g_lut_ready = 1;
}

/* transform.cpp:264-284: scale AFTER the sum */
static void fdct_line(const int16_t *src, int sp, int16_t *dst, int dp)
{
    for (int i = 0; i < 8; ++i)
    {
        int32_t t = 0;
        for (int k = 0; k < 8; ++k) t += src[k * sp] * g_lut[i * 8 + k];
        t = i == 0 ? (t * 45) / 128 : t / 2;
        dst[i * dp] = (int16_t) rdiv(t, 128);
    }
}

/* transform.cpp:330-349: scale PER TERM */
static void idct_line(const int16_t *src, int sp, int16_t *dst, int dp, const int16_t *add, int ap)
{
    for (int i = 0; i < 8; ++i)
    {
        int32_t t = ((src[0] * g_lut[i]) * 45) / 128;
        for (int k = 1; k < 8; ++k) t += (src[k * sp] * g_lut[k * 8 + i]) / 2;
        t = rdiv(t, 128);
        dst[i * dp] = (int16_t) (add ? t + add[i * ap] : t);
    }
}

/* transform.cpp:286-301 (rows, then columns; int16 scratch between) and :435-452 (residual first) */
static void fdct8(const int16_t *src, int sp, const int16_t *sub, int subp, int16_t *dst, int dp)
{
    int16_t res[64], tmp[64];
    for (int j = 0; j < 8; ++j)
    for (int i = 0; i < 8; ++i) res[j * 8 + i] = (int16_t) (sub ? src[j * sp + i] - sub[j * subp + i] : src[j * sp + i]);
    for (int j = 0; j < 8; ++j) fdct_line(res + j * 8, 1, tmp + j * 8, 1);
    for (int j = 0; j < 8; ++j) fdct_line(tmp + j, 8, dst + j, dp);
}

/* transform.cpp:351-366 and :418-433 (columns, then rows; prediction added in the last pass) */
static void idct8(const int16_t *src, int sp, const int16_t *add, int ap, int16_t *dst, int dp)
{
    int16_t tmp[64];
    for (int j = 0; j < 8; ++j) idct_line(src + j, sp, tmp + j, 8, NULL, 0);
    for (int j = 0; j < 8; ++j) idct_line(tmp + j * 8, 1, dst + j * dp, 1, add ? add + j * ap : NULL, 1);
}

/* macroblock.h:265-295: luma = four 8x8 (transform.cpp:485-494, 572-594), chroma 8x8 each */
static void fdct_mb(const view_t *src, const view_t *sub, mbuf_t *out)
{
    for (int q = 0; q < 4; ++q)
    {
        int ox = (q & 1) * 8, oy = (q >> 1) * 8;
        fdct8(src->y + oy * src->stride + ox, src->stride, sub ? sub->y + oy * sub->stride + ox : NULL, sub ? sub->stride : 0,
              out->y + oy * MB + ox, MB);
    }
    fdct8(src->u, src->stride >> 1, sub ? sub->u : NULL, sub ? sub->stride >> 1 : 0, out->u, 8);
    fdct8(src->v, src->stride >> 1, sub ? sub->v : NULL, sub ? sub->stride >> 1 : 0, out->v, 8);
}

static void idct_mb(const mbuf_t *in, const view_t *add, const view_t *dst)
{
    for (int q = 0; q < 4; ++q)
    {
        int ox = (q & 1) * 8, oy = (q >> 1) * 8;
        idct8(in->y + oy * MB + ox, MB, add ? add->y + oy * add->stride + ox : NULL, add ? add->stride : 0,
              dst->y + oy * dst->stride + ox, dst->stride);
    }
    idct8(in->u, 8, add ? add->u : NULL, add ? add->stride >> 1 : 0, dst->u, dst->stride >> 1);
    idct8(in->v, 8, add ? add->v : NULL, add ? add->stride >> 1 : 0, dst->v, dst->stride >> 1);
}

/* ------------------------------------------------------------------ quantiser */

/* quantize.cpp:13-35 (MPEG-style weighting matrices) */
static const int16_t QM_INTRA[64] = {
     8, 17, 18, 19, 21, 23, 25, 27,   17, 18, 19, 21, 23, 25, 27, 28,
    20, 21, 22, 23, 24, 26, 28, 30,   21, 22, 23, 24, 26, 28, 30, 32,
    22, 23, 24, 26, 28, 30, 32, 35,   23, 24, 26, 28, 30, 32, 35, 38,
    25, 26, 28, 30, 32, 35, 38, 41,   27, 28, 30, 32, 35, 38, 41, 45 };
static const int16_t QM_INTER[64] = {
    16, 17, 18, 19, 20, 21, 22, 23,   17, 18, 19, 20, 21, 22, 23, 24,
    18, 19, 20, 21, 22, 23, 24, 25,   19, 20, 21, 22, 23, 24, 26, 27,
    20, 21, 22, 23, 25, 26, 27, 28,   21, 22, 23, 24, 26, 27, 28, 30,
    22, 23, 24, 26, 27, 28, 30, 31,   23, 24, 25, 27, 28, 30, 31, 33 };

/* quantize.cpp:37-55 */
static int16_t luma_dc_scale(int qp) { return (int16_t) (qp < 5 ? 8 : qp < 9 ? qp << 1 : qp < 25 ? qp + 8 : (qp << 1) - 16); }
static int16_t chroma_dc_scale(int qp) { return (int16_t) (qp < 5 ? 8 : qp < 25 ? (qp + 13) >> 1 : qp - 6); }

/* analysis.h:176-198: 16x16 luma coefficients, only element (0,0) skipped; int32 wrap */
static int32_t variance2(const int16_t *y)
{
    uint32_t sum = 0, sq = 0; int32_t count = 0;
    for (int k = 1; k < 256; ++k)
        if (y[k]) { int32_t t = y[k]; sum += (uint32_t) t; sq += (uint32_t) (t * t); count++; }
    if (count <= 0) return 0;
    int32_t ss = (int32_t) (sum * sum);
    return (int32_t) (sq - (uint32_t) rdiv(ss, count));
}

/* quantize.cpp:60-77 */
static uint8_t block_qp(int quality, const int16_t *y)
{
    uint32_t var = (uint32_t) variance2(y);
    uint8_t q = (uint8_t) quality;
    uint8_t idx = (uint8_t) clip16((int16_t) (ilog2(var) >> 1), 1, 31);
    if (idx > q) return (uint8_t) clip16((int16_t) (q + ((idx - q) >> 1)), 1, 31);
    if (idx < q) return (uint8_t) clip16((int16_t) (q - ((q - idx) >> 1)), 1, 31);
    return q;
}

/* quantize.cpp:79-180, one 8x8 block.  mode: 0 intra luma, 1 intra chroma, 2 inter */
static void quant8(const evxo_ctx *c, int mode, int qp, const int16_t *src, int sp, int16_t *dst, int dp)
{
    for (int j = 0; j < 8; ++j)
    for (int k = 0; k < 8; ++k)
    {
        int16_t s = src[j * sp + k], out;
        if (c->linear)
        {
            if (mode < 2) out = (int16_t) rdiv(s, qp << 1);                                   /* :131-144 */
            else { int16_t m = (int16_t) (iabs16(s) - (qp >> 1)); out = (int16_t) rdiv(m, qp << 1); out = (int16_t) (out * isign16(s)); }   /* :165-180 */
        }
        else if (mode < 2) out = (int16_t) rdiv(rdiv(s * 16, QM_INTRA[j * 8 + k]), qp << 1);  /* :79-129 */
        else { int16_t f = (int16_t) rdiv(s * 16, QM_INTER[j * 8 + k]); out = (int16_t) rdiv(f - isign16(f) * qp, qp << 1); }   /* :146-163 */
        dst[j * dp + k] = out;
    }
    if (!c->linear && mode < 2) dst[0] = (int16_t) rdiv(src[0], mode == 0 ? luma_dc_scale(qp) : chroma_dc_scale(qp));
}

/* quantize.cpp:182-243 */
static void dequant8(const evxo_ctx *c, int mode, int qp, const int16_t *src, int sp, int16_t *dst, int dp)
{
    for (int j = 0; j < 8; ++j)
    for (int k = 0; k < 8; ++k)
    {
        int16_t s = src[j * sp + k], out;
        if (c->linear)
        {
            out = 0;
            if (s) { int16_t modq = (int16_t) ((qp + 1) % 2); int16_t m = (int16_t) ((iabs16(s) << 1) + 1); out = (int16_t) (m * qp - modq); out = (int16_t) (out * isign16(s)); }   /* :214-231 */
        }
        else out = (int16_t) ((2 * s * (mode < 2 ? QM_INTRA : QM_INTER)[j * 8 + k] * qp) / 16);
        dst[j * dp + k] = out;
    }
    if (!c->linear && mode < 2) dst[0] = (int16_t) (src[0] * (mode == 0 ? luma_dc_scale(qp) : chroma_dc_scale(qp)));
}

/* quantize.cpp:357-379: the intra matrices apply to INTRA_DEFAULT only */
static void quant_mb(const evxo_ctx *c, int qp, int type, const mbuf_t *in, const view_t *dst)
{
    int intra = (type & T_INTRA) && !(type & T_MOTION);
    for (int q = 0; q < 4; ++q)
    {
        int ox = (q & 1) * 8, oy = (q >> 1) * 8;
        quant8(c, intra ? 0 : 2, qp, in->y + oy * MB + ox, MB, dst->y + oy * dst->stride + ox, dst->stride);
    }
    quant8(c, intra ? 1 : 2, qp, in->u, 8, dst->u, dst->stride >> 1);
    quant8(c, intra ? 1 : 2, qp, in->v, 8, dst->v, dst->stride >> 1);
}

static void dequant_mb(const evxo_ctx *c, int qp, int type, const view_t *src, mbuf_t *out)
{
    int intra = (type & T_INTRA) && !(type & T_MOTION);
    for (int q = 0; q < 4; ++q)
    {
        int ox = (q & 1) * 8, oy = (q >> 1) * 8;
        dequant8(c, intra ? 0 : 2, qp, src->y + oy * src->stride + ox, src->stride, out->y + oy * MB + ox, MB);
    }
    dequant8(c, intra ? 1 : 2, qp, src->u, src->stride >> 1, out->u, 8);
    dequant8(c, intra ? 1 : 2, qp, src->v, src->stride >> 1, out->v, 8);
}

/* ------------------------------------------------------------------ block engines */

/* the prediction a block type implies (encode.cpp:83-141, decode.cpp:27-135), always
 * materialised into scratch so that an in-frame source can never alias the target */
static int build_prediction(evxo_ctx *c, uint32_t index, const evxo_block_desc *d, int px, int py, mbuf_t *pred)
{
    int type = d->block_type;
    if (type == T_INTRA) return 0;                                   /* INTRA_DEFAULT: no prediction */
    int offset = (type & T_INTRA) ? 0 : d->prediction_target;
    const planes_t *pl = &c->ring[ring_slot(c, index, offset)];
    int mx = (type & T_MOTION) ? d->motion_x : 0, my = (type & T_MOTION) ? d->motion_y : 0;
    view_t base = view_at(pl, c->w, px + mx, py + my);
    if ((type & T_MOTION) && d->sp_pred)
    {
        int dx, dy;
        frac_direction(d->sp_index, &dx, &dy);
        view_t nb = view_at(pl, c->w, px + mx + dx, py + my + dy);
        lerp_block(&base, &nb, d->sp_amount, pred);                 /* macroblock.h:243-259 */
    }
    else
    {
        for (int j = 0; j < MB; ++j) memcpy(pred->y + j * MB, base.y + j * base.stride, MB * 2);
        for (int j = 0; j < 8; ++j) { memcpy(pred->u + j * 8, base.u + j * (base.stride >> 1), 16); memcpy(pred->v + j * 8, base.v + j * (base.stride >> 1), 16); }
    }
    return 1;
}

/* encode.cpp:69-163 */
static void encode_block(evxo_ctx *c, uint32_t index, int quality, const view_t *src, int px, int py, evxo_block_desc *d)
{
    if (d->block_type & T_COPY) return;
    mbuf_t pred, tc;
    view_t pv = view_of(&pred);
    int has_pred = build_prediction(c, index, d, px, py, &pred);
    fdct_mb(src, has_pred ? &pv : NULL, &tc);
    d->q_index = block_qp(quality, tc.y);
    d->variance = (int16_t) variance2(tc.y);
    view_t dst = view_at(&c->coef, c->w, px, py);
    quant_mb(c, d->q_index, d->block_type, &tc, &dst);
}

/* decode.cpp:15-144 */
static void decode_block(evxo_ctx *c, uint32_t index, const evxo_block_desc *d, int px, int py)
{
    mbuf_t pred, tc;
    view_t pv = view_of(&pred);
    view_t dst = view_at(&c->ring[ring_slot(c, index, 0)], c->w, px, py);
    int has_pred = build_prediction(c, index, d, px, py, &pred);
    if (d->block_type & T_COPY)
    {
        for (int j = 0; j < MB; ++j) memcpy(dst.y + j * dst.stride, pred.y + j * MB, MB * 2);
        for (int j = 0; j < 8; ++j) { memcpy(dst.u + j * (dst.stride >> 1), pred.u + j * 8, 16); memcpy(dst.v + j * (dst.stride >> 1), pred.v + j * 8, 16); }
        return;
    }
    view_t cv = view_at(&c->coef, c->w, px, py);
    dequant_mb(c, d->q_index, d->block_type, &cv, &tc);
    idct_mb(&tc, has_pred ? &pv : NULL, &dst);
}

static void encode_one(evxo_ctx *c, int frame_type, uint32_t index, int quality, int bx, int by)
{
    int px = bx * MB, py = by * MB;
    evxo_block_desc *d = &c->table[by * c->mbw + bx];
    view_t src = view_at(&c->src, c->w, px, py);
    classify(c, frame_type, index, quality, &src, px, py, d);
    encode_block(c, index, quality, &src, px, py, d);
    decode_block(c, index, d, px, py);            /* the encoder's reconstruction loop, encode.cpp:194-199 */
}

/* encode.cpp:165-203 */
int evxo_encode_slice(evxo_ctx *c, int frame_type, uint32_t index, int quality)
{
    if (!g_lut_ready) build_lut();
    for (int by = 0; by < c->mbh; ++by)
    for (int bx = 0; bx < c->mbw; ++bx) encode_one(c, frame_type, index, quality, bx, by);
    return 0;
}

/* the same work in wavefront order, step = bx + 3*by */
int evxo_encode_slice_wavefront(evxo_ctx *c, int frame_type, uint32_t index, int quality, int reverse)
{
    if (!g_lut_ready) build_lut();
    int steps = c->mbw + 3 * (c->mbh - 1);
    for (int s = 0; s < steps; ++s)
    {
        if (!reverse)
        {
            for (int by = 0; by < c->mbh; ++by) { int bx = s - 3 * by; if (bx >= 0 && bx < c->mbw) encode_one(c, frame_type, index, quality, bx, by); }
        }
        else
        {
            for (int by = c->mbh - 1; by >= 0; --by) { int bx = s - 3 * by; if (bx >= 0 && bx < c->mbw) encode_one(c, frame_type, index, quality, bx, by); }
        }
    }
    return 0;
}

/* decode.cpp:146-170 */
int evxo_decode_slice(evxo_ctx *c, int frame_type, uint32_t index)
{
    (void) frame_type;
    if (!g_lut_ready) build_lut();
    for (int by = 0; by < c->mbh; ++by)
    for (int bx = 0; bx < c->mbw; ++bx) decode_block(c, index, &c->table[by * c->mbw + bx], bx * MB, by * MB);
    return 0;
}

/* ------------------------------------------------------------------ deblocking */

/* deblock.cpp:13-27 */
static const int16_t ALPHA[32] = { 0, 0, 0, 0, 0, 0, 0, 1, 1, 1, 2, 2, 3, 3, 4, 5, 6, 7, 8, 9, 10, 12, 14, 16, 18, 20, 22, 24, 26, 29, 32, 35 };
static const int16_t BETA[32]  = { 0, 0, 0, 0, 0, 0, 0, 0, 1, 1, 1, 1, 2, 2, 2, 3, 3, 3, 4, 4, 4, 5, 5, 6, 6, 7, 7, 8, 8, 9, 10, 11 };

/* deblock.cpp:49-79 */
static void edge_params(const evxo_block_desc *a, const evxo_block_desc *b, int *qp, int *strength)
{
    int ac = (a->block_type & T_COPY) != 0, bc = (b->block_type & T_COPY) != 0;
    if (!ac && !bc) *qp = (a->q_index + b->q_index) >> 1;
    else if (!ac) *qp = a->q_index;
    else if (!bc) *qp = b->q_index;
    else *qp = 0;
    *strength = (ac && bc) ? 0 : (ac != bc) ? 1 : 2;
}

/* deblock.cpp:81-129; s[] holds p3 p2 p1 p0 q0 q1 q2 q3 at stride `st` */
static void filter8(int16_t *s, int st, int qp, int strength, int luma)
{
    int16_t p3 = s[0], p2 = s[st], p1 = s[2 * st], p0 = s[3 * st], q0 = s[4 * st], q1 = s[5 * st], q2 = s[6 * st], q3 = s[7 * st];
    int16_t d0 = (int16_t) iabs32(p0 - q0), d1 = (int16_t) iabs32(p1 - p0), d2 = (int16_t) iabs32(q1 - q0);
    if (d0 >= ALPHA[qp] || d1 >= BETA[qp] || d2 >= BETA[qp]) return;
    if (strength == 2)
    {
        s[3 * st] = (int16_t) rdiv(p2 + 2 * p1 + 2 * p0 + 2 * q0 + q1, 8);
        s[2 * st] = (int16_t) rdiv(p2 + p1 + p0 + q0, 4);
        s[4 * st] = (int16_t) rdiv(p1 + 2 * p0 + 2 * q0 + 2 * q1 + q2, 8);
        s[5 * st] = (int16_t) rdiv(p0 + q0 + q1 + q2, 4);
        if (luma)
        {
            s[st] = (int16_t) rdiv(2 * p3 + 3 * p2 + p1 + p0 + q0, 8);
            s[6 * st] = (int16_t) rdiv(2 * q3 + 3 * q2 + q1 + q0 + p0, 8);
        }
    }
    else if (strength == 1)
    {
        s[3 * st] = (int16_t) rdiv(((q0 + p0) * 4) + p1 - q1, 8);
        s[4 * st] = (int16_t) rdiv(((q0 + p0) * 4) + q1 - p1, 8);
        if (luma)
        {
            s[2 * st] = (int16_t) rdiv((p2 * 4) + (p0 * 2) + (q0 * 2), 8);
            s[5 * st] = (int16_t) rdiv((q2 * 4) + (q0 * 2) + (p0 * 2), 8);
        }
    }
}

/* vertical edge at column i, rows j..j+7 (deblock.cpp:131-152, 177-187) */
static void v_edge(const evxo_ctx *c, int16_t *img, int w, int mbs, int i, int j, int luma, int rows_from, int rows_to)
{
    int wb = w / mbs, qp, st;
    edge_params(&c->table[(uint16_t) ((i - 1) / mbs + (j / mbs) * wb)], &c->table[(uint16_t) (i / mbs + (j / mbs) * wb)], &qp, &st);
    if (!st) return;
    for (int r = rows_from; r < rows_to; ++r) filter8(img + (size_t) (j + r) * w + i - 4, 1, qp, st, luma);
}

/* horizontal edge at row j, columns i..i+7 (deblock.cpp:154-175, 189-199) */
static void h_edge(const evxo_ctx *c, int16_t *img, int w, int mbs, int i, int j, int luma, int cols_from, int cols_to)
{
    int wb = w / mbs, qp, st;
    edge_params(&c->table[(uint16_t) (i / mbs + ((j - 1) / mbs) * wb)], &c->table[(uint16_t) (i / mbs + (j / mbs) * wb)], &qp, &st);
    if (!st) return;
    for (int k = cols_from; k < cols_to; ++k) filter8(img + (size_t) (j - 4) * w + i + k, w, qp, st, luma);
}

/* deblock.cpp:201-254: in-place raster sweep in the reference's own order */
static void deblock_plane(const evxo_ctx *c, int16_t *img, int w, int h, int mbs, int luma)
{
    for (int i = 8; i < w; i += 8) v_edge(c, img, w, mbs, i, 0, luma, 0, 8);
    for (int j = 8; j < h; j += 8)
    {
        h_edge(c, img, w, mbs, 0, j, luma, 0, 8);
        for (int i = 8; i < w; i += 8)
        {
            h_edge(c, img, w, mbs, i, j, luma, 0, 8);
            v_edge(c, img, w, mbs, i, j, luma, 0, 8);
        }
    }
}

/* Same result, order used by the GPU: independent 8x8 tiles centred on the grid
 * crossings (rows j-4..j+3, columns i-4..i+3).  Inside a tile: the upper band's
 * vertical edge on rows j-4..j-1, then the horizontal edge on all 8 columns, then
 * the lower band's vertical edge on rows j..j+3.  No tile reads or writes outside
 * itself, so tiles may run in any order (or in parallel). */
static void deblock_plane_tiled(const evxo_ctx *c, int16_t *img, int w, int h, int mbs, int luma)
{
    for (int j = h; j >= 0; j -= 8)            /* deliberately bottom-up / right-to-left */
    for (int i = w; i >= 0; i -= 8)
    {
        if (i > 0 && i < w && j > 0) v_edge(c, img, w, mbs, i, j - 8, luma, 4, 8);
        if (j > 0 && j < h)
        {
            if (i > 0) h_edge(c, img, w, mbs, i - 8, j, luma, 4, 8);
            if (i < w) h_edge(c, img, w, mbs, i, j, luma, 0, 4);
        }
        if (i > 0 && i < w && j < h) v_edge(c, img, w, mbs, i, j, luma, 0, 4);
    }
}

void evxo_deblock(evxo_ctx *c, uint32_t index)
{
    if (!c->deblocking) return;
    planes_t *pl = &c->ring[ring_slot(c, index, 0)];
    deblock_plane(c, pl->p[0], c->w, c->h, MB, 1);
    deblock_plane(c, pl->p[1], c->w >> 1, c->h >> 1, MB >> 1, 0);       /* deblock.cpp:264-269 */
    deblock_plane(c, pl->p[2], c->w >> 1, c->h >> 1, MB >> 1, 0);
}

void evxo_deblock_tiled(evxo_ctx *c, uint32_t index)
{
    if (!c->deblocking) return;
    planes_t *pl = &c->ring[ring_slot(c, index, 0)];
    deblock_plane_tiled(c, pl->p[0], c->w, c->h, MB, 1);
    deblock_plane_tiled(c, pl->p[1], c->w >> 1, c->h >> 1, MB >> 1, 0);
    deblock_plane_tiled(c, pl->p[2], c->w >> 1, c->h >> 1, MB >> 1, 0);
}

/* ------------------------------------------------------------------ entropy: bits, Exp-Golomb, ABAC */

typedef struct { uint8_t *data; uint32_t cap_bits, wr, rd; } bits_t;

static void put_bit(bits_t *b, int v)      /* bitstream.cpp:181-200: LSB first inside each byte */
{
    if (b->wr >= b->cap_bits) return;
    uint8_t *p = &b->data[b->wr >> 3];
    int k = b->wr & 7;
    *p = (uint8_t) ((*p & ~(1u << k)) | ((unsigned) (v & 1) << k));
    b->wr++;
}

static int bits_empty(const bits_t *b) { return b->rd >= b->wr; }
static int get_bit(bits_t *b) { int v = (b->data[b->rd >> 3] >> (b->rd & 7)) & 1; b->rd++; return v; }

/* abac.h:61-68, abac.cpp:10-18 */
typedef struct { uint32_t e3, hist[2], value, low, high, mid; } abac_t;
#define AB_MAX  0xFFFFu
#define AB_HALF 0x7FFFu
#define AB_QTR  0x3FFFu
#define AB_3QTR (3u * AB_QTR)

static void abac_clear(abac_t *a) { a->low = 0; a->value = 0; a->e3 = 0; a->hist[0] = a->hist[1] = 1; a->high = AB_MAX; a->mid = AB_HALF; }   /* abac.cpp:59-78 */

static void abac_model(abac_t *a)          /* abac.cpp:80-95 */
{
    uint64_t range = a->high - a->low;
    a->mid = a->low + (uint32_t) (range * a->hist[0] / (a->hist[0] + a->hist[1]));
}

static void abac_emit(abac_t *a, bits_t *out, int bit)      /* write_bit + flush_inverse_bits, abac.cpp:156-178 */
{
    put_bit(out, bit);
    for (uint32_t i = 0; i < a->e3; ++i) put_bit(out, !bit);
    a->e3 = 0;
}

static void abac_encode_bit(abac_t *a, bits_t *out, int bit)   /* abac.cpp:97-121, 180-224 */
{
    abac_model(a);
    if (bit) a->low = a->mid + 1; else a->high = a->mid;
    a->hist[bit]++;
    for (;;)
    {
        if ((a->high & 0x8000u) == (a->low & 0x8000u))
        {
            uint32_t msb = (a->high >> 15) & 1;
            a->low -= 0x8000u * msb; a->high -= 0x8000u * msb;
            abac_emit(a, out, (int) msb);
        }
        else if (a->high <= AB_3QTR && a->low > AB_QTR) { a->high -= AB_QTR + 1; a->low -= AB_QTR + 1; a->e3++; }
        else break;
        a->high = ((a->high << 1) & AB_MAX) | 1;
        a->low = (a->low << 1) & AB_MAX;
    }
}

static void abac_finish(abac_t *a, bits_t *out)             /* abac.cpp:281-313 */
{
    a->e3++;
    abac_emit(a, out, a->low < AB_QTR ? 0 : 1);
    abac_clear(a);
}

static void abac_start_decode(abac_t *a, bits_t *in)        /* abac.cpp:398-420 */
{
    int bit = 0;
    abac_clear(a);
    for (int i = 0; i < 16; ++i) { if (!bits_empty(in)) bit = get_bit(in); a->value = (a->value << 1) | (uint32_t) bit; }
}

static int abac_decode_bit(abac_t *a, bits_t *in)           /* abac.cpp:123-154, 226-279 */
{
    int out = 0, bit = 0;
    abac_model(a);
    if (a->value >= a->low && a->value <= a->mid) { a->high = a->mid; a->hist[0]++; out = 0; }
    else if (a->value > a->mid && a->value <= a->high) { a->low = a->mid + 1; a->hist[1]++; out = 1; }
    for (;;)
    {
        if (a->high <= AB_HALF) { }
        else if (a->low > AB_HALF) { a->high -= AB_HALF + 1; a->low -= AB_HALF + 1; a->value -= AB_HALF + 1; }
        else if (a->high <= AB_3QTR && a->low > AB_QTR) { a->high -= AB_QTR + 1; a->low -= AB_QTR + 1; a->value -= AB_QTR + 1; }
        else break;
        if (!bits_empty(in)) bit = get_bit(in);
        a->high = ((a->high << 1) & AB_MAX) | 1;
        a->low = (a->low << 1) & AB_MAX;
        a->value = ((a->value << 1) & AB_MAX) | (uint32_t) bit;
    }
    return out;
}

/* golomb.cpp:8-91: Exp-Golomb, emitted zeros-first then the value MSB-first (the
 * reference stores the code bit-reversed and writes it LSB-first, same thing) */
static void code_unsigned(abac_t *a, bits_t *out, uint32_t v)
{
    uint32_t x = v + 1;
    int n = ilog2(x) + 1;
    for (int i = 0; i < n - 1; ++i) abac_encode_bit(a, out, 0);
    for (int i = n - 1; i >= 0; --i) abac_encode_bit(a, out, (int) ((x >> i) & 1));
}

static void code_signed(abac_t *a, bits_t *out, int16_t v)
{
    uint32_t x = v == 0 ? 1u : (((uint32_t) iabs32(v) << 1) | (v < 0 ? 1u : 0u));
    int n = ilog2(x) + 1;
    for (int i = 0; i < n - 1; ++i) abac_encode_bit(a, out, 0);
    for (int i = n - 1; i >= 0; --i) abac_encode_bit(a, out, (int) ((x >> i) & 1));
}

/* stream.cpp:292-436 */
static uint16_t read_code(abac_t *a, bits_t *in, int *nbits)
{
    int zeros = 0;
    int bit = abac_decode_bit(a, in);
    while (!bit && zeros < 40) { zeros++; bit = abac_decode_bit(a, in); }
    uint16_t r = 0;
    for (int i = 0; i < zeros + 1; ++i)
    {
        r = (uint16_t) ((r << 1) | (bit & 1));
        if (i < zeros) bit = abac_decode_bit(a, in);
    }
    *nbits = zeros + 1;
    return r;
}

static uint16_t decode_unsigned(abac_t *a, bits_t *in) { int n; return (uint16_t) (read_code(a, in, &n) - 1); }

static int16_t decode_signed(abac_t *a, bits_t *in)
{
    int n;
    int16_t r = (int16_t) read_code(a, in, &n);
    int16_t sign = (int16_t) (1 - 2 * (r & 1));
    r = (int16_t) (sign * ((r >> 1) & 0x7FFF));
    if (n + (n - 1) > 0x20) r = (int16_t) (r | 0x8000);
    return r;
}

/* scan.h:60-70: the classic 8x8 zig-zag, generated by walking the anti-diagonals */
static uint8_t g_zigzag[64];
static int g_zigzag_ready = 0;

static void build_zigzag(void)
{
    int n = 0;
    for (int s = 0; s < 15; ++s)
    {
        if (s & 1) { for (int y = (s < 8 ? 0 : s - 7); y <= (s < 8 ? s : 7); ++y) g_zigzag[n++] = (uint8_t) (y * 8 + (s - y)); }
        else       { for (int x = (s < 8 ? 0 : s - 7); x <= (s < 8 ? s : 7); ++x) g_zigzag[n++] = (uint8_t) ((s - x) * 8 + x); }
    }
    g_zigzag_ready = 1;
}

/* stream.cpp:550-581 + serialize.cpp:10-23 */
static void put_block8(abac_t *a, bits_t *out, const int16_t *src, int stride, int16_t last_dc)
{
    int16_t blk[64];
    for (int j = 0; j < 8; ++j) memcpy(blk + j * 8, src + j * stride, 16);
    blk[0] = (int16_t) (blk[0] - last_dc);
    int run = 63;
    for (; run >= 0; --run) if (blk[g_zigzag[run]]) break;
    run++;
    code_unsigned(a, out, (uint16_t) run);
    for (int k = 0; k < run; ++k) code_signed(a, out, blk[g_zigzag[k]]);
}

/* stream.cpp:583-605 + unserialize.cpp:10-22 */
static void get_block8(abac_t *a, bits_t *in, int16_t *dst, int stride, int16_t last_dc)
{
    int16_t blk[64];
    memset(blk, 0, sizeof(blk));
    uint16_t run = decode_unsigned(a, in);
    for (uint32_t k = 0; k < run && k < 64; ++k) blk[g_zigzag[k]] = decode_signed(a, in);
    blk[0] = (int16_t) (blk[0] + last_dc);
    for (int j = 0; j < 8; ++j) memcpy(dst + j * stride, blk + j * 8, 16);
}

/* serialize.cpp:36-123 / unserialize.cpp:36-121: one coefficient plane.  The DC
 * predictor reads the neighbour's stored coefficients even when that neighbour was a
 * copy block this frame and therefore holds an older frame's data. */
static void code_plane(evxo_ctx *c, abac_t *a, bits_t *bs, int comp, int decode)
{
    int w = comp ? c->w >> 1 : c->w, h = comp ? c->h >> 1 : c->h, bsz = comp ? 8 : 16;
    int16_t *img = c->coef.p[comp];
    int idx = 0;
    for (int j = 0; j < h; j += bsz)
    for (int i = 0; i < w; i += bsz)
    {
        const evxo_block_desc *d = &c->table[idx++];
        if (d->block_type & T_COPY) continue;
        int16_t last_dc = 0;
        if (i >= bsz) last_dc = img[(size_t) j * w + (i - 8)];
        else if (j >= bsz) last_dc = img[(size_t) (j - 8) * w + i];
        int16_t *b = img + (size_t) j * w + i;
        if (!comp)
        {
            /* serialize.cpp:25-34: the three later 8x8 blocks predict from DCs inside the macroblock */
            if (!decode)
            {
                put_block8(a, bs, b, w, last_dc);
                put_block8(a, bs, b + 8, w, b[0]);
                put_block8(a, bs, b + 8 * w, w, b[0]);
                put_block8(a, bs, b + 8 * w + 8, w, b[8 * w]);
            }
            else
            {
                get_block8(a, bs, b, w, last_dc);
                get_block8(a, bs, b + 8, w, b[0]);
                get_block8(a, bs, b + 8 * w, w, b[0]);
                get_block8(a, bs, b + 8 * w + 8, w, b[8 * w]);
            }
        }
        else if (!decode) put_block8(a, bs, b, w, last_dc);
        else get_block8(a, bs, b, w, last_dc);
    }
}

static int target_bits(const evxo_ctx *c) { return ilog2((uint32_t) (c->R & 0xFF)); }      /* serialize.cpp:179 */

/* serialize.cpp:156-340 */
uint32_t evxo_serialize_slice(evxo_ctx *c, uint8_t *out, uint32_t cap_bytes)
{
    if (!g_zigzag_ready) build_zigzag();
    bits_t bs = { out, cap_bytes * 8, 0, 0 };
    abac_t a;
    abac_clear(&a);
    int n = c->mbw * c->mbh;
    const evxo_block_desc *t = c->table;
    for (int i = 0; i < n; ++i) for (int k = 0; k < 3; ++k) abac_encode_bit(&a, &bs, (t[i].block_type >> k) & 1);
    for (int i = 0; i < n; ++i)
    {
        if (t[i].block_type & T_INTRA) continue;
        for (int k = 0; k < target_bits(c); ++k) abac_encode_bit(&a, &bs, (t[i].prediction_target >> k) & 1);
    }
    int16_t last = 0;
    for (int i = 0; i < n; ++i) if (t[i].block_type & T_MOTION) { code_signed(&a, &bs, (int16_t) (t[i].motion_x - last)); last = t[i].motion_x; }
    last = 0;
    for (int i = 0; i < n; ++i) if (t[i].block_type & T_MOTION) { code_signed(&a, &bs, (int16_t) (t[i].motion_y - last)); last = t[i].motion_y; }
    for (int i = 0; i < n; ++i) if (t[i].block_type & T_MOTION) abac_encode_bit(&a, &bs, t[i].sp_pred & 1);
    for (int i = 0; i < n; ++i) if ((t[i].block_type & T_MOTION) && t[i].sp_pred) abac_encode_bit(&a, &bs, t[i].sp_amount & 1);
    for (int i = 0; i < n; ++i) if ((t[i].block_type & T_MOTION) && t[i].sp_pred) for (int k = 0; k < 3; ++k) abac_encode_bit(&a, &bs, (t[i].sp_index >> k) & 1);
    last = 0;
    for (int i = 0; i < n; ++i) if (!(t[i].block_type & T_COPY)) { code_signed(&a, &bs, (int16_t) (t[i].q_index - last)); last = t[i].q_index; }
    for (int comp = 0; comp < 3; ++comp) code_plane(c, &a, &bs, comp, 0);
    abac_finish(&a, &bs);
    return bs.wr;
}

/* unserialize.cpp:123-341.  Fields the stream does not carry for a block keep their
 * previous-frame values, as in the reference (the table is persistent). */
int evxo_unserialize_slice(evxo_ctx *c, const uint8_t *in, uint32_t nbits)
{
    if (!g_zigzag_ready) build_zigzag();
    bits_t bs = { (uint8_t *) in, nbits, nbits, 0 };
    abac_t a;
    abac_start_decode(&a, &bs);
    int n = c->mbw * c->mbh;
    evxo_block_desc *t = c->table;
    for (int i = 0; i < n; ++i)
    {
        int v = t[i].block_type & ~7;
        for (int k = 0; k < 3; ++k) v |= abac_decode_bit(&a, &bs) << k;
        t[i].block_type = v;
    }
    for (int i = 0; i < n; ++i)
    {
        if (t[i].block_type & T_INTRA) continue;
        int nb = target_bits(c), v = t[i].prediction_target & ~((1 << nb) - 1);
        for (int k = 0; k < nb; ++k) v |= abac_decode_bit(&a, &bs) << k;
        t[i].prediction_target = (uint8_t) v;
    }
    int16_t last = 0;
    for (int i = 0; i < n; ++i) if (t[i].block_type & T_MOTION) { t[i].motion_x = (int16_t) (last + decode_signed(&a, &bs)); last = t[i].motion_x; }
    last = 0;
    for (int i = 0; i < n; ++i) if (t[i].block_type & T_MOTION) { t[i].motion_y = (int16_t) (last + decode_signed(&a, &bs)); last = t[i].motion_y; }
    for (int i = 0; i < n; ++i) if (t[i].block_type & T_MOTION) t[i].sp_pred = (uint8_t) ((t[i].sp_pred & 0xFE) | abac_decode_bit(&a, &bs));
    for (int i = 0; i < n; ++i) if ((t[i].block_type & T_MOTION) && t[i].sp_pred) t[i].sp_amount = (uint8_t) ((t[i].sp_amount & 0xFE) | abac_decode_bit(&a, &bs));
    for (int i = 0; i < n; ++i) if ((t[i].block_type & T_MOTION) && t[i].sp_pred)
    {
        int v = t[i].sp_index & ~7;
        for (int k = 0; k < 3; ++k) v |= abac_decode_bit(&a, &bs) << k;
        t[i].sp_index = (uint8_t) v;
    }
    last = 0;
    for (int i = 0; i < n; ++i) if (!(t[i].block_type & T_COPY)) { t[i].q_index = (uint8_t) (decode_signed(&a, &bs) + last); last = t[i].q_index; }
    for (int comp = 0; comp < 3; ++comp) code_plane(c, &a, &bs, comp, 1);
    return 0;
}
