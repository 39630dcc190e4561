// evx_device.cuh -- integer arithmetic contract and warp-level block metrics shared by
// the sm_100a kernels.  Every routine cites the reference line whose arithmetic it must
// reproduce bit for bit.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#define EVX_MB 16
#define EVX_SAD_CAP 8192u          // EVX_MOTION_SAD_THRESHOLD, motion.cpp:19
#define EVX_SEARCH_RADIUS 16       // motion.cpp:24
#define EVX_BIG 0x7FFFFFFF

enum { EVX_T_INTRA = 1, EVX_T_MOTION = 2, EVX_T_COPY = 4 };   // types.h:68-71

struct EvxGeom
{
    int w, h;        // 16-aligned luma plane size (evx1enc.cpp:79-80)
    int vw, vh;      // visible size
    int mbw, mbh;
};

struct EvxPlanes { int16_t *y, *u, *v; };

// ------------------------------------------------------------------ scalar helpers

// math.h:228-236
__device__ __forceinline__ int evx_rdiv(int n, int d)
{
    return ((n ^ d) < 0) ? (n - d / 2) / d : (n + d / 2) / d;
}

// rounded_div by a positive power of two 2^s (d/2 = 2^(s-1)); truncating division of the biased numerator
__device__ __forceinline__ int evx_rdiv_pow2(int n, int s)
{
    int half = 1 << (s - 1);
    int t = n < 0 ? n - half : n + half;
    // C division truncates toward zero
    return t < 0 ? -((-t) >> s) : (t >> s);
}

__device__ __forceinline__ int evx_tdiv_pow2(int n, int s) { return n < 0 ? -((-n) >> s) : (n >> s); }   // C '/' by 2^s

__device__ __forceinline__ int evx_ilog2(uint32_t v) { return v ? 31 - __clz(v) : 0; }                  // math.h:69-138
__device__ __forceinline__ int evx_clip(int v, int lo, int hi) { return v < lo ? lo : (v > hi ? hi : v); }
__device__ __forceinline__ int evx_abs16(int v) { return v == -32768 ? 32767 : (v < 0 ? -v : v); }          // math.h:197-203
__device__ __forceinline__ int evx_sign(int v) { return (v > 0) - (v < 0); }

// macroblock.h:203-241: (a+b+-1)/2 and (3a+b+-2)/4 with C truncation == add, bias away from zero, arithmetic shift
__device__ __forceinline__ int evx_lerp_half(int a, int b) { int t = a + b; return (t + 1 + (t >> 31)) >> 1; }
__device__ __forceinline__ int evx_lerp_quarter(int a, int b) { int t = 3 * a + b; return (t + 2 + (t >> 31)) >> 2; }

__device__ __forceinline__ int evx_lo16(uint32_t v) { return (int) (short) (v & 0xFFFFu); }
__device__ __forceinline__ int evx_hi16(uint32_t v) { return ((int) v) >> 16; }
__device__ __forceinline__ uint32_t evx_pack16(int lo, int hi) { return ((uint32_t) lo & 0xFFFFu) | ((uint32_t) hi << 16); }

// motion.cpp:61-84
__device__ __forceinline__ int evx_frac_index(int i, int j)
{
    i++; j++;
    if (j == 0) return i;
    if (j == 1) return i == 0 ? 3 : 4;
    return i + 5;
}

// motion.cpp:86-109
__device__ __forceinline__ void evx_frac_direction(int idx, int &dx, int &dy)
{
    if (idx <= 2) { dy = -1; dx = idx - 1; }
    else if (idx == 3) { dx = -1; dy = 0; }
    else if (idx == 4) { dx = 1; dy = 0; }
    else { dy = 1; dx = idx - 6; }
}

// ------------------------------------------------------------------ block descriptor (16 bytes, common.h:78-95)

struct __align__(16) EvxDesc
{
    uint32_t w0, w1, w2, w3;
    __device__ __forceinline__ int type() const { return (int) w0; }
    __device__ __forceinline__ int target() const { return (int) (w1 & 0xFF); }
    __device__ __forceinline__ int mx() const { return evx_hi16(w1); }
    __device__ __forceinline__ int my() const { return evx_lo16(w2); }
    __device__ __forceinline__ int sp_pred() const { return (int) ((w2 >> 16) & 0xFF); }
    __device__ __forceinline__ int sp_amount() const { return (int) ((w2 >> 24) & 0xFF); }
    __device__ __forceinline__ int sp_index() const { return (int) (w3 & 0xFF); }
    __device__ __forceinline__ int q_index() const { return (int) ((w3 >> 8) & 0xFF); }
    __device__ __forceinline__ void set_q(int q, int var) { w3 = (w3 & 0xFFu) | (((uint32_t) q & 0xFFu) << 8) | (((uint32_t) var & 0xFFFFu) << 16); }
};

__device__ __forceinline__ EvxDesc evx_make_desc(int type, int target, int mx, int my, int sp_pred, int sp_amount, int sp_index)
{
    EvxDesc d;
    d.w0 = (uint32_t) type;
    d.w1 = ((uint32_t) target & 0xFFu) | ((uint32_t) mx << 16);
    d.w2 = ((uint32_t) my & 0xFFFFu) | (((uint32_t) sp_pred & 0xFFu) << 16) | (((uint32_t) sp_amount & 0xFFu) << 24);
    d.w3 = (uint32_t) sp_index & 0xFFu;
    return d;
}

// ------------------------------------------------------------------ search state (motion.cpp:47-59)

struct EvxSel
{
    int bx, by;              // best full-pel position, frame coordinates
    int sad, mad, ssd;
    int sp_index, sp_amount, sp_enabled;
};

// ------------------------------------------------------------------ closed-form acceptance
//
// The reference accepts the candidates of a round one after the other (motion.cpp:111-149):
//   holding a copy candidate (best_mad < thr):  take iff mad < best_mad || (mad == best_mad && ssd < best_ssd)
//   otherwise: take iff sad < best_sad || (sad == best_sad && ssd < best_ssd && sad < 8192) || mad < thr
// (operator precedence as written in the source); sub-pel tests (motion.cpp:151-223):
//   copy mode: take iff mad < best_mad;  otherwise: take iff (sad < best_sad && sad < 8192) || mad < thr.
// With each candidate as sortable keys that fold has a closed form:
//   key1 = sad:ssd'  (ssd' = 4095 = "infinite" when sad >= 8192: that disables the tie rule exactly
//                     where the reference's `&& sad < 8192` does),   key2 = mad:ssd,
//   * state already in copy mode (best_mad < thr): every take needs (mad,ssd) < (best_mad,best_ssd)
//     -> survivor = FIRST minimum of key2, if it beats the state;
//   * else, if some legal candidate has mad < thr: the first such is taken unconditionally and flips
//     the state to copy mode; later ones need a smaller key2 -> survivor = first minimum of key2 among
//     the candidates with mad < thr (the others can never be smaller);
//   * else every take needs (sad,ssd') < (best_sad,best_ssd) -> survivor = first minimum of key1.
// ssd <= 32^2 + 48^2 = 3328 < 4095 for every reachable position; sad, mad are clamped to 20 bits.
struct EvxKeys { uint32_t k1, k2; bool legal, lt; };

__device__ __forceinline__ EvxKeys evx_make_keys(int sad, int mad, int ssd, int thr, bool legal)
{
    EvxKeys k;
    const uint32_t z = (uint32_t) min(ssd, 4095);
    k.k1 = ((uint32_t) min(sad, 0xFFFFE) << 12) | ((uint32_t) sad < EVX_SAD_CAP ? z : 4095u);
    k.k2 = ((uint32_t) min(mad, 0xFFFFE) << 12) | z;
    k.legal = legal;
    k.lt = legal && mad < thr;
    return k;
}

// Lane c holds candidate c (in the reference's visiting order); lanes without a candidate pass
// legal = false.  Returns the surviving candidate's lane, or -1 when the state stands.
__device__ __forceinline__ int evx_select_fullpel(const EvxSel &s, const EvxKeys &k, int lane, int thr, uint32_t &n_legal)
{
    const unsigned legal_mask = __ballot_sync(0xFFFFFFFFu, k.legal);
    const unsigned lt_mask = __ballot_sync(0xFFFFFFFFu, k.lt);
    n_legal += __popc(legal_mask);
    const bool copy0 = s.mad < thr;
    uint32_t key, init;
    unsigned elig;
    if (copy0) { key = k.k2; elig = legal_mask; init = ((uint32_t) min(s.mad, 0xFFFFE) << 12) | (uint32_t) min(s.ssd, 4095); }
    else if (lt_mask) { key = k.k2; elig = lt_mask; init = 0xFFFFFFFFu; }
    else { key = k.k1; elig = legal_mask; init = ((uint32_t) min(s.sad, 0xFFFFE) << 12) | (uint32_t) min(s.ssd, 4095); }
    const uint32_t mykey = ((elig >> lane) & 1u) ? key : 0xFFFFFFFFu;
    const uint32_t m = __reduce_min_sync(0xFFFFFFFFu, mykey);
    if (m >= init) return -1;
    return __ffs(__ballot_sync(0xFFFFFFFFu, mykey == m)) - 1;
}

// Sub-pel tests (motion.cpp:151-223), lane t = test t in reference order (direction-major, half
// before quarter): copy mode -> first minimum of mad; else a test with mad < thr exists -> first
// minimum of mad among those; else -> first minimum of sad among tests with sad < 8192.
__device__ __forceinline__ int evx_select_subpel(const EvxSel &s, int sad, int mad, bool legal, int lane, int thr, uint32_t &n_legal)
{
    const unsigned legal_mask = __ballot_sync(0xFFFFFFFFu, legal);
    const unsigned lt_mask = __ballot_sync(0xFFFFFFFFu, legal && mad < thr);
    const unsigned cap_mask = __ballot_sync(0xFFFFFFFFu, legal && (uint32_t) sad < EVX_SAD_CAP);
    n_legal += __popc(legal_mask);
    const bool copy0 = s.mad < thr;
    int key, init;
    unsigned elig;
    if (copy0) { key = mad; elig = legal_mask; init = s.mad; }
    else if (lt_mask) { key = mad; elig = lt_mask; init = EVX_BIG; }
    else { key = sad; elig = cap_mask; init = s.sad; }
    const int mykey = ((elig >> lane) & 1u) ? key : EVX_BIG;
    const int m = __reduce_min_sync(0xFFFFFFFFu, mykey);
    if (m >= init) return -1;
    return __ffs(__ballot_sync(0xFFFFFFFFu, mykey == m)) - 1;
}

__device__ __forceinline__ EvxDesc evx_desc_from_sel(const EvxSel &s, int intra, int target, int px, int py, int thr)
{
    int type = intra ? EVX_T_INTRA : 0;
    if (s.bx != px || s.by != py || s.sp_enabled) type |= EVX_T_MOTION;
    if (s.mad < thr) type |= EVX_T_COPY;
    return evx_make_desc(type, target, s.bx - px, s.by - py, s.sp_enabled, s.sp_amount, s.sp_index);
}

// ------------------------------------------------------------------ warp-level block metrics
//
// A search window lives in shared memory as packed int16 pairs (one 32-bit word = two
// horizontally adjacent samples).  A warp evaluates one 16x16+8x8+8x8 candidate:
//   luma   lane -> (rr = lane>>3, cc = lane&7): words (row rr+4k, pixel pair cc), k=0..3
//   chroma lane -> (row lane>>2, pixel pair lane&3), one word of U and one of V
// With a row pitch of 8 (mod 32) words for luma and 4*odd (mod 32) for chroma every one
// of these loads is bank-conflict free.

struct EvxWin
{
    const uint32_t *y, *u, *v;   // shared memory
    int pw_y, pw_c;              // row pitch in 32-bit words
    int ox, oy;                  // frame coordinates of the window's first luma sample (ox even)
    int cox, coy;                // same for chroma (cox even)
};

struct EvxLaneBlock { uint32_t w[6]; };   // 4 luma words, U word, V word of this lane

__device__ __forceinline__ void evx_load_block(const EvxWin &win, int x, int y, int lane, EvxLaneBlock &b)
{
    int wx = x - win.ox, wy = y - win.oy;
    const uint32_t *p = win.y + (wy + (lane >> 3)) * win.pw_y + (wx >> 1) + (lane & 7);
    if (wx & 1)
    {
#pragma unroll
        for (int k = 0; k < 4; ++k) b.w[k] = __byte_perm(p[4 * k * win.pw_y], p[4 * k * win.pw_y + 1], 0x5432);
    }
    else
    {
#pragma unroll
        for (int k = 0; k < 4; ++k) b.w[k] = p[4 * k * win.pw_y];
    }
    int cx = (x >> 1) - win.cox, cy = (y >> 1) - win.coy;
    int off = (cy + (lane >> 2)) * win.pw_c + (cx >> 1) + (lane & 3);
    if (cx & 1)
    {
        b.w[4] = __byte_perm(win.u[off], win.u[off + 1], 0x5432);
        b.w[5] = __byte_perm(win.v[off], win.v[off + 1], 0x5432);
    }
    else
    {
        b.w[4] = win.u[off];
        b.w[5] = win.v[off];
    }
}

// Same lane layout against a window that is CIRCULAR in x (K3): rows of 64 luma words
// (128 samples) / 32 chroma words (64 samples) addressed by absolute frame x modulo the
// ring, row pitch padded to 72 / 36 words so that the loads stay bank-conflict free.
struct EvxRingWin
{
    const uint32_t *y, *u, *v;
    int oy, coy;                 // frame row of window row 0 (luma / chroma)
};
#define EVX_RING_PWY 72
#define EVX_RING_PWC 36

// Both neighbouring words are always fetched and the parity picks the byte selector: no branch,
// so several candidates can be scheduled in one basic block.
__device__ __forceinline__ void evx_load_block_ring_bf(const EvxRingWin &win, int x, int y, int lane, EvxLaneBlock &b)
{
    const uint32_t *row = win.y + (y - win.oy + (lane >> 3)) * EVX_RING_PWY;
    const int w0 = (x >> 1) + (lane & 7);
    const uint32_t sel = (x & 1) ? 0x5432u : 0x3210u;
#pragma unroll
    for (int k = 0; k < 4; ++k)
        b.w[k] = __byte_perm(row[4 * k * EVX_RING_PWY + (w0 & 63)], row[4 * k * EVX_RING_PWY + ((w0 + 1) & 63)], sel);
    const int cx = x >> 1;
    const int crow = ((y >> 1) - win.coy + (lane >> 2)) * EVX_RING_PWC;
    const int c0 = (cx >> 1) + (lane & 3);
    const uint32_t csel = (cx & 1) ? 0x5432u : 0x3210u;
    b.w[4] = __byte_perm(win.u[crow + (c0 & 31)], win.u[crow + ((c0 + 1) & 31)], csel);
    b.w[5] = __byte_perm(win.v[crow + (c0 & 31)], win.v[crow + ((c0 + 1) & 31)], csel);
}

// The same for positions whose parity is known when the code is written: the full-pel rounds with steps 16, 8 and 4 only
// ever name positions with x = 0 (mod 4) (macroblock origin + multiples of the step), the step-2 round x = 0 (mod 2): the
// words are aligned and the neighbouring word and the byte permute are not needed.
template <bool LUMA_ONLY_ALIGNED>      // false: luma and chroma words aligned (x % 4 == 0); true: luma aligned, chroma by parity
__device__ __forceinline__ void evx_load_block_ring_al(const EvxRingWin &win, int x, int y, int lane, EvxLaneBlock &b)
{
    const uint32_t *row = win.y + (y - win.oy + (lane >> 3)) * EVX_RING_PWY;
    const int w0 = (x >> 1) + (lane & 7);
#pragma unroll
    for (int k = 0; k < 4; ++k) b.w[k] = row[4 * k * EVX_RING_PWY + (w0 & 63)];
    const int cx = x >> 1;
    const int crow = ((y >> 1) - win.coy + (lane >> 2)) * EVX_RING_PWC;
    const int c0 = (cx >> 1) + (lane & 3);
    if (LUMA_ONLY_ALIGNED)
    {
        const uint32_t csel = (cx & 1) ? 0x5432u : 0x3210u;
        b.w[4] = __byte_perm(win.u[crow + (c0 & 31)], win.u[crow + ((c0 + 1) & 31)], csel);
        b.w[5] = __byte_perm(win.v[crow + (c0 & 31)], win.v[crow + ((c0 + 1) & 31)], csel);
    }
    else
    {
        b.w[4] = win.u[crow + (c0 & 31)];
        b.w[5] = win.v[crow + (c0 & 31)];
    }
}

// The source macroblock of a warp, in the lane layout above: packed negation (for
// VIADDMNMX) and the lane's luma sum.
struct EvxLaneSrc
{
    uint32_t neg[6];     // -src, packed
    uint32_t pos[6];     // src, packed
    int lsum;            // sum of this lane's 8 luma samples
};

// Packed negation, halfword by halfword in scalar arithmetic (six words per macroblock: the cost is nil).
// The packed form (~v + 0x00010001 through add.u16x2) is what ptxas 12.9 rematerialises in split halves,
// dropping the immediate of one half (cairo_b200/build.py:lint_sass scans for the symptom).
__device__ __forceinline__ uint32_t evx_neg16x2(uint32_t v)
{
    return ((0u - (v & 0xFFFFu)) & 0xFFFFu) | ((0u - (v >> 16)) << 16);
}

__device__ __forceinline__ void evx_make_src(const EvxLaneBlock &b, EvxLaneSrc &s)
{
    s.lsum = 0;
#pragma unroll
    for (int k = 0; k < 6; ++k) { s.pos[k] = b.w[k]; s.neg[k] = evx_neg16x2(b.w[k]); }
#pragma unroll
    for (int k = 0; k < 4; ++k) s.lsum += evx_lo16(b.w[k]) + evx_hi16(b.w[k]);
}

// SAD over luma (analysis.h:42-55) and MAD over luma+chroma (analysis.h:103-125) of one
// candidate against the warp's source block.  Per packed word: three VIADDMNMX.S16x2
// (running max and min of ref-src, and relu(ref-src)) and two IDP.2A:
//   sum|d| = 2*sum relu(d) - sum d,   max|d| = max(max d, -min d).
// SAD only: one VIADDMNMX.S16x2.RELU and two IDP.2A per packed word.
__device__ __forceinline__ int evx_block_sad_lane(const EvxLaneBlock &ref, const EvxLaneSrc &src)
{
    int acc = src.lsum;
#pragma unroll
    for (int k = 0; k < 4; ++k)
    {
        uint32_t r = __viaddmax_s16x2_relu(ref.w[k], src.neg[k], 0u);
        acc = __dp2a_lo((int) r, 0x0202, acc);
        acc = __dp2a_lo((int) ref.w[k], 0xFFFF, acc);
    }
    return acc;
}

// MAD only (luma + chroma): two VIADDMNMX.S16x2 per packed word.
__device__ __forceinline__ int evx_block_mad_lane(const EvxLaneBlock &ref, const EvxLaneSrc &src)
{
    uint32_t amx = 0x80008000u, amn = 0x7FFF7FFFu;
#pragma unroll
    for (int k = 0; k < 6; ++k)
    {
        amx = __viaddmax_s16x2(ref.w[k], src.neg[k], amx);
        amn = __viaddmin_s16x2(ref.w[k], src.neg[k], amn);
    }
    return max(max(evx_lo16(amx), evx_hi16(amx)), -min(evx_lo16(amn), evx_hi16(amn)));
}

// The MAD feeds the reference's decisions only through `mad < thr` and, once a copy candidate
// is held, through comparisons among values below thr (motion.cpp:123-140, 165-213): its exact
// value never matters when it is >= thr.  And mad < thr implies every one of the 256 luma
// differences is below thr, i.e. sad < 256*thr.  So the MAD pass is run only for candidates
// whose SAD is under that bound; all others report EVX_BIG.  (Bit-exact by construction; the
// parity tests compare every descriptor and SAD with the oracle.)
__device__ __forceinline__ void evx_block_cost(const EvxLaneBlock &ref, const EvxLaneSrc &src, int thr, int &sad, int &mad)
{
    sad = __reduce_add_sync(0xFFFFFFFFu, evx_block_sad_lane(ref, src));
    mad = EVX_BIG;
    if (sad < 256 * thr) mad = __reduce_max_sync(0xFFFFFFFFu, evx_block_mad_lane(ref, src));
}

// Latency-optimised form for the serial kernel (K3): both reductions are issued back to back
// instead of making the MAD pass wait for the SAD's warp reduction.
__device__ __forceinline__ void evx_block_cost_both(const EvxLaneBlock &ref, const EvxLaneSrc &src, int &sad, int &mad)
{
    const int a = evx_block_sad_lane(ref, src), m = evx_block_mad_lane(ref, src);
    sad = __reduce_add_sync(0xFFFFFFFFu, a);
    mad = __reduce_max_sync(0xFFFFFFFFu, m);
}

// The exact scalar form of one sub-pel direction for this lane's twelve samples (samples outside [0, 16383]: never in real
// video).  A call: inlined into every caller it doubled the size of the sub-pel code on the hot path.
__device__ __noinline__ int4 evx_subpel_exact_lane(EvxLaneBlock best, EvxLaneBlock nb, EvxLaneBlock pos)
{
    int sh = 0, mh = 0, sq = 0, mq = 0;
#pragma unroll 1
    for (int k = 0; k < 6; ++k)
    {
#pragma unroll
        for (int half = 0; half < 2; ++half)
        {
            int a = half ? evx_hi16(best.w[k]) : evx_lo16(best.w[k]);
            int b = half ? evx_hi16(nb.w[k]) : evx_lo16(nb.w[k]);
            int s = half ? evx_hi16(pos.w[k]) : evx_lo16(pos.w[k]);
            // the blended sample is stored as int16 before it is compared (macroblock.h:210, 230)
            int dh = abs(s - (int) (short) evx_lerp_half(a, b));
            int dq = abs(s - (int) (short) evx_lerp_quarter(a, b));
            if (k < 4) { sh += dh; sq += dq; }
            mh = max(mh, dh);
            mq = max(mq, dq);
        }
    }
    return make_int4(sh, mh, sq, mq);
}

// One sub-pel direction: both the half- and the quarter-pel blend of `best` with its
// neighbour `nb` (macroblock.h:203-241), SAD/MAD of each against the source.
//
// Fast path (every sample of both blocks in [0, 16383], i.e. always for real video): the
// blends are computed on packed pairs.  For t >= 0 the reference's (t+1)/2 and (t+2)/4 are
// plain floors, so   half = (a+b+1) >> 1   and   quarter = a + floor((b-a+2)/4), the latter
// with a +0x4000 bias so the per-halfword shift can be a logical one.  Otherwise the exact
// scalar form (negative sums round away from zero) is used.
// thr < 0 selects the latency-optimised variant (every MAD computed, reductions overlapped).
__device__ __forceinline__ void evx_subpel_cost(const EvxLaneBlock &best, const EvxLaneBlock &nb, const EvxLaneSrc &src, int thr,
                                                int &sad_h, int &mad_h, int &sad_q, int &mad_q)
{
    uint32_t any = 0;
#pragma unroll
    for (int k = 0; k < 6; ++k) any |= best.w[k] | nb.w[k];
    int sh, mh, sq, mq;
    if (!__any_sync(0xFFFFFFFFu, (any & 0xC000C000u) != 0))
    {
        EvxLaneBlock hb, qb;
#pragma unroll
        for (int k = 0; k < 6; ++k)
        {
            const uint32_t a = best.w[k], b = nb.w[k];
            hb.w[k] = (__vadd2(__vadd2(a, b), 0x00010001u) >> 1) & 0x7FFF7FFFu;
            const uint32_t d = __vadd2(b, __vadd2(~a, 0x40034003u));            // b - a + 2 + 0x4000 per halfword
            qb.w[k] = __vadd2((d >> 2) & 0x3FFF3FFFu, __vadd2(a, 0xF000F000u));  // floor(d/4) - 0x1000 + a
        }
        if (thr < 0)
        {   // latency-optimised (K3): four independent reductions in flight
            const int a1 = evx_block_sad_lane(hb, src), a2 = evx_block_sad_lane(qb, src);
            const int m1 = evx_block_mad_lane(hb, src), m2 = evx_block_mad_lane(qb, src);
            sad_h = __reduce_add_sync(0xFFFFFFFFu, a1); sad_q = __reduce_add_sync(0xFFFFFFFFu, a2);
            mad_h = __reduce_max_sync(0xFFFFFFFFu, m1); mad_q = __reduce_max_sync(0xFFFFFFFFu, m2);
            return;
        }
        sad_h = __reduce_add_sync(0xFFFFFFFFu, evx_block_sad_lane(hb, src));
        sad_q = __reduce_add_sync(0xFFFFFFFFu, evx_block_sad_lane(qb, src));
        mad_h = EVX_BIG; mad_q = EVX_BIG;
        if (sad_h < 256 * thr) mad_h = __reduce_max_sync(0xFFFFFFFFu, evx_block_mad_lane(hb, src));
        if (sad_q < 256 * thr) mad_q = __reduce_max_sync(0xFFFFFFFFu, evx_block_mad_lane(qb, src));
        return;
    }
    else
    {
        EvxLaneBlock pos;
#pragma unroll
        for (int k = 0; k < 6; ++k) pos.w[k] = src.pos[k];
        const int4 r = evx_subpel_exact_lane(best, nb, pos);
        sh = r.x; mh = r.y; sq = r.z; mq = r.w;
    }
    sad_h = __reduce_add_sync(0xFFFFFFFFu, sh);
    mad_h = __reduce_max_sync(0xFFFFFFFFu, mh);
    sad_q = __reduce_add_sync(0xFFFFFFFFu, sq);
    mad_q = __reduce_max_sync(0xFFFFFFFFu, mq);
}

// ------------------------------------------------------------------ transform / quantiser tables

// xftables.h:57-67: LUT[j*8+i] = round(128 cos((2i+1) j pi / 16)), from the 8 magnitudes
__constant__ int16_t EVX_C16[9] = { 128, 126, 118, 106, 91, 71, 49, 25, 0 };     // (in constant memory: a local array indexed at run time lives on the stack)
__device__ __forceinline__ int evx_dct_lut(int j, int i)
{
    const int16_t *c16 = EVX_C16;
    int k = ((2 * i + 1) * j) & 31;
    if (k <= 8) return c16[k];
    if (k <= 16) return -c16[16 - k];
    if (k <= 24) return -c16[k - 16];
    return c16[32 - k];
}

// quantize.cpp:13-35
__constant__ int16_t EVX_QM_INTRA[64] = {
     8, 17, 18, 19, 21, 23, 25, 27,   17, 18, 19, 21, 23, 25, 27, 28,
    20, 21, 22, 23, 24, 26, 28, 30,   21, 22, 23, 24, 26, 28, 30, 32,
    22, 23, 24, 26, 28, 30, 32, 35,   23, 24, 26, 28, 30, 32, 35, 38,
    25, 26, 28, 30, 32, 35, 38, 41,   27, 28, 30, 32, 35, 38, 41, 45 };
__constant__ int16_t EVX_QM_INTER[64] = {
    16, 17, 18, 19, 20, 21, 22, 23,   17, 18, 19, 20, 21, 22, 23, 24,
    18, 19, 20, 21, 22, 23, 24, 25,   19, 20, 21, 22, 23, 24, 26, 27,
    20, 21, 22, 23, 25, 26, 27, 28,   21, 22, 23, 24, 26, 27, 28, 30,
    22, 23, 24, 26, 27, 28, 30, 31,   23, 24, 25, 27, 28, 30, 31, 33 };
// deblock.cpp:13-27
__constant__ int16_t EVX_ALPHA[32] = { 0, 0, 0, 0, 0, 0, 0, 1, 1, 1, 2, 2, 3, 3, 4, 5, 6, 7, 8, 9, 10, 12, 14, 16, 18, 20, 22, 24, 26, 29, 32, 35 };
__constant__ int16_t EVX_BETA[32]  = { 0, 0, 0, 0, 0, 0, 0, 0, 1, 1, 1, 1, 2, 2, 2, 3, 3, 3, 4, 4, 4, 5, 5, 6, 6, 7, 7, 8, 8, 9, 10, 11 };

// quantize.cpp:37-55
__device__ __forceinline__ int evx_luma_dc_scale(int qp) { return qp < 5 ? 8 : qp < 9 ? qp << 1 : qp < 25 ? qp + 8 : (qp << 1) - 16; }
__device__ __forceinline__ int evx_chroma_dc_scale(int qp) { return qp < 5 ? 8 : qp < 25 ? (qp + 13) >> 1 : qp - 6; }

// quantize.cpp:182-243
__device__ __forceinline__ int evx_dequant(int s, int pos, int mode, int qp, int linear, const int16_t *qm_intra, const int16_t *qm_inter)
{
    int out;
    if (linear)
    {
        out = 0;
        if (s) { int modq = (qp + 1) % 2; int m = (short) ((evx_abs16(s) << 1) + 1); out = (short) (m * qp - modq); out = (short) (out * evx_sign(s)); }
    }
    else if (mode < 2 && pos == 0) out = (short) (s * (mode == 0 ? evx_luma_dc_scale(qp) : evx_chroma_dc_scale(qp)));
    else out = (short) evx_tdiv_pow2(2 * s * (mode < 2 ? qm_intra[pos] : qm_inter[pos]) * qp, 4);
    return out;
}
