// entropy.cpp -- see entropy.h.
#include "entropy.h"

#include <string.h>

#include <algorithm>
#include <atomic>
#include <chrono>
#include <mutex>
#include <stdio.h>
#include <stdlib.h>

namespace evx {

namespace {

enum { T_INTRA = 1, T_MOTION = 2, T_COPY = 4 };      // types.h:68-71

const uint32_t AB_MAX = 0xFFFFu, AB_HALF = 0x7FFFu, AB_QTR = 0x3FFFu, AB_3QTR = 3u * 0x3FFFu;   // abac.cpp:4-10

inline int bit_length(uint32_t v) { return 32 - __builtin_clz(v); }

// 8x8 zig-zag (scan.h:60-70), built by walking the anti-diagonals; entries are offsets into a
// stride-16 (luma record) or stride-8 (chroma record) block
struct zigzag_tables
{
    uint8_t pos[64];          // row*8+col
    uint8_t luma[64];         // row*16+col
    zigzag_tables()
    {
        int n = 0;
        for (int s = 0; s < 15; ++s)
        {
            int lo = s < 8 ? 0 : s - 7, hi = s < 8 ? s : 7;
            for (int k = lo; k <= hi; ++k)
            {
                int row = (s & 1) ? k : s - k, col = s - row;
                pos[n] = (uint8_t) (row * 8 + col);
                luma[n] = (uint8_t) (row * 16 + col);
                ++n;
            }
        }
    }
};
const zigzag_tables ZZ;

// ---------------------------------------------------------------- encoder side

// The model divides by the number of symbols seen so far + 2 (abac.cpp:80-95), a divisor that
// simply counts up.  floor(a / t) for a < 2^40 is the high half of a * ceil(2^64 / t) exactly
// (error term a * (M*t - 2^64) < 2^40 * 2^24 < 2^64), so one table of reciprocals, shared by
// every coder of the process and grown on demand, replaces the hardware divide on the bin path.
class reciprocal_table
{
    static const size_t kMax = size_t(1) << 23;      // (the fast loop takes tot < 2^23 only; address space, touched on demand)
    uint64_t *m_;
    std::atomic<size_t> size_;
    std::mutex lock_;

public:
    reciprocal_table() : m_(new uint64_t[kMax]), size_(0) { grow(size_t(1) << 18); }
    const uint64_t *data() const { return m_; }
    size_t size() const { return size_.load(std::memory_order_acquire); }
    void grow(size_t want)
    {
        std::lock_guard<std::mutex> g(lock_);
        size_t have = size_.load(std::memory_order_relaxed);
        if (want > kMax) want = kMax;
        for (size_t t = have; t < want; ++t)
            m_[t] = t < 2 ? 0 : (uint64_t) ((((unsigned __int128) 1 << 64) + t - 1) / t);
        if (want > have) size_.store(want, std::memory_order_release);
    }
};
reciprocal_table &recips() { static reciprocal_table t; return t; }

// bit reversal of a byte (the stream is LSB-first, the coder's bits leave MSB-first)
struct rev8_table { uint8_t v[256]; rev8_table() { for (int i = 0; i < 256; ++i) { int r = 0; for (int b = 0; b < 8; ++b) if (i & (1 << b)) r |= 0x80 >> b; v[i] = (uint8_t) r; } } };
const rev8_table REV8_T;
#define REV8 REV8_T.v

// Phase 1 of a slice: everything the coder will be fed, as one string of bins (bit i of the
// string = the i-th call of the reference's encode_symbol).  Binarisation is independent per
// syntax element, so it is a handful of shifts per CODE instead of a coder step per bin.
class bin_string
{
    std::vector<uint64_t> &w_;
    size_t pos_;
    uint64_t acc_;
    uint32_t n_;

public:
    explicit bin_string(std::vector<uint64_t> &words) : w_(words), pos_(0), acc_(0), n_(0) {}

    // room for at least `bins` more bins
    inline void reserve(size_t bins)
    {
        const size_t need = pos_ + (bins >> 6) + 3;
        if (need > w_.size()) w_.resize(need * 2);
    }
    // n <= 40 bins, bit 0 of v first
    inline void put(uint64_t v, uint32_t n)
    {
        acc_ |= v << n_;
        if (n_ + n >= 64)
        {
            w_[pos_++] = acc_;
            acc_ = n_ ? v >> (64 - n_) : 0;
            n_ = n_ + n - 64;
        }
        else n_ += n;
    }
    inline void put_bits_lsb(uint32_t v, int n) { put(v & ((1u << n) - 1u), (uint32_t) n); }
    // Exp-Golomb (golomb.cpp:8-91): n-1 zeros, then the n bits of x, most significant first
    inline void put_code(uint32_t x)
    {
        const uint32_t n = (uint32_t) bit_length(x);
        uint32_t r = x;                                   // reverse the low n bits
        r = ((r >> 1) & 0x55555555u) | ((r & 0x55555555u) << 1);
        r = ((r >> 2) & 0x33333333u) | ((r & 0x33333333u) << 2);
        r = ((r >> 4) & 0x0F0F0F0Fu) | ((r & 0x0F0F0F0Fu) << 4);
        r = __builtin_bswap32(r) >> (32 - n);
        put((uint64_t) r << (n - 1), 2 * n - 1);
    }
    inline void put_unsigned(uint32_t v) { put_code(v + 1); }
    inline void put_signed(int v) { put_code(v == 0 ? 1u : (((uint32_t) (v < 0 ? -v : v) << 1) | (v < 0 ? 1u : 0u))); }

    size_t finish()
    {
        const size_t bins = pos_ * 64 + n_;
        w_[pos_] = acc_;
        return bins;
    }
};

// stream.cpp:550-581 + serialize.cpp:10-23: one 8x8 block of a record
inline void put_block(bin_string &w, const int16_t *blk, const uint8_t *zz, int16_t last_dc)
{
    const int16_t dc = (int16_t) (blk[0] - last_dc);
    int run = 63;
    for (; run >= 1; --run) if (blk[zz[run]]) break;
    if (run == 0 && dc == 0) run = -1;
    run++;
    w.put_unsigned((uint32_t) run);
    if (run > 0)
    {
        w.put_signed(dc);
        for (int k = 1; k < run; ++k) w.put_signed(blk[zz[k]]);
    }
}

// Phase 2: the adaptive binary arithmetic coder over the bin string -- encode_symbol +
// resolve_encode_scaling (abac.cpp:97-121, 180-224) per bin, flush_encoder (abac.cpp:281-313)
// at the end.  The coder is one serial dependency chain through (low, high); everything here is
// about keeping that chain short and free of unpredictable branches (bin values are coin flips
// for a branch predictor, and one mispredict costs as much as the arithmetic of a whole bin):
//  * the model's split point floor((high-low) * h0 / tot) is ONE 64-bit multiply on the chain:
//    F = floor(h0 * 2^48 / tot) + (1..257) is prepared off the chain from the reciprocal table
//    (h0 and tot depend on the bins only), and ((high-low) * F) >> 48 equals the exact quotient
//    for tot < 2^23 (the excess of F adds less than 1/tot to a value whose fractional part is
//    at most 1 - 1/tot) -- profiles/abac_bench.cpp checks this identity exhaustively per tot;
//  * interval update and the first E3 step are selects written as masks;
//  * all leading bits on which low and high agree leave at once (E1/E2): with k such bits,
//    T = the top k bits of high and e3 pending bits, the bits that leave are, most significant
//    first, T + ((2^e3 - 1) << (k-1)) in k + e3 bits (first bit, e3 inverted copies, the
//    other k-1 bits).  They are shifted into a most-significant-first accumulator that is
//    stored every bin and advanced by whole bytes -- no "accumulator full" branch; one pass at
//    the end turns each byte around, because the stream is least-significant-bit first.
// Rare cases (collapsed interval, a long run of pending E3 bits, tot >= 2^23) take the
// bit-at-a-time path.  Returns the number of bits written to out.
// (built twice: a portable clone and one for BMI2/LZCNT machines, whose three-operand shifts
// and leading-zero count shave a fifth off the instruction count; picked at load time)
#if defined(__x86_64__) && defined(__GNUC__) && !defined(__clang__)
__attribute__((noinline, target_clones("default", "arch=haswell")))
#else
__attribute__((noinline))
#endif
uint64_t abac_encode_bins(const uint64_t *bins, size_t nbins, std::vector<uint8_t> &outv)
{
    uint64_t low = 0, high = AB_MAX, h0 = 1, tot = 2;
    uint64_t e3 = 0;
    uint64_t acc = 0;                                   // the low nacc bits are pending output, oldest on top
    uint64_t nacc = 0;                                  // < 8 between bins
    if (outv.size() < 4096) outv.resize(4096);
    uint8_t *wp = outv.data();                          // next output byte

    reciprocal_table &rt = recips();
    const size_t kFastTot = size_t(1) << 23;
    if (rt.size() < std::min(nbins + 4, kFastTot)) rt.grow(std::min(nbins + 4, kFastTot));
    const uint64_t *recip = rt.data();
    const uint64_t fast_limit = std::min<uint64_t>(rt.size(), kFastTot);

#define EVX_DRAIN() do { const uint64_t top_ = __builtin_bswap64(acc << ((64 - nacc) & 63)); memcpy(wp, &top_, 8); wp += nacc >> 3; nacc &= 7; } while (0)
#define EVX_PUT(b) do { acc = (acc << 1) | (uint64_t) (b); ++nacc; EVX_DRAIN(); } while (0)
#define EVX_GROW(need) do { const size_t pos_ = (size_t) (wp - outv.data()); if (pos_ + (need) > outv.size()) { outv.resize((pos_ + (need)) * 2); wp = outv.data() + pos_; } } while (0)
    // E1/E2 one bit at a time, as the reference does it (write_bit + flush_inverse_bits, abac.cpp:156-178)
#define EVX_RENORM_SLOW() \
    for (;;) \
    { \
        uint64_t b_; \
        if (high <= AB_HALF) b_ = 0; \
        else if (low > AB_HALF) { b_ = 1; low -= AB_HALF + 1; high -= AB_HALF + 1; } \
        else break; \
        EVX_GROW(64 + (e3 >> 3)); \
        EVX_PUT(b_); \
        for (; e3; --e3) EVX_PUT(b_ ^ 1u); \
        low = (low << 1) & AB_MAX; \
        high = ((high << 1) & AB_MAX) | 1u; \
    }
    // E3: low = 01..., high = 10... (with the reference's 3*QTR = 0xBFFD quirk); the MSBs differ here.
    // One step maps low -> 2*low - 0x8000, high -> 2*high - 0x7FFF.
#define EVX_E3() \
    do { \
        const uint64_t c_ = (uint64_t) (low > AB_QTR) & (uint64_t) (high <= AB_3QTR), cm_ = 0 - c_; \
        low += (low - 0x8000u) & cm_; \
        high += (high - 0x7FFFu) & cm_; \
        e3 += c_; \
        while (__builtin_expect(c_ && low > AB_QTR && high <= AB_3QTR, 0)) { low = 2 * low - 0x8000u; high = 2 * high - 0x7FFFu; e3++; } \
    } while (0)

    size_t base = 0;
    for (; base < nbins && tot + 64 <= fast_limit; base += 64)
    {
        uint64_t word = bins[base >> 6];
        const int cnt = nbins - base < 64 ? (int) (nbins - base) : 64;
        // a bin leaves at most 16 E1/E2 bits plus the E3 bits pending before it (one per earlier step)
        EVX_GROW((size_t) 64 * 4 + (e3 >> 3) + 64);
        for (int j = 0; j < cnt; ++j, word >>= 1)
        {
            const uint64_t bm = 0 - (word & 1u);
            const uint64_t F = (uint64_t) (((unsigned __int128) h0 * recip[tot]) >> 16) + 1;
            const uint64_t q = ((high - low) * F) >> 48;
            const uint64_t mid = low + q;
            low += (q + 1) & bm;                      // bit ? mid + 1 : low
            high = (high & bm) | (mid & ~bm);         // bit ? high : mid
            h0 += 1 + bm;                             // counts the zeros
            tot++;

            // k = number of leading bits low and high share (16 = collapsed interval)
            const uint64_t k = (uint64_t) __builtin_clz((uint32_t) (((low ^ high) << 16) + 0x8000u));
            if (__builtin_expect((k == 16) | (e3 > 40), 0)) EVX_RENORM_SLOW()
            else
            {
                const uint64_t nzm = 0 - ((k + 15) >> 4);                           // any bit leaving? (k <= 15 here)
                const uint64_t e3f = e3 & nzm;                                      // pending bits leave with the first one
                const uint64_t v = (high >> (16 - k)) + ((((uint64_t(1) << e3f) - 1) << k) >> 1);
                const uint64_t len = k + e3f;
                acc = (acc << len) | v;              // nacc <= 7, len <= 15 + 40
                nacc += len;
                e3 -= e3f;
                EVX_DRAIN();
                low = (low << k) & AB_MAX;
                high = ((high << k) & AB_MAX) | ((uint64_t(1) << k) - 1u);
            }
            EVX_E3();
        }
    }
    // beyond the reciprocal table (slices of more than 8M bins): the plain form
    for (; base < nbins; base += 64)
    {
        uint64_t word = bins[base >> 6];
        const int cnt = nbins - base < 64 ? (int) (nbins - base) : 64;
        for (int j = 0; j < cnt; ++j, word >>= 1)
        {
            const uint64_t bit = word & 1u;
            const uint64_t mid = low + ((high - low) * h0) / tot;
            if (bit) low = mid + 1; else { high = mid; h0++; }
            tot++;
            EVX_RENORM_SLOW()
            EVX_E3();
        }
    }
    // flush_encoder: one more pending bit, then the quarter the interval sits in
    EVX_GROW((e3 >> 3) + 64);
    e3++;
    {
        const uint64_t b = low < AB_QTR ? 0u : 1u;
        EVX_PUT(b);
        for (; e3; --e3) EVX_PUT(b ^ 1u);
    }
    uint8_t *out = outv.data();
    const uint64_t bits = (uint64_t) (wp - out) * 8 + nacc;
    { const uint64_t top_ = __builtin_bswap64(acc << ((64 - nacc) & 63)); memcpy(wp, &top_, 8); if (!nacc) *wp = 0; }
#undef EVX_DRAIN
#undef EVX_PUT
#undef EVX_GROW
#undef EVX_RENORM_SLOW
#undef EVX_E3
    // the stream is least-significant-bit first within a byte
    const size_t nbytes = (size_t) ((bits + 7) >> 3);
    for (size_t i = 0; i < nbytes; ++i) out[i] = REV8[out[i]];
    return bits;
}

}  // namespace

void slice_writer::configure(int mbw, int mbh, int ref_count)
{
    mbw_ = mbw; mbh_ = mbh;
    target_bits_ = bit_length((uint32_t) (ref_count & 0xFF)) - 1;     // log2((uint8) R), serialize.cpp:179
    dc_.resize((size_t) mbw * mbh);
    bins_.assign((size_t) mbw * mbh * 8 + 1024, 0);                   // grown on demand
    buf_.assign((size_t) mbw * mbh * 64 + 4096, 0);
}

void slice_writer::reset() { dc_.resize((size_t) mbw_ * mbh_); }

uint32_t slice_writer::serialize(const evxgpu_block_desc *t, const int16_t *records, uint32_t n_noncopy)
{
    const int n = mbw_ * mbh_;
    static const bool prof = getenv("EVX_ENTROPY_PROFILE") != NULL;
    auto now = [] { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count(); };
    const double tp0 = prof ? now() : 0;

    // refresh the DC mirror with this frame's non-copy macroblocks (serialisation reads the
    // coefficient planes AFTER the whole slice was encoded, encode.cpp:214-220)
    {
        uint32_t k = 0;
        for (int i = 0; i < n; ++i)
        {
            if (t[i].block_type & T_COPY) continue;
            const int16_t *r = records + (size_t) k * 384;
            dc_.y_tr[i] = r[8]; dc_.y_bl[i] = r[8 * 16]; dc_.u[i] = r[256]; dc_.v[i] = r[320];
            ++k;
        }
        if (k != n_noncopy) return 0;
    }

    bin_string w(bins_);
    w.reserve((size_t) n * 160);
    // block table by field, serialize.cpp:156-317
    for (int i = 0; i < n; ++i) w.put_bits_lsb((uint32_t) t[i].block_type, 3);
    for (int i = 0; i < n; ++i) if (!(t[i].block_type & T_INTRA)) w.put_bits_lsb(t[i].prediction_target, target_bits_);
    int last = 0;
    for (int i = 0; i < n; ++i) if (t[i].block_type & T_MOTION) { w.put_signed((int16_t) (t[i].motion_x - last)); last = t[i].motion_x; }
    last = 0;
    for (int i = 0; i < n; ++i) if (t[i].block_type & T_MOTION) { w.put_signed((int16_t) (t[i].motion_y - last)); last = t[i].motion_y; }
    for (int i = 0; i < n; ++i) if (t[i].block_type & T_MOTION) w.put(t[i].sp_pred & 1u, 1);
    for (int i = 0; i < n; ++i) if ((t[i].block_type & T_MOTION) && t[i].sp_pred) w.put(t[i].sp_amount & 1u, 1);
    for (int i = 0; i < n; ++i) if ((t[i].block_type & T_MOTION) && t[i].sp_pred) w.put_bits_lsb(t[i].sp_index, 3);
    last = 0;
    for (int i = 0; i < n; ++i) if (!(t[i].block_type & T_COPY)) { w.put_signed((int16_t) (t[i].q_index - last)); last = t[i].q_index; }

    // residuals: all luma, then all U, then all V (serialize.cpp:125-154)
    for (int comp = 0; comp < 3; ++comp)
    {
        uint32_t k = 0;
        int idx = 0;
        for (int by = 0; by < mbh_; ++by)
        for (int bx = 0; bx < mbw_; ++bx, ++idx)
        {
            if (t[idx].block_type & T_COPY) continue;
            const int16_t *r = records + (size_t) (k++) * 384;
            w.reserve(384 * 40);
            if (comp == 0)
            {
                int16_t last_dc = bx >= 1 ? dc_.y_tr[idx - 1] : (by >= 1 ? dc_.y_bl[idx - mbw_] : 0);
                put_block(w, r, ZZ.luma, last_dc);                     // serialize.cpp:25-34
                put_block(w, r + 8, ZZ.luma, r[0]);
                put_block(w, r + 8 * 16, ZZ.luma, r[0]);
                put_block(w, r + 8 * 16 + 8, ZZ.luma, r[8 * 16]);
            }
            else
            {
                const std::vector<int16_t> &m = comp == 1 ? dc_.u : dc_.v;
                int16_t last_dc = bx >= 1 ? m[idx - 1] : (by >= 1 ? m[idx - mbw_] : 0);
                put_block(w, r + 256 + (comp - 1) * 64, ZZ.pos, last_dc);
            }
        }
    }
    const size_t nbins = w.finish();
    const double tp1 = prof ? now() : 0;
    const uint32_t total_bits = (uint32_t) abac_encode_bins(bins_.data(), nbins, buf_);
    if (prof) fprintf(stderr, "[entropy] binarise %.3f ms (%zu bins), coder %.3f ms, %u bits\n", tp1 - tp0, nbins, now() - tp1, total_bits);
    return total_bits;
}

uint32_t slice_writer::serialize_bins(const uint64_t *bins, uint64_t nbins)
{
    return (uint32_t) abac_encode_bins(bins, (size_t) nbins, buf_);
}

// ---------------------------------------------------------------- decoder side

namespace {

// decode_symbol + resolve_decode_scaling (abac.cpp:123-154, 226-279).  Like the encoder, the
// decoder is one dependency chain through (low, high, value); the common path uses the same
// reciprocal split point (see abac_encode_bins), mask selects, and takes all E1/E2 shifts of a
// symbol at once: with k leading bits shared by low and high, the next k stream bits enter
// `value` together.  The stream is least-significant-bit first, so the slice is copied once with
// every byte turned around and the coder peeks 17 bits from a big-endian load.  Everything that
// is not the common case -- the last 17 bits of the slice (past its end the reference repeats
// the last bit read within the call), a collapsed interval, a value outside [low, high] (corrupt
// streams: the reference then updates nothing) -- goes bit at a time exactly as the reference does.
class abac_reader
{
    const uint8_t *data_;           // the caller's bits (LSB-first), absolute bit positions
    const uint8_t *rev_;            // the same bytes, each bit-reversed (MSB-first), padded
    uint32_t pos_, end_;
    uint32_t low_, high_, value_;
    uint64_t h0_, tot_;
    const uint64_t *recip_;
    uint64_t fast_limit_;

    inline bool empty() const { return pos_ >= end_; }
    inline uint32_t get() { uint32_t b = (data_[pos_ >> 3] >> (pos_ & 7)) & 1u; pos_++; return b; }

    // the reference's loop, one scaling step per turn
    inline void renorm_slow()
    {
        uint32_t bit = 0;
        for (;;)
        {
            if (high_ <= AB_HALF) { }
            else if (low_ > AB_HALF) { high_ -= AB_HALF + 1; low_ -= AB_HALF + 1; value_ -= AB_HALF + 1; }
            else if (high_ <= AB_3QTR && low_ > AB_QTR) { high_ -= AB_QTR + 1; low_ -= AB_QTR + 1; value_ -= AB_QTR + 1; }
            else break;
            if (!empty()) bit = get();
            high_ = ((high_ << 1) & AB_MAX) | 1u;
            low_ = (low_ << 1) & AB_MAX;
            value_ = ((value_ << 1) & AB_MAX) | bit;
        }
    }

public:
    abac_reader(const uint8_t *data, const uint8_t *rev, uint32_t pos, uint32_t end)
        : data_(data), rev_(rev), pos_(pos), end_(end), low_(0), high_(AB_MAX), value_(0), h0_(1), tot_(2)
    {   // start_decode, abac.cpp:398-420: past the end the LAST bit read is repeated
        uint32_t bit = 0;
        for (int i = 0; i < 16; ++i) { if (!empty()) bit = get(); value_ = (value_ << 1) | bit; }
        reciprocal_table &rt = recips();
        const size_t want = std::min<size_t>((size_t) (end - pos) * 2 + 64, size_t(1) << 23);     // a bin costs >= ~1/2 bit on real slices; grown below if not
        if (rt.size() < want) rt.grow(want);
        recip_ = rt.data();
        fast_limit_ = std::min<uint64_t>(rt.size(), uint64_t(1) << 23);
    }

    inline uint32_t decode()
    {
        const uint64_t range = high_ - low_;
        uint64_t q;
        if (__builtin_expect(tot_ < fast_limit_, 1)) q = (range * ((uint64_t) (((unsigned __int128) h0_ * recip_[tot_]) >> 16) + 1)) >> 48;
        else
        {
            q = (range * h0_) / tot_;
            if (tot_ < (uint64_t(1) << 23)) { recips().grow((size_t) tot_ * 2); recip_ = recips().data(); fast_limit_ = std::min<uint64_t>(recips().size(), uint64_t(1) << 23); }
        }
        const uint32_t mid = low_ + (uint32_t) q;
        const uint32_t in0 = (uint32_t) (value_ >= low_) & (uint32_t) (value_ <= mid);
        const uint32_t in1 = (uint32_t) (value_ > mid) & (uint32_t) (value_ <= high_);
        const uint32_t m0 = 0u - in0, m1 = 0u - in1;
        low_ += ((uint32_t) q + 1) & m1;              // bit 1: low = mid + 1
        high_ = (high_ & ~m0) | (mid & m0);           // bit 0: high = mid
        h0_ += in0;
        tot_ += in0 | in1;                            // neither (corrupt stream): nothing moves

        const uint32_t k = (uint32_t) __builtin_clz(((low_ ^ high_) << 16) + 0x8000u);           // shared leading bits; 16 = collapsed
        if (__builtin_expect((k == 16) | (pos_ + 17 > end_), 0)) { renorm_slow(); return in1; }
        // 17 stream bits from pos_, first one on top
        uint64_t w;
        memcpy(&w, rev_ + (pos_ >> 3), 8);
        w = (__builtin_bswap64(w) << (pos_ & 7)) >> 47;
        low_ = (low_ << k) & AB_MAX;
        high_ = ((high_ << k) & AB_MAX) | ((1u << k) - 1u);
        value_ = ((value_ << k) & AB_MAX) | (uint32_t) (w >> (17 - k));
        // E3 (3*QTR = 0xBFFD as in the reference): low -> 2*low - 0x8000, high -> 2*high - 0x7FFF, value -> 2*value - 0x8000 (mod 2^16) | next bit
        const uint32_t c = (uint32_t) (low_ > AB_QTR) & (uint32_t) (high_ <= AB_3QTR), cm = 0u - c;
        const uint32_t b1 = (uint32_t) (w >> (16 - k)) & 1u;
        low_ += (low_ - 0x8000u) & cm;
        high_ += (high_ - 0x7FFFu) & cm;
        value_ = (value_ & ~cm) | (((((value_ ^ 0x4000u) << 1) & AB_MAX) | b1) & cm);
        pos_ += k + c;
        if (__builtin_expect(c && low_ > AB_QTR && high_ <= AB_3QTR, 0)) renorm_slow();          // further E3 steps
        return in1;
    }

    inline uint32_t decode_bits_lsb(int n) { uint32_t v = 0; for (int k = 0; k < n; ++k) v |= decode() << k; return v; }

    // stream.cpp:292-436
    inline uint16_t decode_code(int *nbits)
    {
        int zeros = 0;
        uint32_t bit = decode();
        while (!bit && zeros < 48) { zeros++; bit = decode(); }
        uint16_t r = 0;
        for (int i = 0; i < zeros + 1; ++i) { r = (uint16_t) ((r << 1) | (bit & 1u)); if (i < zeros) bit = decode(); }
        *nbits = zeros + 1;
        return r;
    }
    inline uint16_t decode_unsigned() { int n; return (uint16_t) (decode_code(&n) - 1); }
    inline int16_t decode_signed()
    {
        int n;
        int16_t r = (int16_t) decode_code(&n);
        int16_t sign = (int16_t) (1 - 2 * (r & 1));
        r = (int16_t) (sign * ((r >> 1) & 0x7FFF));
        if (n + (n - 1) > 0x20) r = (int16_t) (r | 0x8000);
        return r;
    }
};

// stream.cpp:583-605 + unserialize.cpp:10-22
inline void get_block(abac_reader &rd, int16_t *blk, int stride, int16_t last_dc)
{
    for (int j = 0; j < 8; ++j) memset(blk + j * stride, 0, 16);
    uint16_t run = rd.decode_unsigned();
    for (uint32_t k = 0; k < run && k < 64; ++k)
    {
        int p = ZZ.pos[k];
        blk[(p >> 3) * stride + (p & 7)] = rd.decode_signed();
    }
    blk[0] = (int16_t) (blk[0] + last_dc);
}

}  // namespace

void slice_reader::configure(int mbw, int mbh, int ref_count)
{
    mbw_ = mbw; mbh_ = mbh;
    target_bits_ = bit_length((uint32_t) (ref_count & 0xFF)) - 1;
    dc_.resize((size_t) mbw * mbh);
}

void slice_reader::reset() { dc_.resize((size_t) mbw_ * mbh_); }

int slice_reader::parse(const uint8_t *data, uint32_t pos, uint32_t end, parsed_slice &out) const
{
    const int n = mbw_ * mbh_;
    {   // the slice's bytes, each turned around (see abac_reader); absolute byte positions, 16 bytes of padding
        const size_t first = pos >> 3, last = ((size_t) end + 7) >> 3;
        if (out.rev.size() < last + 16) out.rev.resize(last + 16);
        for (size_t i = first; i < last; ++i) out.rev[i] = REV8[data[i]];
        memset(out.rev.data() + last, 0, 16);
    }
    evxgpu_block_desc zero;
    memset(&zero, 0, sizeof(zero));
    out.fields.assign((size_t) n, zero);
    evxgpu_block_desc *t = out.fields.data();
    abac_reader rd(data, out.rev.data(), pos, end);
    for (int i = 0; i < n; ++i) t[i].block_type = (int32_t) rd.decode_bits_lsb(3);
    for (int i = 0; i < n; ++i)
        if (!(t[i].block_type & T_INTRA)) t[i].prediction_target = (uint8_t) rd.decode_bits_lsb(target_bits_);
    int16_t last = 0;
    for (int i = 0; i < n; ++i) if (t[i].block_type & T_MOTION) { t[i].motion_x = (int16_t) (last + rd.decode_signed()); last = t[i].motion_x; }
    last = 0;
    for (int i = 0; i < n; ++i) if (t[i].block_type & T_MOTION) { t[i].motion_y = (int16_t) (last + rd.decode_signed()); last = t[i].motion_y; }
    for (int i = 0; i < n; ++i) if (t[i].block_type & T_MOTION) t[i].sp_pred = (uint8_t) rd.decode();
    for (int i = 0; i < n; ++i) if ((t[i].block_type & T_MOTION) && t[i].sp_pred) t[i].sp_amount = (uint8_t) rd.decode();
    for (int i = 0; i < n; ++i) if ((t[i].block_type & T_MOTION) && t[i].sp_pred) t[i].sp_index = (uint8_t) rd.decode_bits_lsb(3);
    last = 0;
    for (int i = 0; i < n; ++i) if (!(t[i].block_type & T_COPY)) { t[i].q_index = (uint8_t) (rd.decode_signed() + last); last = t[i].q_index; }

    uint32_t count = 0;
    for (int i = 0; i < n; ++i) if (!(t[i].block_type & T_COPY)) count++;
    out.n_noncopy = count;
    if (out.records.size() < (size_t) count * 384) out.records.resize((size_t) count * 384);

    for (int comp = 0; comp < 3; ++comp)
    {
        uint32_t k = 0;
        for (int idx = 0; idx < n; ++idx)
        {
            if (t[idx].block_type & T_COPY) continue;
            int16_t *r = out.records.data() + (size_t) (k++) * 384;
            if (comp == 0)
            {   // unserialize.cpp:24-33; the first block's neighbour DC is added by apply()
                get_block(rd, r, 16, 0);
                get_block(rd, r + 8, 16, r[0]);
                get_block(rd, r + 8 * 16, 16, r[0]);
                get_block(rd, r + 8 * 16 + 8, 16, r[8 * 16]);
            }
            else get_block(rd, r + 256 + (comp - 1) * 64, 8, 0);
        }
    }
    return 0;
}

int slice_reader::apply(const parsed_slice &in, evxgpu_block_desc *t, int16_t *records, uint32_t *n_noncopy)
{
    const int n = mbw_ * mbh_;
    if (in.fields.size() != (size_t) n) return 1;
    const evxgpu_block_desc *f = in.fields.data();
    const uint8_t tmask = (uint8_t) ((1u << target_bits_) - 1u);
    for (int i = 0; i < n; ++i)
    {   // a field the frame does not carry keeps its old value (unserialize.cpp:123-319)
        t[i].block_type = (t[i].block_type & ~7) | f[i].block_type;
        if (!(f[i].block_type & T_INTRA)) t[i].prediction_target = (uint8_t) ((t[i].prediction_target & ~tmask) | f[i].prediction_target);
        if (f[i].block_type & T_MOTION)
        {
            t[i].motion_x = f[i].motion_x; t[i].motion_y = f[i].motion_y;
            t[i].sp_pred = (uint8_t) ((t[i].sp_pred & 0xFE) | f[i].sp_pred);
            if (f[i].sp_pred)
            {
                t[i].sp_amount = (uint8_t) ((t[i].sp_amount & 0xFE) | f[i].sp_amount);
                t[i].sp_index = (uint8_t) ((t[i].sp_index & ~7u) | f[i].sp_index);
            }
        }
        if (!(f[i].block_type & T_COPY)) t[i].q_index = f[i].q_index;
    }
    *n_noncopy = in.n_noncopy;
    if (records != in.records.data()) memcpy(records, in.records.data(), (size_t) in.n_noncopy * 384 * sizeof(int16_t));
    // DC prediction across macroblocks (unserialize.cpp:24-72): the neighbour's DC as of now -- from this frame if
    // it was coded, else what the mirror kept from the last frame that coded it.  The prediction is additive in
    // int16 arithmetic, so the neighbour's DC is added to every DC that was chained from it inside the macroblock.
    uint32_t k = 0;
    int idx = 0;
    for (int by = 0; by < mbh_; ++by)
    for (int bx = 0; bx < mbw_; ++bx, ++idx)
    {
        if (f[idx].block_type & T_COPY) continue;
        int16_t *r = records + (size_t) (k++) * 384;
        const int16_t ly = bx >= 1 ? dc_.y_tr[idx - 1] : (by >= 1 ? dc_.y_bl[idx - mbw_] : 0);
        r[0] = (int16_t) (r[0] + ly); r[8] = (int16_t) (r[8] + ly); r[8 * 16] = (int16_t) (r[8 * 16] + ly); r[8 * 16 + 8] = (int16_t) (r[8 * 16 + 8] + ly);
        dc_.y_tr[idx] = r[8]; dc_.y_bl[idx] = r[8 * 16];
        const int16_t lu = bx >= 1 ? dc_.u[idx - 1] : (by >= 1 ? dc_.u[idx - mbw_] : 0);
        r[256] = (int16_t) (r[256] + lu); dc_.u[idx] = r[256];
        const int16_t lv = bx >= 1 ? dc_.v[idx - 1] : (by >= 1 ? dc_.v[idx - mbw_] : 0);
        r[320] = (int16_t) (r[320] + lv); dc_.v[idx] = r[320];
    }
    return 0;
}

int slice_reader::unserialize(const uint8_t *data, uint32_t pos, uint32_t end, evxgpu_block_desc *t, int16_t *records, uint32_t *n_noncopy)
{
    int rc = parse(data, pos, end, tmp_);
    return rc ? rc : apply(tmp_, t, records, n_noncopy);
}

}  // namespace evx
