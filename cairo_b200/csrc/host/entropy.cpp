// entropy.cpp -- see entropy.h.
#include "entropy.h"

#include <string.h>

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
    static const size_t kMax = size_t(1) << 24;
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

class abac_writer
{
    uint32_t low_, high_, e3_, h0_, tot_;
    uint64_t acc_;
    uint32_t nacc_;
    uint8_t *out_;
    size_t pos_;
    const uint64_t *recip_;
    size_t recip_size_;

    inline void flush_acc() { memcpy(out_ + pos_, &acc_, 8); pos_ += 8; acc_ = 0; nacc_ = 0; }
    inline void put(uint32_t bit)
    {
        acc_ |= (uint64_t) bit << nacc_;
        if (++nacc_ == 64) flush_acc();
    }
    // n <= 16 bits, value's bit (n-1) goes out first (the stream is LSB-first, so reverse them)
    inline void put_msb_first(uint32_t value, uint32_t n)
    {
        for (uint32_t i = n; i-- > 0;) put((value >> i) & 1u);
    }
    inline void emit(uint32_t bit)        // write_bit + flush_inverse_bits, abac.cpp:156-178
    {
        put(bit);
        for (; e3_; --e3_) put(bit ^ 1u);
    }

public:
    explicit abac_writer(uint8_t *out) : low_(0), high_(AB_MAX), e3_(0), h0_(1), tot_(2), acc_(0), nacc_(0), out_(out), pos_(0)
    {
        recip_ = recips().data();
        recip_size_ = recips().size();
    }

    size_t bytes_pending() const { return pos_ + 8; }
    void rebase(uint8_t *out) { out_ = out; }

    // encode_symbol + resolve_encode_scaling, abac.cpp:97-121, 180-224.
    // The bin values are close to coin flips for the branch predictor, so the common path is
    // written without data-dependent branches: the interval update is a pair of selects, the
    // E1/E2 bits that leave (all leading bits on which low and high agree) are assembled into one
    // word -- first bit, then the pending E3 bits inverted, then the remaining k-1 bits -- and
    // OR-ed into the accumulator, and the first E3 step is a select as well.
    inline void encode(uint32_t bit)
    {
        const uint64_t a = (uint64_t) (high_ - low_) * h0_;
        uint32_t q;
        if (__builtin_expect(tot_ < recip_size_, 1)) q = (uint32_t) (((unsigned __int128) a * recip_[tot_]) >> 64);
        else
        {
            if (tot_ < (size_t(1) << 24)) { recips().grow((size_t) tot_ * 2); recip_size_ = recips().size(); }
            q = (uint32_t) (a / tot_);
        }
        const uint32_t mid = low_ + q;
        low_ = bit ? mid + 1 : low_;
        high_ = bit ? high_ : mid;
        h0_ += bit ^ 1u;
        tot_++;

        uint32_t k = (uint32_t) __builtin_clz((high_ ^ low_) | 1u) - 16;      // 0..15 common leading bits
        if (__builtin_expect(e3_ > 40 || nacc_ > 64 - 57 + 0 && nacc_ + k + e3_ > 64 || (high_ == low_), 0))
        {   // rare: long E3 run, accumulator about to wrap, or a collapsed interval -- bit at a time
            for (;;)
            {
                k = (uint32_t) __builtin_clz((high_ ^ low_) | 1u) - 16;
                if (!k) break;
                emit(high_ >> 15);
                if (k > 1) put_msb_first((high_ >> (16 - k)) & ((1u << (k - 1)) - 1u), k - 1);
                low_ = (low_ << k) & AB_MAX;
                high_ = ((high_ << k) & AB_MAX) | ((1u << k) - 1u);
            }
        }
        else
        {
            const uint32_t msb = high_ >> 15;
            const uint32_t km1 = k - (k != 0);
            const uint32_t rest = (high_ >> (16 - k)) & ((1u << km1) - 1u);                     // k-1 bits, first-out at the top
            const uint32_t rest_rev = (uint32_t) ((REV8[rest & 0xFF] << 8) | REV8[rest >> 8]) >> (16 - km1);
            const uint64_t pend = msb ? 0 : ((uint64_t(1) << e3_) - 1);                          // e3 copies of !msb
            uint64_t v = (uint64_t) msb | (pend << 1) | ((uint64_t) rest_rev << (1 + e3_));
            const uint32_t len = k ? k + e3_ : 0;
            v = k ? v : 0;
            acc_ |= v << nacc_;
            nacc_ += len;
            e3_ = k ? 0 : e3_;
            if (nacc_ >= 64)
            {   // exactly full (the guard above keeps nacc_ + len <= 64)
                memcpy(out_ + pos_, &acc_, 8); pos_ += 8; acc_ = 0; nacc_ = 0;
            }
            low_ = (low_ << k) & AB_MAX;
            high_ = ((high_ << k) & AB_MAX) | ((1u << k) - 1u);
        }
        // E3: low = 01..., high = 10... (with the reference's 3*QTR = 0xBFFD quirk); MSBs differ here
        uint32_t c = (low_ > AB_QTR) & (high_ <= AB_3QTR);
        low_ = c ? ((low_ - (AB_QTR + 1)) << 1) & AB_MAX : low_;
        high_ = c ? ((((high_ - (AB_QTR + 1)) << 1) & AB_MAX) | 1u) : high_;
        e3_ += c;
        while (__builtin_expect(c && low_ > AB_QTR && high_ <= AB_3QTR, 0))
        {
            low_ = ((low_ - (AB_QTR + 1)) << 1) & AB_MAX;
            high_ = (((high_ - (AB_QTR + 1)) << 1) & AB_MAX) | 1u;
            e3_++;
        }
    }

    inline void encode_bits_lsb(uint32_t v, int n) { for (int k = 0; k < n; ++k) encode((v >> k) & 1u); }

    // Exp-Golomb (golomb.cpp:8-91): n-1 zeros, then x MSB-first
    inline void encode_code(uint32_t x)
    {
        const int n = bit_length(x);
        for (int i = 0; i < n - 1; ++i) encode(0);
        for (int i = n - 1; i >= 0; --i) encode((x >> i) & 1u);
    }
    inline void encode_unsigned(uint32_t v) { encode_code(v + 1); }
    inline void encode_signed(int v) { encode_code(v == 0 ? 1u : (((uint32_t) (v < 0 ? -v : v) << 1) | (v < 0 ? 1u : 0u))); }

    // flush_encoder, abac.cpp:281-313.  Returns the total number of bits written.
    uint64_t finish()
    {
        e3_++;
        emit(low_ < AB_QTR ? 0u : 1u);
        uint64_t bits = (uint64_t) pos_ * 8 + nacc_;
        memcpy(out_ + pos_, &acc_, 8);
        return bits;
    }
};

// stream.cpp:550-581 + serialize.cpp:10-23: one 8x8 block of a record
inline void put_block(abac_writer &w, const int16_t *blk, const uint8_t *zz, int16_t last_dc)
{
    const int16_t dc = (int16_t) (blk[0] - last_dc);
    int run = 63;
    for (; run >= 1; --run) if (blk[zz[run]]) break;
    if (run == 0 && dc == 0) run = -1;
    run++;
    w.encode_unsigned((uint32_t) run);
    if (run > 0)
    {
        w.encode_signed(dc);
        for (int k = 1; k < run; ++k) w.encode_signed(blk[zz[k]]);
    }
}

}  // namespace

void slice_writer::configure(int mbw, int mbh, int ref_count)
{
    mbw_ = mbw; mbh_ = mbh;
    target_bits_ = bit_length((uint32_t) (ref_count & 0xFF)) - 1;     // log2((uint8) R), serialize.cpp:179
    dc_.resize((size_t) mbw * mbh);
    // worst case: every coefficient an escape-length code; grown on demand below
    buf_.assign((size_t) mbw * mbh * 384 * 5 + 4096, 0);
}

void slice_writer::reset() { dc_.resize((size_t) mbw_ * mbh_); }

uint32_t slice_writer::serialize(const evxgpu_block_desc *t, const int16_t *records, uint32_t n_noncopy)
{
    const int n = mbw_ * mbh_;
    abac_writer w(buf_.data());
    static const bool prof = getenv("EVX_ENTROPY_PROFILE") != NULL;
    auto now = [] { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count(); };
    double tp0 = prof ? now() : 0, tp1 = 0, tp2 = 0;

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

    if (prof) tp1 = now();
    // block table by field, serialize.cpp:156-317
    for (int i = 0; i < n; ++i) w.encode_bits_lsb((uint32_t) t[i].block_type, 3);
    for (int i = 0; i < n; ++i) if (!(t[i].block_type & T_INTRA)) w.encode_bits_lsb(t[i].prediction_target, target_bits_);
    int last = 0;
    for (int i = 0; i < n; ++i) if (t[i].block_type & T_MOTION) { w.encode_signed((int16_t) (t[i].motion_x - last)); last = t[i].motion_x; }
    last = 0;
    for (int i = 0; i < n; ++i) if (t[i].block_type & T_MOTION) { w.encode_signed((int16_t) (t[i].motion_y - last)); last = t[i].motion_y; }
    for (int i = 0; i < n; ++i) if (t[i].block_type & T_MOTION) w.encode(t[i].sp_pred & 1u);
    for (int i = 0; i < n; ++i) if ((t[i].block_type & T_MOTION) && t[i].sp_pred) w.encode(t[i].sp_amount & 1u);
    for (int i = 0; i < n; ++i) if ((t[i].block_type & T_MOTION) && t[i].sp_pred) w.encode_bits_lsb(t[i].sp_index, 3);
    last = 0;
    for (int i = 0; i < n; ++i) if (!(t[i].block_type & T_COPY)) { w.encode_signed((int16_t) (t[i].q_index - last)); last = t[i].q_index; }

    if (prof) tp2 = now();
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
            if (w.bytes_pending() + 8192 > buf_.size()) { buf_.resize(buf_.size() * 2); w.rebase(buf_.data()); }
        }
    }
    uint32_t total_bits = (uint32_t) w.finish();
    if (prof) fprintf(stderr, "[entropy] mirror %.3f ms, table %.3f ms, residuals %.3f ms, %u bits\n", tp1 - tp0, tp2 - tp1, now() - tp2, total_bits);
    return total_bits;
}

// ---------------------------------------------------------------- decoder side

namespace {

class abac_reader
{
    const uint8_t *data_;
    uint32_t pos_, end_;
    uint32_t low_, high_, value_, h0_, h1_;

    inline bool empty() const { return pos_ >= end_; }
    inline uint32_t get() { uint32_t b = (data_[pos_ >> 3] >> (pos_ & 7)) & 1u; pos_++; return b; }

public:
    abac_reader(const uint8_t *data, uint32_t pos, uint32_t end) : data_(data), pos_(pos), end_(end), low_(0), high_(AB_MAX), value_(0), h0_(1), h1_(1)
    {   // start_decode, abac.cpp:398-420: past the end the LAST bit read is repeated
        uint32_t bit = 0;
        for (int i = 0; i < 16; ++i) { if (!empty()) bit = get(); value_ = (value_ << 1) | bit; }
    }

    // decode_symbol + resolve_decode_scaling, abac.cpp:123-154, 226-279
    inline uint32_t decode()
    {
        const uint32_t range = high_ - low_;
        const uint32_t mid = low_ + (uint32_t) (((uint64_t) range * h0_) / (h0_ + h1_));
        uint32_t out = 0;
        if (value_ >= low_ && value_ <= mid) { high_ = mid; h0_++; }
        else if (value_ > mid && value_ <= high_) { low_ = mid + 1; h1_++; out = 1; }
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
        return out;
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

int slice_reader::unserialize(const uint8_t *data, uint32_t pos, uint32_t end, evxgpu_block_desc *t, int16_t *records, uint32_t *n_noncopy)
{
    const int n = mbw_ * mbh_;
    abac_reader rd(data, pos, end);
    for (int i = 0; i < n; ++i) t[i].block_type = (t[i].block_type & ~7) | (int32_t) rd.decode_bits_lsb(3);
    for (int i = 0; i < n; ++i)
        if (!(t[i].block_type & T_INTRA))
            t[i].prediction_target = (uint8_t) ((t[i].prediction_target & ~((1u << target_bits_) - 1u)) | rd.decode_bits_lsb(target_bits_));
    int16_t last = 0;
    for (int i = 0; i < n; ++i) if (t[i].block_type & T_MOTION) { t[i].motion_x = (int16_t) (last + rd.decode_signed()); last = t[i].motion_x; }
    last = 0;
    for (int i = 0; i < n; ++i) if (t[i].block_type & T_MOTION) { t[i].motion_y = (int16_t) (last + rd.decode_signed()); last = t[i].motion_y; }
    for (int i = 0; i < n; ++i) if (t[i].block_type & T_MOTION) t[i].sp_pred = (uint8_t) ((t[i].sp_pred & 0xFE) | rd.decode());
    for (int i = 0; i < n; ++i) if ((t[i].block_type & T_MOTION) && t[i].sp_pred) t[i].sp_amount = (uint8_t) ((t[i].sp_amount & 0xFE) | rd.decode());
    for (int i = 0; i < n; ++i) if ((t[i].block_type & T_MOTION) && t[i].sp_pred) t[i].sp_index = (uint8_t) ((t[i].sp_index & ~7u) | rd.decode_bits_lsb(3));
    last = 0;
    for (int i = 0; i < n; ++i) if (!(t[i].block_type & T_COPY)) { t[i].q_index = (uint8_t) (rd.decode_signed() + last); last = t[i].q_index; }

    uint32_t count = 0;
    for (int i = 0; i < n; ++i) if (!(t[i].block_type & T_COPY)) count++;
    *n_noncopy = count;

    for (int comp = 0; comp < 3; ++comp)
    {
        uint32_t k = 0;
        int idx = 0;
        for (int by = 0; by < mbh_; ++by)
        for (int bx = 0; bx < mbw_; ++bx, ++idx)
        {
            if (t[idx].block_type & T_COPY) continue;
            int16_t *r = records + (size_t) (k++) * 384;
            if (comp == 0)
            {
                int16_t last_dc = bx >= 1 ? dc_.y_tr[idx - 1] : (by >= 1 ? dc_.y_bl[idx - mbw_] : 0);
                get_block(rd, r, 16, last_dc);                         // unserialize.cpp:24-33
                get_block(rd, r + 8, 16, r[0]);
                get_block(rd, r + 8 * 16, 16, r[0]);
                get_block(rd, r + 8 * 16 + 8, 16, r[8 * 16]);
                dc_.y_tr[idx] = r[8]; dc_.y_bl[idx] = r[8 * 16];
            }
            else
            {
                std::vector<int16_t> &m = comp == 1 ? dc_.u : dc_.v;
                int16_t last_dc = bx >= 1 ? m[idx - 1] : (by >= 1 ? m[idx - mbw_] : 0);
                int16_t *b = r + 256 + (comp - 1) * 64;
                get_block(rd, b, 8, last_dc);
                m[idx] = b[0];
            }
        }
    }
    return 0;
}

}  // namespace evx
