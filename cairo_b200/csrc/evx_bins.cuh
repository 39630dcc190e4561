// evx_bins.cuh -- K8: the slice as a string of bins, built on the device.
//
// serialize_slice (serialize.cpp:156-340) feeds the arithmetic coder field by field: all block
// types, all prediction targets, motion x deltas, motion y deltas, sub-pel flags, amounts and
// directions, quantiser deltas, then every luma block of every non-copy macroblock, every U
// block, every V block -- each syntax element binarised as raw bits or as an Exp-Golomb code
// (golomb.cpp:8-91, stream.cpp:550-581).  Only the coder itself is serial.  Binarisation is
// independent per element once each element knows (a) the value its delta is taken against
// and (b) where its bins start, so it runs here, and the host receives ~20 KB of bins per
// 1080p P-frame instead of 1.2 MB of coefficient records:
//
//   (evx_wavefront)    while it walks a row, K3 notes for every macroblock the previous one OF THAT
//                      ROW with a motion vector / with coefficients, and each row's last such block
//   evx_bins_lengths   one thread per ITEM (14 per macroblock: 8 table fields, 4 luma blocks,
//                      U, V; laid out field-major = stream order): its bin count; per-tile sums
//   evx_bins_emit      tile base = sum of the tiles before it, exclusive scan inside the tile,
//                      then every item ORs its bins into the zeroed string
// DC prediction (serialize.cpp:25-72) reads a neighbour's DC as of the end of this frame: from the
// neighbour's record if it was coded in this frame, else from the persistent DC mirror -- a copy
// block keeps the DC values of the last frame that coded it.  evx_bins_emit refreshes
// the mirror entries of the blocks coded in this frame (entries nobody reads during this frame).
//
// Bin i of the slice is bit (i & 31) of 32-bit word (i >> 5): the layout the host coder
// (entropy.cpp, abac_encode_bins) walks.
#pragma once

#include "evx_device.cuh"

#define EVX_BINS_ITEMS 14                 // items per macroblock
#define EVX_BINS_TILE 256                 // items per CTA of the lengths / emit kernels

__constant__ uint8_t EVX_ZIGZAG[64] = {   // scan.h:60-70 as row*8+col
    0, 1, 8, 16, 9, 2, 3, 10, 17, 24, 32, 25, 18, 11, 4, 5, 12, 19, 26, 33, 40, 48, 41, 34, 27, 20, 13, 6, 7, 14, 21, 28,
    35, 42, 49, 56, 57, 50, 43, 36, 29, 22, 15, 23, 30, 37, 44, 51, 58, 59, 52, 45, 38, 31, 39, 46, 53, 60, 61, 54, 47, 55, 62, 63 };

struct EvxBinsParams
{
    const EvxDesc *table;
    const int16_t *records;       // [nmb][384], slot = macroblock index (K3's layout)
    int16_t *dc;                  // persistent mirror: [4][nmb] = luma top-right DC, luma bottom-left DC, U DC, V DC
    const int *prev_motion, *prev_coded;   // [nmb] previous flagged macroblock of the same row, or -1 (written by K3)
    const int *row_last;          // [2][mbh] last flagged macroblock per row, or -1
    const int *row_records;       // [mbh] non-copy macroblocks per row
    uint32_t *len;                // [14*nmb]
    uint32_t *tile_sum;           // [ntiles]
    uint32_t *bins;               // the string, zeroed before evx_bins_emit
    uint32_t *total;              // [0] = number of bins, [1] = 1 if the string did not fit, [2] = non-copy macroblocks
    uint32_t cap_bits;
    int mbw, mbh, nmb, target_bits;
};

// ---------------------------------------------------------------- Exp-Golomb

__device__ __forceinline__ uint32_t evx_code_signed(int v) { return v == 0 ? 1u : (((uint32_t) (v < 0 ? -v : v) << 1) | (v < 0 ? 1u : 0u)); }
__device__ __forceinline__ int evx_code_len(uint32_t x) { return 2 * (32 - __clz(x)) - 1; }
// the code of x as bins, first bin in bit 0: n-1 zeros, then the n bits of x most significant first
__device__ __forceinline__ uint64_t evx_code_bins(uint32_t x)
{
    const int n = 32 - __clz(x);
    return (uint64_t) (__brev(x) >> (32 - n)) << (n - 1);
}

// ---------------------------------------------------------------- bin sink: one item's bins, ORed into the string

struct EvxBinSink
{
    uint32_t *buf;
    uint64_t acc;
    uint32_t word;
    int fill;
    __device__ __forceinline__ void open(uint32_t *b, uint32_t bitpos) { buf = b; acc = 0; word = bitpos >> 5; fill = (int) (bitpos & 31); }
    __device__ __forceinline__ void put(uint64_t v, int n)       // n <= 33
    {
        acc |= v << fill;
        fill += n;
        if (fill >= 32)
        {
            atomicOr(buf + word, (uint32_t) acc);
            ++word; acc >>= 32; fill -= 32;
            if (fill >= 32) { atomicOr(buf + word, (uint32_t) acc); ++word; acc >>= 32; fill -= 32; }
        }
    }
    __device__ __forceinline__ void close() { if (fill > 0 && (uint32_t) acc) atomicOr(buf + word, (uint32_t) acc); }
};

struct EvxBinCount
{
    uint32_t n;
    __device__ __forceinline__ void put(uint64_t, int k) { n += (uint32_t) k; }
};

// ---------------------------------------------------------------- one item

// item -> (kind, macroblock, block).  Layout = stream order: 8 table fields x nmb, then the
// luma blocks macroblock-major, then U, then V.
__device__ __forceinline__ void evx_bins_item_decode(uint32_t item, int nmb, int &field, int &mb, int &blk)
{
    if (item < 8u * nmb) { field = (int) (item / nmb); mb = (int) (item % nmb); blk = 0; }
    else if (item < 12u * nmb) { field = 8; mb = (int) ((item - 8u * nmb) >> 2); blk = (int) ((item - 8u * nmb) & 3u); }
    else { field = item < 13u * nmb ? 9 : 10; mb = (int) (item - (field == 9 ? 12u : 13u) * nmb); blk = 0; }
}

// previous flagged macroblock in raster order: in the row, else the last one of the nearest row above that has any
__device__ __forceinline__ int evx_bins_prev(const int *prev_in_row, const int *row_last, int mb, int mbw)
{
    const int p = prev_in_row[mb];
    if (p >= 0) return p;
    for (int r = mb / mbw - 1; r >= 0; --r) { const int l = row_last[r]; if (l >= 0) return l; }
    return -1;
}

// DC number k (0 luma top-right, 1 luma bottom-left, 2 U, 3 V) of macroblock nb as of the end of this frame
__device__ __forceinline__ int evx_bins_dc(const EvxBinsParams &p, int nb, int k)
{
    if (p.table[nb].type() & EVX_T_COPY) return p.dc[(size_t) k * p.nmb + nb];
    const int16_t *r = p.records + (size_t) nb * 384;
    return r[k == 0 ? 8 : k == 1 ? 8 * 16 : k == 2 ? 256 : 320];
}

// stream.cpp:550-581 + serialize.cpp:10-23: one 8x8 block
template <class Sink>
__device__ __forceinline__ void evx_bins_block(Sink &s, const int16_t *blk, int stride, int last_dc)
{
    const int dc = (int) (int16_t) (blk[0] - last_dc);
    int run = 63;
    for (; run >= 1; --run)
    {
        const int z = EVX_ZIGZAG[run];
        if (blk[(z >> 3) * stride + (z & 7)]) break;
    }
    if (run == 0 && dc == 0) run = -1;
    run++;
    { const uint32_t x = (uint32_t) run + 1u; s.put(evx_code_bins(x), evx_code_len(x)); }
    if (run > 0)
    {
        { const uint32_t x = evx_code_signed(dc); s.put(evx_code_bins(x), evx_code_len(x)); }
        for (int k = 1; k < run; ++k)
        {
            const int z = EVX_ZIGZAG[k];
            const uint32_t x = evx_code_signed(blk[(z >> 3) * stride + (z & 7)]);
            s.put(evx_code_bins(x), evx_code_len(x));
        }
    }
}

template <class Sink>
__device__ __forceinline__ void evx_bins_item(Sink &s, const EvxBinsParams &p, uint32_t item)
{
    int field, mb, b;
    evx_bins_item_decode(item, p.nmb, field, mb, b);
    const EvxDesc d = p.table[mb];
    const int type = d.type();
    const bool motion = (type & EVX_T_MOTION) != 0, coded = !(type & EVX_T_COPY);
    switch (field)
    {
    case 0: s.put((uint64_t) (type & 7), 3); break;                                              // serialize.cpp:156-170
    case 1: if (!(type & EVX_T_INTRA) && p.target_bits) s.put((uint64_t) (d.target() & ((1 << p.target_bits) - 1)), p.target_bits); break;
    case 2: case 3:
        if (motion)
        {
            const int pm = evx_bins_prev(p.prev_motion, p.row_last, mb, p.mbw);
            int last = 0;
            if (pm >= 0) { const EvxDesc q = p.table[pm]; last = field == 2 ? q.mx() : q.my(); }
            const uint32_t x = evx_code_signed((int) (int16_t) ((field == 2 ? d.mx() : d.my()) - last));
            s.put(evx_code_bins(x), evx_code_len(x));
        }
        break;
    case 4: if (motion) s.put((uint64_t) (d.sp_pred() & 1), 1); break;
    case 5: if (motion && d.sp_pred()) s.put((uint64_t) (d.sp_amount() & 1), 1); break;
    case 6: if (motion && d.sp_pred()) s.put((uint64_t) (d.sp_index() & 7), 3); break;
    case 7:
        if (coded)
        {
            const int pc = evx_bins_prev(p.prev_coded, p.row_last + p.mbh, mb, p.mbw);
            const int last = pc >= 0 ? p.table[pc].q_index() : 0;
            const uint32_t x = evx_code_signed((int) (int16_t) (d.q_index() - last));
            s.put(evx_code_bins(x), evx_code_len(x));
        }
        break;
    default:
        if (coded)
        {
            const int16_t *r = p.records + (size_t) mb * 384;
            const int bx = mb % p.mbw, by = mb / p.mbw;
            if (field == 8)
            {   // serialize.cpp:25-34: the four luma blocks and what each one's DC is predicted from
                int last_dc;
                if (b == 0) last_dc = bx >= 1 ? evx_bins_dc(p, mb - 1, 0) : (by >= 1 ? evx_bins_dc(p, mb - p.mbw, 1) : 0);
                else if (b == 3) last_dc = r[8 * 16];
                else last_dc = r[0];
                evx_bins_block(s, r + (b >> 1) * 8 * 16 + (b & 1) * 8, 16, last_dc);
            }
            else
            {
                const int k = field == 9 ? 2 : 3;
                const int last_dc = bx >= 1 ? evx_bins_dc(p, mb - 1, k) : (by >= 1 ? evx_bins_dc(p, mb - p.mbw, k) : 0);
                evx_bins_block(s, r + 256 + (field - 9) * 64, 8, last_dc);
            }
        }
        break;
    }
}

// ---------------------------------------------------------------- kernels

__global__ void __launch_bounds__(EVX_BINS_TILE) evx_bins_lengths(EvxBinsParams p)
{
    __shared__ uint32_t s_w[EVX_BINS_TILE / 32];
    const uint32_t item = blockIdx.x * EVX_BINS_TILE + threadIdx.x, nitems = (uint32_t) EVX_BINS_ITEMS * p.nmb;
    EvxBinCount c; c.n = 0;
    if (item < nitems) { evx_bins_item(c, p, item); p.len[item] = c.n; }
    const uint32_t ws = __reduce_add_sync(0xFFFFFFFFu, c.n);
    if ((threadIdx.x & 31) == 0) s_w[threadIdx.x >> 5] = ws;
    __syncthreads();
    if (threadIdx.x == 0)
    {
        uint32_t t = 0;
        for (int w = 0; w < EVX_BINS_TILE / 32; ++w) t += s_w[w];
        p.tile_sum[blockIdx.x] = t;
    }
}

__global__ void __launch_bounds__(EVX_BINS_TILE) evx_bins_emit(EvxBinsParams p)
{
    __shared__ uint32_t s_w[EVX_BINS_TILE / 32];
    __shared__ uint32_t s_base;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const uint32_t item = blockIdx.x * EVX_BINS_TILE + tid, nitems = (uint32_t) EVX_BINS_ITEMS * p.nmb;
    // bins in the tiles before this one
    uint32_t part = 0;
    for (uint32_t t = tid; t < blockIdx.x; t += EVX_BINS_TILE) part += p.tile_sum[t];
    part = __reduce_add_sync(0xFFFFFFFFu, part);
    if (lane == 0) s_w[warp] = part;
    __syncthreads();
    if (tid == 0)
    {
        uint32_t t = 0;
        for (int w = 0; w < EVX_BINS_TILE / 32; ++w) t += s_w[w];
        s_base = t;
    }
    __syncthreads();
    const uint32_t base = s_base;
    __syncthreads();
    // exclusive scan of the tile's lengths
    const uint32_t n = item < nitems ? p.len[item] : 0u;
    uint32_t inc = n;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const uint32_t a = __shfl_up_sync(0xFFFFFFFFu, inc, o); if (lane >= o) inc += a; }
    if (lane == 31) s_w[warp] = inc;
    __syncthreads();
    uint32_t off = base + inc - n;
    for (int w = 0; w < warp; ++w) off += s_w[w];
    if (blockIdx.x == gridDim.x - 1 && tid == EVX_BINS_TILE - 1)
    {
        const uint32_t total = off + n;
        p.total[0] = total;
        p.total[1] = total > p.cap_bits ? 1u : 0u;
        uint32_t coded = 0;
        for (int r = 0; r < p.mbh; ++r) coded += (uint32_t) p.row_records[r];
        p.total[2] = coded;
    }
    if (item >= 7u * p.nmb && item < 8u * p.nmb)
    {   // the quantiser-delta item exists once per macroblock: refresh the DC mirror of the coded ones
        const int mb = (int) (item - 7u * p.nmb);
        if (!(p.table[mb].type() & EVX_T_COPY))
        {
            const int16_t *r = p.records + (size_t) mb * 384;
            p.dc[0 * p.nmb + mb] = r[8]; p.dc[1 * p.nmb + mb] = r[8 * 16]; p.dc[2 * p.nmb + mb] = r[256]; p.dc[3 * p.nmb + mb] = r[320];
        }
    }
    if (n == 0 || off + n > p.cap_bits) return;      // an overflowing string is re-emitted into a larger buffer by the host side
    EvxBinSink s;
    s.open(p.bins, off);
    evx_bins_item(s, p, item);
    s.close();
}
