// entropy.h -- host entropy stage of the EVX-1 stream: the slice layout of
// serialize.cpp:156-340 / unserialize.cpp:123-341 (block table by field, then the Y, U and
// V residual planes), Exp-Golomb binarisation (golomb.cpp:8-91) and the single-context
// adaptive binary arithmetic coder (abac.cpp).  Written from scratch around a word-wide bit
// writer instead of the reference's bit-at-a-time feed stream; the produced bits are identical.
#ifndef CAIRO_B200_ENTROPY_H
#define CAIRO_B200_ENTROPY_H

#include <stdint.h>

#include <vector>

#include "evxgpu.h"

namespace evx {

// What the reference keeps implicitly in its persistent coefficient planes: the four DC
// values per macroblock that a NEIGHBOUR's DC prediction can read, whatever frame they were
// last written in (copy blocks do not refresh them -- serialize.cpp:59-72).
struct dc_mirror
{
    std::vector<int16_t> y_tr, y_bl, u, v;      // luma top-right / bottom-left 8x8 DC, chroma DCs
    void resize(size_t n) { y_tr.assign(n, 0); y_bl.assign(n, 0); u.assign(n, 0); v.assign(n, 0); }
};

class slice_writer
{
    int mbw_, mbh_, target_bits_;
    dc_mirror dc_;
    std::vector<uint64_t> bins_;    // the slice as a string of bins (phase 1)
    std::vector<uint8_t> buf_;      // the coded bits (phase 2)

public:
    void configure(int mbw, int mbh, int ref_count);
    void reset();
    // table: mbw*mbh descriptors; records: the non-copy macroblocks' 384 coefficients in raster
    // order.  Returns the slice's bit count; the bits are in data() (LSB-first).
    uint32_t serialize(const evxgpu_block_desc *table, const int16_t *records, uint32_t n_noncopy);
    // The bin string built on the device (evxgpu_encode_collect_bins): the coder only.
    uint32_t serialize_bins(const uint64_t *bins, uint64_t nbins);
    const uint8_t *data() const { return buf_.data(); }
};

// One slice, entropy-decoded but not yet merged into the stream's state.  Everything the arithmetic decoder
// and the syntax need is inside the slice itself (the coder is reset per frame, serialize.cpp:323; motion and
// quantiser deltas chain within the frame), so slices of consecutive frames can be parsed concurrently; what
// depends on earlier frames -- table fields a frame does not carry, and the DC prediction from a neighbour
// that may be a stale copy block -- is resolved by slice_reader::apply, in frame order.
struct parsed_slice
{
    std::vector<evxgpu_block_desc> fields;   // per macroblock, only the fields this frame carries
    std::vector<int16_t> records;            // non-copy macroblocks, raster order; DCs relative to a zero neighbour
    uint32_t n_noncopy;
    std::vector<uint8_t> rev;                // scratch: the slice's bytes bit-reversed (see abac_reader)
};

class slice_reader
{
    int mbw_, mbh_, target_bits_;
    dc_mirror dc_;
    parsed_slice tmp_;

public:
    void configure(int mbw, int mbh, int ref_count);
    void reset();
    // Decodes one slice from bits [pos, end) of data.  `table` is persistent across frames
    // (fields a frame does not carry keep their old values, as in the reference); `records`
    // receives the non-copy macroblocks' coefficients in raster order.  = parse + apply.
    int unserialize(const uint8_t *data, uint32_t pos, uint32_t end, evxgpu_block_desc *table, int16_t *records, uint32_t *n_noncopy);
    // The two halves.  parse touches no state of the reader (thread-safe, any order); apply must run in frame order.
    int parse(const uint8_t *data, uint32_t pos, uint32_t end, parsed_slice &out) const;
    int apply(const parsed_slice &in, evxgpu_block_desc *table, int16_t *records, uint32_t *n_noncopy);
};

}  // namespace evx

#endif
