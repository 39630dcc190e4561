// evx1.h -- public API of the B200 build, source-compatible with the reference's evx1.h
// (evx1.h:66-122) and bitstream.h (bitstream.h:43-92): same namespace, class names, method
// names, argument meaning and status codes, so a caller of the reference recompiles against
// this header unchanged.  The pixel pipeline behind encode()/decode() runs on the GPU
// (include/evxgpu.h); only the serial entropy stage runs on the host.
//
// Additions (not in the reference): evx1_config + create_encoder_ex/create_decoder_ex make the
// reference's compile-time switches (config.h:38-53) run-time values, bit_stream gains
// query_read_index()/query_write_index(), the encoder exposes last_frame_stats(), and encode() is
// also available as its two halves, submit() + collect(), so that the host entropy stage of frame n
// overlaps the device's work on frame n+1 (one frame of latency, the same bytes).
#ifndef CAIRO_B200_EVX1_H
#define CAIRO_B200_EVX1_H

#include <stdint.h>
#include <stddef.h>

namespace evx {

typedef int64_t int64;
typedef int32_t int32;
typedef int16_t int16;
typedef int8_t int8;
typedef uint64_t uint64;
typedef uint32_t uint32;
typedef uint16_t uint16;
typedef uint8_t uint8;

typedef uint8 evx_status;          // base.h:150

}  // namespace evx

// status codes, base.h:152-169
#define EVX_SUCCESS                    (0)
#define EVX_ERROR_INVALIDARG           (1)
#define EVX_ERROR_NOTIMPL              (2)
#define EVX_ERROR_OUTOFMEMORY          (3)
#define EVX_ERROR_UNDEFINED            (4)
#define EVX_ERROR_HARDWAREFAIL         (5)
#define EVX_ERROR_INVALID_INDEX        (6)
#define EVX_ERROR_CAPACITY_LIMIT       (7)
#define EVX_ERROR_INVALID_RESOURCE     (8)
#define EVX_ERROR_OPERATION_TIMEDOUT   (9)
#define EVX_ERROR_EXECUTION_FAILURE    (10)
#define EVX_ERROR_NOT_READY            (15)

#define evx_succeeded(status) ((status) == EVX_SUCCESS)
#define evx_failed(status) (!evx_succeeded(status))

namespace evx {

// LSB-first bit FIFO (bitstream.h:43-92).  Capacity and indices are in BITS.
class bit_stream
{
    uint32 read_index;
    uint32 write_index;
    uint32 data_capacity;      // bytes
    uint8 *data_store;

public:
    bit_stream();
    bit_stream(uint32 size);
    bit_stream(void *bytes, uint32 size);
    virtual ~bit_stream();

    uint8 *query_data() const;
    uint32 query_capacity() const;
    uint32 query_occupancy() const;
    uint32 query_byte_occupancy() const;
    uint32 resize_capacity(uint32 size_in_bits);
    evx_status assign(void *bytes, uint32 size);

    void seek(uint32 offset);      // read index only (and it overshoots exactly like bitstream.cpp:87-95)
    void clear();
    void empty();
    bool is_empty() const;
    bool is_full() const;

    evx_status write_byte(uint8 value);
    evx_status write_bit(uint8 value);
    evx_status write_bytes(void *data, uint32 count);
    evx_status write_bits(void *data, uint32 count);

    evx_status read_byte(void *data);
    evx_status read_bit(void *data);
    evx_status read_bytes(void *data, uint32 count);
    evx_status read_bits(void *data, uint32 count);

    evx_status peek_byte(void *data);
    evx_status peek_bit(void *data);
    evx_status peek_bytes(void *data, uint32 count);
    evx_status peek_bits(void *data, uint32 count);

    // additions
    uint32 query_read_index() const { return read_index; }
    uint32 query_write_index() const { return write_index; }

private:
    bit_stream(const bit_stream &);
    bit_stream &operator=(const bit_stream &);
};

enum EVX_PEEK_STATE      // evx1.h:53-62
{
    EVX_PEEK_SOURCE = 0,
    EVX_PEEK_PREDICTION,
    EVX_PEEK_BLOCK_TABLE,
    EVX_PEEK_QUANT_TABLE,
    EVX_PEEK_SPMP_TABLE,
    EVX_PEEK_BLOCK_VARIANCE,
    EVX_PEEK_DESTINATION,
};

// addition: per-frame timing/size facts of the last encode() call
struct evx1_frame_stats
{
    double gpu_ms;           // submit -> records on the host (H2D, kernels, D2H)
    double entropy_ms;       // host serialisation (Exp-Golomb + ABAC)
    uint32 slice_bits;
    uint32 noncopy_blocks;
    uint32 d2h_bytes;        // what came back from the device for this frame (bin string, or table + records)
    double wait_ms;          // of gpu_ms: how long the host waited for the device when it asked for the frame's results
};

class evx1_encoder
{
protected:
    virtual ~evx1_encoder() {}

public:
    virtual evx_status clear() = 0;
    virtual evx_status insert_intra() = 0;
    virtual evx_status set_quality(uint8 quality) = 0;                                  // clipped to 1..31
    virtual evx_status encode(void *image, uint32 width, uint32 height, bit_stream *output) = 0;   // R8G8B8 in, appends to output
    virtual evx_status peek(EVX_PEEK_STATE peek_state, void *output) = 0;              // debug views, evx1enc.cpp:170-305
    virtual evx_status last_frame_stats(evx1_frame_stats *out) = 0;                    // addition

    // additions: encode() == submit() + collect().  submit queues the frame on the device (colour
    // conversion, motion search, transform/quantisation, reconstruction, deblocking, binarisation) and
    // returns; collect appends what encode() would have appended for the oldest uncollected frame (stream
    // header on the first frame, frame descriptor, slice).  Several frames may be uncollected: as many as the device
    // library has frame slots (evx1_config::frame_slots, ten by default), which follow each other macroblock by macroblock
    // on the device, and up to twelve retired ones whose slices are being entropy-coded on the session's coder threads
    // (evx1_config::coder_threads, six by default; a slice's coder needs nothing from other frames).  With submit(n+k) before
    // collect(n) the host entropy stage, the device's work and the host->device copies of different frames all run at the
    // same time; k should exceed the frame slots by a few frames, or every collect() waits for a coder that has only just
    // started.  collect() returns the frames in order, the same bytes encode() appends.  `image` must stay unchanged until
    // the frame's own collect() returns.  submit returns EVX_ERROR_NOT_READY when the device holds all the frames it takes
    // and twelve retired ones wait to be collected; encode() when any frame is uncollected; collect() when none is.
    virtual evx_status submit(void *image, uint32 width, uint32 height) = 0;
    virtual evx_status collect(bit_stream *output) = 0;
};

class evx1_decoder
{
protected:
    virtual ~evx1_decoder() {}

public:
    virtual evx_status clear() = 0;
    virtual evx_status decode(bit_stream *input, void *output) = 0;
    virtual evx_status last_frame_stats(evx1_frame_stats *out) = 0;                    // addition (entropy_ms = unserialize)

    // additions: decode() == submit() + collect().  submit takes one frame out of input (and empties it, like
    // decode) and hands its slice to a parser thread; collect merges the oldest submitted frame into the stream's
    // state, runs the pixel pipeline and writes its picture.  The arithmetic decoding of a slice needs nothing from
    // other frames, so with submit(n+1) ... submit(n+k) before collect(n) the slices of consecutive frames are decoded
    // concurrently (eight parser threads) while the caller's thread runs the device; collect(n) also hands frame n+1 to the
    // device, if it is parsed, before it waits for picture n.  At most twelve frames may be
    // uncollected (EVX_ERROR_NOT_READY otherwise, and for decode() with any frame uncollected, and for collect
    // with none).  decode() itself parses on the calling thread.
    virtual evx_status submit(bit_stream *input) = 0;
    virtual evx_status collect(void *output) = 0;
};

// addition: the reference's config.h switches at run time
struct evx1_config
{
    int32 device;            // CUDA device ordinal
    int32 ref_count;         // EVX_REFERENCE_FRAME_COUNT (default 4)
    int32 linear_quant;      // EVX_ENABLE_LINEAR_QUANTIZATION (default 0)
    int32 deblocking;        // EVX_ENABLE_DEBLOCKING (default 1)
    int32 periodic_intra;    // EVX_PERIODIC_INTRA_RATE (default 3600; 0 = never)
    int32 default_quality;   // EVX_DEFAULT_QUALITY_LEVEL (default 8)
    int32 frame_slots;       // encoder: frames of the stream in flight on the device at once (0 = the device library's default)
    int32 coder_threads;     // encoder: threads running the arithmetic coder of retired frames (0 = default 6, at most 8)
    int32 device_frames;     // 1: the image pointers of encode()/submit() and decode()/collect() are DEVICE memory
                             //    (RGB8, tightly pitched) on cfg.device -- capture / present paths that never touch the host
};

void default_config(evx1_config *cfg);

evx_status create_encoder(evx1_encoder **output);       // evx1.h:118-122
evx_status create_decoder(evx1_decoder **output);
evx_status destroy_encoder(evx1_encoder *input);
evx_status destroy_decoder(evx1_decoder *input);
evx_status create_encoder_ex(const evx1_config &cfg, evx1_encoder **output);
evx_status create_decoder_ex(const evx1_config &cfg, evx1_decoder **output);

}  // namespace evx

#endif
