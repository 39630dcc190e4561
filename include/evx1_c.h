/* include/evx1_c.h -- flat C view of the C++ public API (cairo_b200/csrc/host/evx1.h, itself
 * source-compatible with the reference's evx1.h:66-122) for callers that bind through an FFI
 * (ctypes in tests/ and bench.py).  One call = one method of evx1_encoder / evx1_decoder; the
 * transport unit is one frame per bit_stream, as in the reference (evx1dec.cpp:120).
 * Status codes are the reference's evx_status values (base.h:152-169). */
#ifndef EVX1_C_H
#define EVX1_C_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct evx1c_encoder evx1c_encoder;
typedef struct evx1c_decoder evx1c_decoder;

/* ref_count / linear_quant / deblocking / periodic_intra / default_quality < 0 select the
 * reference's config.h defaults (4, 0, 1, 3600, 8). */
evx1c_encoder *evx1c_encoder_create(int device, int ref_count, int linear_quant, int deblocking, int periodic_intra, int default_quality);
/* the same plus the additions of this build (evx1_config, cairo_b200/csrc/host/evx1.h): frames of the stream in flight on
 * the device (0 = default), arithmetic-coder threads (0 = default), and device_frames = 1: the rgb pointers given to
 * encode/submit are DEVICE memory on `device` (RGB8, tightly pitched; the frame must stay unchanged until collected). */
evx1c_encoder *evx1c_encoder_create_ex(int device, int ref_count, int linear_quant, int deblocking, int periodic_intra, int default_quality,
                                       int frame_slots, int coder_threads, int device_frames);
void evx1c_encoder_destroy(evx1c_encoder *e);
int evx1c_encoder_clear(evx1c_encoder *e);                       /* evx1_encoder::clear */
int evx1c_encoder_insert_intra(evx1c_encoder *e);                /* evx1_encoder::insert_intra */
int evx1c_encoder_set_quality(evx1c_encoder *e, int quality);    /* evx1_encoder::set_quality */
/* evx1_encoder::encode into a fresh bit_stream; copies ceil(out_bits/8) bytes to out. */
int evx1c_encoder_encode(evx1c_encoder *e, const uint8_t *rgb, uint32_t width, uint32_t height,
                         uint8_t *out, uint32_t out_cap_bytes, uint32_t *out_bits);
/* evx1_encoder::submit / collect (additions of this build: encode == submit + collect; submit(n+1) before
 * collect(n) overlaps the host entropy stage of frame n with the device's work on frame n+1).  rgb must stay
 * unchanged until the next submit or the frame's own collect returns. */
int evx1c_encoder_submit(evx1c_encoder *e, const uint8_t *rgb, uint32_t width, uint32_t height);
int evx1c_encoder_collect(evx1c_encoder *e, uint8_t *out, uint32_t out_cap_bytes, uint32_t *out_bits);
/* evx1_encoder::peek (evx1enc.cpp:170-305); state = EVX_PEEK_STATE (0 source, 2 block table, 3 quant table, 4 sub-pel
 * table, 5 block variance, 6 destination); rgb_out is width*height*3 bytes. */
int evx1c_encoder_peek(evx1c_encoder *e, int state, uint8_t *rgb_out);
double evx1c_encoder_wait_ms(evx1c_encoder *e);      /* last collected frame: host time spent waiting for the device */
int evx1c_encoder_stats(evx1c_encoder *e, double *gpu_ms, double *entropy_ms, uint32_t *slice_bits, uint32_t *noncopy_blocks, uint32_t *d2h_bytes);

evx1c_decoder *evx1c_decoder_create(int device, int linear_quant, int deblocking);
/* device_frames = 1: rgb_out of decode/collect is DEVICE memory on `device` (the picture never visits the host) */
evx1c_decoder *evx1c_decoder_create_ex(int device, int linear_quant, int deblocking, int device_frames);
void evx1c_decoder_destroy(evx1c_decoder *d);
int evx1c_decoder_clear(evx1c_decoder *d);                       /* evx1_decoder::clear */
int evx1c_decoder_stats(evx1c_decoder *d, double *gpu_ms, double *entropy_ms);   /* last frame: submit->collect, unserialize */
/* evx1_decoder::decode of one frame (nbits bits at data); rgb_out is width*height*3 bytes. */
int evx1c_decoder_decode(evx1c_decoder *d, const uint8_t *data, uint32_t nbits, uint8_t *rgb_out);

/* evx1_decoder::submit / collect (additions: decode == submit + collect; submit(n+1) before collect(n) overlaps
 * the host entropy decoding of frame n+1 with the device's work on frame n). */
int evx1c_decoder_submit(evx1c_decoder *d, const uint8_t *data, uint32_t nbits);
int evx1c_decoder_collect(evx1c_decoder *d, uint8_t *rgb_out);

/* The host entropy stage on its own (serialize_slice / unserialize_slice of the reference,
 * serialize.cpp:319-340, unserialize.cpp:321-341).  A writer/reader is persistent per stream:
 * it carries the DC-prediction state that the reference keeps in its coefficient planes.
 * table: evxgpu_block_desc[mbw*mbh]; records: int16[n_noncopy][384] in raster order. */
typedef struct evx1c_slice_writer evx1c_slice_writer;
typedef struct evx1c_slice_reader evx1c_slice_reader;
evx1c_slice_writer *evx1c_slice_writer_create(int mbw, int mbh, int ref_count);
void evx1c_slice_writer_destroy(evx1c_slice_writer *w);
int evx1c_slice_writer_serialize(evx1c_slice_writer *w, const void *table, const int16_t *records, uint32_t n_noncopy,
                                 uint8_t *out, uint32_t out_cap_bytes, uint32_t *out_bits);
/* the coder alone, over a bin string from evxgpu_encode_collect_bins */
int evx1c_slice_writer_serialize_bins(evx1c_slice_writer *w, const uint64_t *bins, uint64_t nbins, uint8_t *out, uint32_t out_cap_bytes, uint32_t *out_bits);
evx1c_slice_reader *evx1c_slice_reader_create(int mbw, int mbh, int ref_count);
void evx1c_slice_reader_destroy(evx1c_slice_reader *r);
/* table is in/out (persistent across frames); records_out must hold mbw*mbh*384 int16. */
int evx1c_slice_reader_unserialize(evx1c_slice_reader *r, const uint8_t *data, uint32_t nbits, void *table,
                                   int16_t *records_out, uint32_t *n_noncopy);

/* The reader's two halves (cairo_b200/csrc/host/entropy.h): parse decodes a slice without touching the reader's
 * state (any order, any thread); apply merges parsed slices into the stream's state, in frame order. */
typedef struct evx1c_parsed_slice evx1c_parsed_slice;
evx1c_parsed_slice *evx1c_parsed_slice_create(void);
void evx1c_parsed_slice_destroy(evx1c_parsed_slice *p);
int evx1c_slice_reader_parse(const evx1c_slice_reader *r, const uint8_t *data, uint32_t nbits, evx1c_parsed_slice *out);
int evx1c_slice_reader_apply(evx1c_slice_reader *r, const evx1c_parsed_slice *in, void *table, int16_t *records_out, uint32_t *n_noncopy);

/* include/evxgpu_records.h exported for FFI callers: records of the non-copy macroblocks <-> the reference's persistent
 * coefficient planes (output_cache on the encoder side, input_cache on the decoder side).  Return the record count. */
uint32_t evx1c_scatter_records(const void *table, const int16_t *records, int aligned_width, int aligned_height, int16_t *y, int16_t *u, int16_t *v);
uint32_t evx1c_gather_records(const void *table, const int16_t *y, const int16_t *u, const int16_t *v, int aligned_width, int aligned_height, int16_t *records);

#ifdef __cplusplus
}
#endif
#endif
