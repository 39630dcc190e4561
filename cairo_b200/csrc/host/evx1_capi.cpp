// evx1_capi.cpp -- include/evx1_c.h on top of the C++ API (evx1.h).
#include <string.h>

#include <new>

#include "entropy.h"
#include "evx1.h"
#include "evx1_c.h"

using namespace evx;

struct evx1c_encoder { evx1_encoder *enc; bit_stream bs; };
struct evx1c_decoder { evx1_decoder *dec; bit_stream bs; };

static evx1_config make_config(int device, int ref_count, int linear_quant, int deblocking, int periodic_intra, int default_quality)
{
    evx1_config c;
    default_config(&c);
    c.device = device;
    if (ref_count >= 0) c.ref_count = ref_count;
    if (linear_quant >= 0) c.linear_quant = linear_quant;
    if (deblocking >= 0) c.deblocking = deblocking;
    if (periodic_intra >= 0) c.periodic_intra = periodic_intra;
    if (default_quality >= 0) c.default_quality = default_quality;
    return c;
}

#include "evxgpu_records.h"

extern "C" {

// include/evxgpu_records.h through the library, for FFI callers (the header itself is all a C or C++ binding needs)
uint32_t evx1c_scatter_records(const void *table, const int16_t *records, int aligned_width, int aligned_height, int16_t *y, int16_t *u, int16_t *v)
{
    return evxgpu_scatter_records(static_cast<const evxgpu_block_desc *>(table), records, aligned_width, aligned_height, y, u, v);
}
uint32_t evx1c_gather_records(const void *table, const int16_t *y, const int16_t *u, const int16_t *v, int aligned_width, int aligned_height, int16_t *records)
{
    return evxgpu_gather_records(static_cast<const evxgpu_block_desc *>(table), y, u, v, aligned_width, aligned_height, records);
}

evx1c_encoder *evx1c_encoder_create(int device, int ref_count, int linear_quant, int deblocking, int periodic_intra, int default_quality)
{
    evx1c_encoder *e = new (std::nothrow) evx1c_encoder();
    if (!e) return NULL;
    if (create_encoder_ex(make_config(device, ref_count, linear_quant, deblocking, periodic_intra, default_quality), &e->enc) != EVX_SUCCESS) { delete e; return NULL; }
    return e;
}

evx1c_encoder *evx1c_encoder_create_ex(int device, int ref_count, int linear_quant, int deblocking, int periodic_intra, int default_quality,
                                       int frame_slots, int coder_threads, int device_frames)
{
    evx1c_encoder *e = new (std::nothrow) evx1c_encoder();
    if (!e) return NULL;
    evx1_config c = make_config(device, ref_count, linear_quant, deblocking, periodic_intra, default_quality);
    if (frame_slots > 0) c.frame_slots = frame_slots;
    if (coder_threads > 0) c.coder_threads = coder_threads;
    c.device_frames = device_frames > 0 ? 1 : 0;
    if (create_encoder_ex(c, &e->enc) != EVX_SUCCESS) { delete e; return NULL; }
    return e;
}

void evx1c_encoder_destroy(evx1c_encoder *e) { if (e) { destroy_encoder(e->enc); delete e; } }
int evx1c_encoder_clear(evx1c_encoder *e) { return e ? e->enc->clear() : EVX_ERROR_INVALIDARG; }
int evx1c_encoder_insert_intra(evx1c_encoder *e) { return e ? e->enc->insert_intra() : EVX_ERROR_INVALIDARG; }
int evx1c_encoder_set_quality(evx1c_encoder *e, int quality) { return e ? e->enc->set_quality((uint8) quality) : EVX_ERROR_INVALIDARG; }

int evx1c_encoder_encode(evx1c_encoder *e, const uint8_t *rgb, uint32_t width, uint32_t height, uint8_t *out, uint32_t out_cap_bytes, uint32_t *out_bits)
{
    if (!e || !out || !out_bits) return EVX_ERROR_INVALIDARG;
    if (e->bs.query_capacity() != out_cap_bytes * 8u) e->bs.resize_capacity(out_cap_bytes * 8u);
    e->bs.empty();
    int st = e->enc->encode(const_cast<uint8_t *>(rgb), width, height, &e->bs);
    uint32 bits = e->bs.query_occupancy();
    memcpy(out, e->bs.query_data(), (bits + 7) >> 3);
    *out_bits = bits;
    return st;
}

int evx1c_encoder_submit(evx1c_encoder *e, const uint8_t *rgb, uint32_t width, uint32_t height)
{
    if (!e) return EVX_ERROR_INVALIDARG;
    return e->enc->submit(const_cast<uint8_t *>(rgb), width, height);
}

int evx1c_encoder_collect(evx1c_encoder *e, uint8_t *out, uint32_t out_cap_bytes, uint32_t *out_bits)
{
    if (!e || !out || !out_bits) return EVX_ERROR_INVALIDARG;
    if (e->bs.query_capacity() != out_cap_bytes * 8u) e->bs.resize_capacity(out_cap_bytes * 8u);
    e->bs.empty();
    int st = e->enc->collect(&e->bs);
    uint32 bits = e->bs.query_occupancy();
    memcpy(out, e->bs.query_data(), (bits + 7) >> 3);
    *out_bits = bits;
    return st;
}

int evx1c_encoder_peek(evx1c_encoder *e, int state, uint8_t *rgb_out)
{
    if (!e) return EVX_ERROR_INVALIDARG;
    return e->enc->peek((EVX_PEEK_STATE) state, rgb_out);
}

double evx1c_encoder_wait_ms(evx1c_encoder *e)
{
    evx1_frame_stats s;
    if (!e || e->enc->last_frame_stats(&s) != EVX_SUCCESS) return 0.0;
    return s.wait_ms;
}

int evx1c_encoder_stats(evx1c_encoder *e, double *gpu_ms, double *entropy_ms, uint32_t *slice_bits, uint32_t *noncopy_blocks, uint32_t *d2h_bytes)
{
    if (!e) return EVX_ERROR_INVALIDARG;
    evx1_frame_stats s;
    int st = e->enc->last_frame_stats(&s);
    if (gpu_ms) *gpu_ms = s.gpu_ms;
    if (entropy_ms) *entropy_ms = s.entropy_ms;
    if (slice_bits) *slice_bits = s.slice_bits;
    if (noncopy_blocks) *noncopy_blocks = s.noncopy_blocks;
    if (d2h_bytes) *d2h_bytes = s.d2h_bytes;
    return st;
}

evx1c_decoder *evx1c_decoder_create(int device, int linear_quant, int deblocking)
{
    evx1c_decoder *d = new (std::nothrow) evx1c_decoder();
    if (!d) return NULL;
    if (create_decoder_ex(make_config(device, -1, linear_quant, deblocking, -1, -1), &d->dec) != EVX_SUCCESS) { delete d; return NULL; }
    return d;
}

evx1c_decoder *evx1c_decoder_create_ex(int device, int linear_quant, int deblocking, int device_frames)
{
    evx1c_decoder *d = new (std::nothrow) evx1c_decoder();
    if (!d) return NULL;
    evx1_config c = make_config(device, -1, linear_quant, deblocking, -1, -1);
    c.device_frames = device_frames > 0 ? 1 : 0;
    if (create_decoder_ex(c, &d->dec) != EVX_SUCCESS) { delete d; return NULL; }
    return d;
}

void evx1c_decoder_destroy(evx1c_decoder *d) { if (d) { destroy_decoder(d->dec); delete d; } }
int evx1c_decoder_clear(evx1c_decoder *d) { return d ? d->dec->clear() : EVX_ERROR_INVALIDARG; }
int evx1c_decoder_stats(evx1c_decoder *d, double *gpu_ms, double *entropy_ms)
{
    if (!d) return EVX_ERROR_INVALIDARG;
    evx1_frame_stats s;
    int st = d->dec->last_frame_stats(&s);
    if (gpu_ms) *gpu_ms = s.gpu_ms;
    if (entropy_ms) *entropy_ms = s.entropy_ms;
    return st;
}

int evx1c_decoder_decode(evx1c_decoder *d, const uint8_t *data, uint32_t nbits, uint8_t *rgb_out)
{
    if (!d || !data || !nbits) return EVX_ERROR_INVALIDARG;
    uint32 need = ((nbits + 7) >> 3) * 8 + 64;
    if (d->bs.query_capacity() < need) d->bs.resize_capacity(need);
    d->bs.empty();
    d->bs.write_bits(const_cast<uint8_t *>(data), nbits);
    return d->dec->decode(&d->bs, rgb_out);
}

int evx1c_decoder_submit(evx1c_decoder *d, const uint8_t *data, uint32_t nbits)
{
    if (!d || !data || !nbits) return EVX_ERROR_INVALIDARG;
    uint32 need = ((nbits + 7) >> 3) * 8 + 64;
    if (d->bs.query_capacity() < need) d->bs.resize_capacity(need);
    d->bs.empty();
    d->bs.write_bits(const_cast<uint8_t *>(data), nbits);
    return d->dec->submit(&d->bs);
}

int evx1c_decoder_collect(evx1c_decoder *d, uint8_t *rgb_out)
{
    if (!d || !rgb_out) return EVX_ERROR_INVALIDARG;
    return d->dec->collect(rgb_out);
}

struct evx1c_slice_writer { slice_writer w; };
struct evx1c_slice_reader { slice_reader r; };

evx1c_slice_writer *evx1c_slice_writer_create(int mbw, int mbh, int ref_count)
{
    evx1c_slice_writer *w = new (std::nothrow) evx1c_slice_writer();
    if (w) w->w.configure(mbw, mbh, ref_count);
    return w;
}
void evx1c_slice_writer_destroy(evx1c_slice_writer *w) { delete w; }

int evx1c_slice_writer_serialize(evx1c_slice_writer *w, const void *table, const int16_t *records, uint32_t n_noncopy,
                                 uint8_t *out, uint32_t out_cap_bytes, uint32_t *out_bits)
{
    if (!w || !table || !out || !out_bits) return EVX_ERROR_INVALIDARG;
    uint32_t bits = w->w.serialize(static_cast<const evxgpu_block_desc *>(table), records, n_noncopy);
    if (!bits) return EVX_ERROR_EXECUTION_FAILURE;
    if (((bits + 7) >> 3) > out_cap_bytes) return EVX_ERROR_CAPACITY_LIMIT;
    memcpy(out, w->w.data(), (bits + 7) >> 3);
    *out_bits = bits;
    return EVX_SUCCESS;
}

int evx1c_slice_writer_serialize_bins(evx1c_slice_writer *w, const uint64_t *bins, uint64_t nbins, uint8_t *out, uint32_t out_cap_bytes, uint32_t *out_bits)
{
    if (!w || !bins || !out || !out_bits) return EVX_ERROR_INVALIDARG;
    uint32_t bits = w->w.serialize_bins(bins, nbins);
    if (!bits) return EVX_ERROR_EXECUTION_FAILURE;
    if (((bits + 7) >> 3) > out_cap_bytes) return EVX_ERROR_CAPACITY_LIMIT;
    memcpy(out, w->w.data(), (bits + 7) >> 3);
    *out_bits = bits;
    return EVX_SUCCESS;
}

evx1c_slice_reader *evx1c_slice_reader_create(int mbw, int mbh, int ref_count)
{
    evx1c_slice_reader *r = new (std::nothrow) evx1c_slice_reader();
    if (r) r->r.configure(mbw, mbh, ref_count);
    return r;
}
void evx1c_slice_reader_destroy(evx1c_slice_reader *r) { delete r; }

int evx1c_slice_reader_unserialize(evx1c_slice_reader *r, const uint8_t *data, uint32_t nbits, void *table,
                                   int16_t *records_out, uint32_t *n_noncopy)
{
    if (!r || !data || !table || !records_out || !n_noncopy) return EVX_ERROR_INVALIDARG;
    return r->r.unserialize(data, 0, nbits, static_cast<evxgpu_block_desc *>(table), records_out, n_noncopy);
}

struct evx1c_parsed_slice { parsed_slice p; };

evx1c_parsed_slice *evx1c_parsed_slice_create(void) { return new (std::nothrow) evx1c_parsed_slice(); }
void evx1c_parsed_slice_destroy(evx1c_parsed_slice *p) { delete p; }

int evx1c_slice_reader_parse(const evx1c_slice_reader *r, const uint8_t *data, uint32_t nbits, evx1c_parsed_slice *out)
{
    if (!r || !data || !out) return EVX_ERROR_INVALIDARG;
    return r->r.parse(data, 0, nbits, out->p);
}

int evx1c_slice_reader_apply(evx1c_slice_reader *r, const evx1c_parsed_slice *in, void *table, int16_t *records_out, uint32_t *n_noncopy)
{
    if (!r || !in || !table || !records_out || !n_noncopy) return EVX_ERROR_INVALIDARG;
    return r->r.apply(in->p, static_cast<evxgpu_block_desc *>(table), records_out, n_noncopy);
}

}
