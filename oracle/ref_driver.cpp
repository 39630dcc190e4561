// oracle/ref_driver.cpp -- TEST INFRASTRUCTURE ONLY (never linked into the product).
//
// A thin extern "C" driver around the UNMODIFIED reference sources in
// /root/reference.  oracle/Makefile compiles this file together with the
// reference's own .cpp files (read where they lie, nothing is copied) into
// oracle/_ref/libevxref_<variant>.so.  Two ways in:
//
//   * the public API  (evx1_encoder::encode / evx1_decoder::decode, evx1.h:66-113)
//   * a staged path that owns an evx_context and calls the reference's own
//     stage functions one by one (same sequence as engine_encode_frame,
//     encode.cpp:205-232 and engine_decode_frame, decode.cpp:172-198) so that
//     tests can read the planes / block table between stages.
//
// The internal prototypes are re-declared exactly the way the reference's own
// translation units do it (encode.cpp:12-15, decode.cpp:12-13, evx1enc.cpp:9).

#include <string.h>
#include <stdint.h>
#include <new>

#include "evx1.h"
#include "common.h"
#include "config.h"
#include "convert.h"
#include "motion.h"
#include "version.h"

namespace evx {
evx_status encode_slice(const evx_frame &frame, evx_context *context);
evx_status decode_slice(const evx_frame &frame, evx_context *context);
evx_status serialize_slice(const evx_frame &frame, evx_context *context, bit_stream *output);
evx_status unserialize_slice(bit_stream *input, evx_context *context);
evx_status deblock_image_filter(evx_block_desc *block_table, image_set *target_image);
}

using namespace evx;

namespace {

struct stage_ctx
{
    evx_context context;
    evx_frame frame;
    uint32 width, height;             // visible size
    uint32 aligned_width, aligned_height;
};

image_set *pick_set(stage_ctx *s, int which, int slot)
{
    switch (which)
    {
        case 0: return &s->context.cache_bank.input_cache;
        case 1: return &s->context.cache_bank.output_cache;
        case 2: return &s->context.cache_bank.prediction_cache[slot % EVX_REFERENCE_FRAME_COUNT];
    }
    return NULL;
}

} // namespace

extern "C" {

// ---------------------------------------------------------------- build facts
int evxref_ref_count(void)            { return EVX_REFERENCE_FRAME_COUNT; }
int evxref_linear_quant(void)         { return EVX_ENABLE_LINEAR_QUANTIZATION; }
int evxref_deblocking(void)           { return EVX_ENABLE_DEBLOCKING; }
int evxref_sizeof_block_desc(void)    { return (int) sizeof(evx_block_desc); }
int evxref_sizeof_header(void)        { return (int) sizeof(evx_header); }
int evxref_sizeof_frame(void)         { return (int) sizeof(evx_frame); }

// ---------------------------------------------------------------- public API
void *evxref_encoder_create(void)
{
    evx1_encoder *enc = NULL;
    if (create_encoder(&enc) != EVX_SUCCESS) return NULL;
    return enc;
}
void evxref_encoder_destroy(void *h)           { destroy_encoder((evx1_encoder *) h); }
int evxref_encoder_clear(void *h)              { return ((evx1_encoder *) h)->clear(); }
int evxref_encoder_insert_intra(void *h)       { return ((evx1_encoder *) h)->insert_intra(); }
int evxref_encoder_set_quality(void *h, int q) { return ((evx1_encoder *) h)->set_quality((uint8) q); }

// Encodes one frame into a fresh bit_stream (the reference's transport unit is
// one frame per stream, evx1dec.cpp:120) and copies the bytes out.
// storage is pre-zeroed by the caller-visible contract: bytes beyond *out_bits
// are whatever the reference left -- compare at bit length.
int evxref_encoder_encode(void *h, const uint8_t *rgb, uint32_t w, uint32_t hgt,
                          uint8_t *out, uint32_t out_cap_bytes, uint32_t *out_bits)
{
    bit_stream bs(out_cap_bytes * 8);
    memset(bs.query_data(), 0, out_cap_bytes);   // determinism of the trailing bits
    int st = ((evx1_encoder *) h)->encode(const_cast<uint8_t *>(rgb), w, hgt, &bs);
    uint32 bits = bs.query_occupancy();
    memcpy(out, bs.query_data(), (bits + 7) >> 3);
    *out_bits = bits;
    return st;
}

int evxref_encoder_peek(void *h, int state, uint8_t *rgb_out)
{
    return ((evx1_encoder *) h)->peek((EVX_PEEK_STATE) state, rgb_out);
}

void *evxref_decoder_create(void)
{
    evx1_decoder *dec = NULL;
    if (create_decoder(&dec) != EVX_SUCCESS) return NULL;
    return dec;
}
void evxref_decoder_destroy(void *h) { destroy_decoder((evx1_decoder *) h); }
int evxref_decoder_clear(void *h)    { return ((evx1_decoder *) h)->clear(); }

int evxref_decoder_decode(void *h, const uint8_t *bytes, uint32_t nbits, uint8_t *rgb_out)
{
    uint32 nbytes = (nbits + 7) >> 3;
    bit_stream bs(nbytes * 8 + 64);
    bs.write_bits(const_cast<uint8_t *>(bytes), nbits);
    return ((evx1_decoder *) h)->decode(&bs, rgb_out);
}

// ---------------------------------------------------------------- staged path
void *evxref_stage_create(uint32_t w, uint32_t h)
{
    stage_ctx *s = new (std::nothrow) stage_ctx;
    if (!s) return NULL;
    s->width = w; s->height = h;
    s->aligned_width = align(w, EVX_MACROBLOCK_SIZE);      // evx1enc.cpp:79-80
    s->aligned_height = align(h, EVX_MACROBLOCK_SIZE);
    clear_frame(&s->frame);
    if (initialize_context(s->aligned_width, s->aligned_height, &s->context) != EVX_SUCCESS)
    {
        delete s;
        return NULL;
    }
    return s;
}

void evxref_stage_destroy(void *h)
{
    stage_ctx *s = (stage_ctx *) h;
    clear_context(&s->context);
    delete s;
}

void evxref_stage_set_frame(void *h, int type, uint32_t index, int quality)
{
    stage_ctx *s = (stage_ctx *) h;
    s->frame.type = (EVX_FRAME_TYPE) type;
    s->frame.index = index;
    s->frame.quality = (uint16) quality;
}

int evxref_stage_convert_in(void *h, const uint8_t *rgb)
{
    stage_ctx *s = (stage_ctx *) h;
    image input;
    if (evx_failed(create_image(EVX_IMAGE_FORMAT_R8G8B8, const_cast<uint8_t *>(rgb), s->width, s->height, &input)))
        return EVX_ERROR_EXECUTION_FAILURE;
    return convert_image(input, &s->context.cache_bank.input_cache);
}

int evxref_stage_encode_slice(void *h)   { stage_ctx *s = (stage_ctx *) h; return encode_slice(s->frame, &s->context); }
int evxref_stage_decode_slice(void *h)   { stage_ctx *s = (stage_ctx *) h; return decode_slice(s->frame, &s->context); }

int evxref_stage_deblock(void *h)
{
    stage_ctx *s = (stage_ctx *) h;
    uint32 dest = query_prediction_index_by_offset(s->frame, 0);
    return deblock_image_filter(s->context.block_table, &s->context.cache_bank.prediction_cache[dest]);
}

int evxref_stage_serialize(void *h, uint8_t *out, uint32_t out_cap_bytes, uint32_t *out_bits)
{
    stage_ctx *s = (stage_ctx *) h;
    bit_stream bs(out_cap_bytes * 8);
    memset(bs.query_data(), 0, out_cap_bytes);
    int st = serialize_slice(s->frame, &s->context, &bs);
    uint32 bits = bs.query_occupancy();
    memcpy(out, bs.query_data(), (bits + 7) >> 3);
    *out_bits = bits;
    return st;
}

int evxref_stage_unserialize(void *h, const uint8_t *bytes, uint32_t nbits)
{
    stage_ctx *s = (stage_ctx *) h;
    uint32 nbytes = (nbits + 7) >> 3;
    bit_stream bs(nbytes * 8 + 64);
    bs.write_bits(const_cast<uint8_t *>(bytes), nbits);
    return unserialize_slice(&bs, &s->context);
}

int evxref_stage_convert_out(void *h, uint8_t *rgb_out)
{
    stage_ctx *s = (stage_ctx *) h;
    image output;
    if (evx_failed(create_image(EVX_IMAGE_FORMAT_R8G8B8, rgb_out, s->width, s->height, &output)))
        return EVX_ERROR_EXECUTION_FAILURE;
    uint32 dest = query_prediction_index_by_offset(s->frame, 0);
    return convert_image(s->context.cache_bank.prediction_cache[dest], &output);
}

// one (macroblock, reference) search against the current ring state
int evxref_stage_inter_prediction(void *h, int px, int py, int offset, void *desc_out, int32_t *sad_out)
{
    stage_ctx *s = (stage_ctx *) h;
    macroblock src;
    create_macroblock(s->context.cache_bank.input_cache, px, py, &src);
    evx_block_desc d;
    memset(&d, 0, sizeof(d));
    *sad_out = calculate_inter_prediction(s->frame, src, px, py, &s->context.cache_bank, (uint16) offset, &d);
    memcpy(desc_out, &d, sizeof(d));
    return 0;
}

int evxref_stage_intra_prediction(void *h, int px, int py, void *desc_out, int32_t *sad_out)
{
    stage_ctx *s = (stage_ctx *) h;
    macroblock src;
    create_macroblock(s->context.cache_bank.input_cache, px, py, &src);
    evx_block_desc d;
    memset(&d, 0, sizeof(d));
    *sad_out = calculate_intra_prediction(s->frame, src, px, py, &s->context.cache_bank, &d);
    memcpy(desc_out, &d, sizeof(d));
    return 0;
}

// plane access: which = 0 input_cache, 1 output_cache (quantised coefficients),
// 2 prediction_cache[slot]; comp = 0 Y, 1 U, 2 V.  Tightly pitched int16.
static image *pick_plane(stage_ctx *s, int which, int slot, int comp)
{
    image_set *set = pick_set(s, which, slot);
    if (!set) return NULL;
    return comp == 0 ? set->query_y_image() : comp == 1 ? set->query_u_image() : set->query_v_image();
}

int evxref_stage_get_plane(void *h, int which, int slot, int comp, int16_t *out)
{
    image *img = pick_plane((stage_ctx *) h, which, slot, comp);
    if (!img) return 1;
    memcpy(out, img->query_data(), (size_t) img->query_width() * img->query_height() * 2);
    return 0;
}

int evxref_stage_set_plane(void *h, int which, int slot, int comp, const int16_t *in)
{
    image *img = pick_plane((stage_ctx *) h, which, slot, comp);
    if (!img) return 1;
    memcpy(img->query_data(), in, (size_t) img->query_width() * img->query_height() * 2);
    return 0;
}

int evxref_stage_get_block_table(void *h, void *out)
{
    stage_ctx *s = (stage_ctx *) h;
    uint32 n = s->context.width_in_blocks * s->context.height_in_blocks;
    memcpy(out, s->context.block_table, n * sizeof(evx_block_desc));
    return (int) n;
}

int evxref_stage_set_block_table(void *h, const void *in)
{
    stage_ctx *s = (stage_ctx *) h;
    uint32 n = s->context.width_in_blocks * s->context.height_in_blocks;
    memcpy(s->context.block_table, in, n * sizeof(evx_block_desc));
    return (int) n;
}

} // extern "C"
