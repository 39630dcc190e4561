/* oracle/evx_oracle.h -- TEST INFRASTRUCTURE ONLY.
 *
 * Plain-C restatement of the reference's (hinike/cairo, EVX-1) per-macroblock
 * pixel pipeline and of its slice (un)serialiser, used as the CHECKER for the
 * CUDA path.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline
 * leg may load this; the product never does.
 *
 * Parity status: PINNED.  tests/test_oracle_vs_ref.py checks every stage of
 * this file bit-for-bit against the unmodified reference compiled by
 * oracle/Makefile (oracle/_ref/libevxref_*.so), and tests/golden/ holds
 * fixtures generated from that reference for boxes without /root/reference.
 *
 * Unlike the reference, whose ring size / quantiser family / deblocking are
 * compile-time switches (config.h:38-53), all three are run-time here.
 */
#ifndef EVX_ORACLE_H
#define EVX_ORACLE_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* Same 16-byte layout as the reference's evx_block_desc (common.h:78-95, #pragma pack(2)). */
#pragma pack(push, 2)
typedef struct evxo_block_desc
{
    int32_t block_type;          /* types.h:68-87 bit codes: intra | motion<<1 | copy<<2 */
    uint8_t prediction_target;   /* ring offset of the reference frame (0 = current) */
    int16_t motion_x;
    int16_t motion_y;
    uint8_t sp_pred;
    uint8_t sp_amount;           /* 0 half, 1 quarter */
    uint8_t sp_index;            /* direction code, motion.cpp:61-109 */
    uint8_t q_index;
    int16_t variance;
} evxo_block_desc;
#pragma pack(pop)

typedef struct evxo_ctx evxo_ctx;

evxo_ctx *evxo_create(int width, int height, int ref_count, int linear_quant, int deblocking);
void evxo_destroy(evxo_ctx *c);
void evxo_reset(evxo_ctx *c);                       /* zero every plane and the block table */

int evxo_aligned_width(const evxo_ctx *c);
int evxo_aligned_height(const evxo_ctx *c);
int evxo_block_count(const evxo_ctx *c);

/* which: 0 source YUV ("input_cache"), 1 quantised coefficients ("output_cache" on the
 * encoder, "input_cache" on the decoder), 2 ring slot `slot`.  comp: 0 Y, 1 U, 2 V. */
int16_t *evxo_plane(evxo_ctx *c, int which, int slot, int comp);
evxo_block_desc *evxo_block_table(evxo_ctx *c);

/* stages (frame_type 0 intra / 1 inter) */
void evxo_convert_in(evxo_ctx *c, const uint8_t *rgb);                               /* convert.cpp:95-160 */
int  evxo_encode_slice(evxo_ctx *c, int frame_type, uint32_t index, int quality);   /* encode.cpp:165-203 */
int  evxo_decode_slice(evxo_ctx *c, int frame_type, uint32_t index);                /* decode.cpp:146-170 */
void evxo_deblock(evxo_ctx *c, uint32_t index);                                      /* deblock.cpp:277-284 */
void evxo_deblock_tiled(evxo_ctx *c, uint32_t index);   /* same result, 8x8 tile order used by the GPU */
void evxo_convert_out(evxo_ctx *c, uint32_t index, uint8_t *rgb);                    /* convert.cpp:162-223 */
uint32_t evxo_serialize_slice(evxo_ctx *c, uint8_t *out, uint32_t cap_bytes);        /* serialize.cpp:319-340 */
int  evxo_unserialize_slice(evxo_ctx *c, const uint8_t *in, uint32_t nbits);        /* unserialize.cpp:321-341 */

/* single searches against the current ring state (motion.cpp:354-494) */
int32_t evxo_inter_prediction(evxo_ctx *c, uint32_t index, int quality, int px, int py, int offset, evxo_block_desc *out);
int32_t evxo_intra_prediction(evxo_ctx *c, uint32_t index, int quality, int px, int py, evxo_block_desc *out);

/* work counters for the roofline figure: evaluated full-pel candidates
 * and sub-pel tests (half and quarter counted separately) since the last reset */
void evxo_get_counters(const evxo_ctx *c, uint64_t *fullpel, uint64_t *subpel);
void evxo_reset_counters(evxo_ctx *c);

/* the wavefront schedule check of SURVEY H3: encode_slice visiting macroblocks in
 * wavefront order (step = bx + 3*by; reverse!=0 walks each step backwards) */
int evxo_encode_slice_wavefront(evxo_ctx *c, int frame_type, uint32_t index, int quality, int reverse);

#ifdef __cplusplus
}
#endif
#endif
