/* include/evxgpu.h -- C-ABI of the B200 pixel pipeline for the EVX-1 ("Cairo") codec.
 *
 * This is the device boundary introduced under the reference's frame engine.  The
 * reference has no FFI or plugin seam of its own; its pixel pipeline is reached through
 * two internal C++ functions, and this library replaces what those two call:
 *
 *   engine_encode_frame (encode.cpp:205-232)  = convert_image -> encode_slice ->
 *       [serialize_slice, host] -> deblock_image_filter
 *   engine_decode_frame (decode.cpp:172-198)  = [unserialize_slice, host] -> decode_slice ->
 *       deblock_image_filter -> convert_image
 *
 * Everything crossing the boundary is plain data: RGB8 frames in, and per-macroblock
 * records out (the reference's evx_block_desc table, common.h:78-95, plus the quantised
 * coefficients of the non-copy macroblocks) -- exactly what serialize_slice reads
 * (serialize.cpp:125-154, 288-340).  Reference frames never leave the device.
 *
 * One handle = one video stream = one CUDA stream.  Calls return 0 on success or an
 * evx_status-compatible code (base.h:152-169): 1 invalid argument, 3 out of memory,
 * 5 hardware (CUDA) failure, 8 invalid resource, 15 not ready.
 * There is NO CPU fallback: with no usable CUDA device evxgpu_create fails with 5.
 */
#ifndef EVXGPU_H
#define EVXGPU_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* Layout-identical to the reference's evx_block_desc (common.h:78-95, #pragma pack(2)):
 * offsets type@0 target@4 mx@6 my@8 sp_pred@10 sp_amount@11 sp_index@12 q@13 var@14. */
#pragma pack(push, 2)
typedef struct evxgpu_block_desc
{
    int32_t block_type;          /* types.h:68-87: intra | motion<<1 | copy<<2 */
    uint8_t prediction_target;   /* ring offset, 0 = the frame under construction */
    int16_t motion_x;
    int16_t motion_y;
    uint8_t sp_pred;
    uint8_t sp_amount;           /* 0 half-pel, 1 quarter-pel */
    uint8_t sp_index;            /* direction code, motion.cpp:61-109 */
    uint8_t q_index;
    int16_t variance;
} evxgpu_block_desc;
#pragma pack(pop)

/* The reference's compile-time switches (config.h:38-53) as run-time values. */
typedef struct evxgpu_config
{
    int32_t ref_count;        /* EVX_REFERENCE_FRAME_COUNT: ring slots incl. the current frame, 2..8 */
    int32_t linear_quant;     /* EVX_ENABLE_LINEAR_QUANTIZATION */
    int32_t deblocking;       /* EVX_ENABLE_DEBLOCKING */
    int32_t frame_slots;      /* encoder, bin-string output: frames of the stream in flight on the device at once (0 = default 10, at most 16);
                               * 1 = no frame pipeline: frame after frame with the stand-alone kernels (a second frame may still be queued
                               * behind the first) -- the better choice when many streams share one device */
} evxgpu_config;

#define EVXGPU_MB_COEFFS 384   /* 16x16 luma (row-major, stride 16) + 8x8 U + 8x8 V, int16 */

/* kernels timed by evxgpu_get_timing() */
enum { EVXGPU_T_CONVERT_IN = 0, EVXGPU_T_INTER_SEARCH, EVXGPU_T_WAVEFRONT, EVXGPU_T_DEBLOCK,
       EVXGPU_T_DECODE_RECON, EVXGPU_T_CONVERT_OUT, EVXGPU_T_BINS, EVXGPU_T_COUNT };

typedef struct evxgpu_handle evxgpu_handle;

int evxgpu_device_count(void);

/* cuda_stream: a cudaStream_t to run on (e.g. torch's current stream), or NULL to own one. */
int evxgpu_create(int device, int width, int height, const evxgpu_config *cfg, void *cuda_stream, evxgpu_handle **out);
int evxgpu_destroy(evxgpu_handle *h);
int evxgpu_reset(evxgpu_handle *h);                 /* zero the ring, as a fresh context (image.cpp:89) */
int evxgpu_block_count(const evxgpu_handle *h);
int evxgpu_synchronize(evxgpu_handle *h);

/* pinned host memory for callers that want true asynchronous copies */
void *evxgpu_host_alloc(uint64_t bytes);
void evxgpu_host_free(void *p);
void *evxgpu_device_alloc(uint64_t bytes);
void evxgpu_device_free(void *p);
int evxgpu_upload(evxgpu_handle *h, void *dst_device, const void *src_host, uint64_t bytes);

/* ---- encoder: replaces convert_image + encode_slice + deblock_image_filter ----
 * submit: queues H2D of the frame (rgb_is_device=0) or uses the device pointer, then the
 * kernels; returns immediately.  collect: waits, then hands back the block table
 * (block_count entries) and the coefficient records of the non-copy macroblocks in
 * raster order (n_noncopy * 384 int16). */
int evxgpu_encode_submit(evxgpu_handle *h, const uint8_t *rgb, int rgb_is_device,
                         int frame_type, uint32_t frame_index, int quality);
int evxgpu_encode_collect(evxgpu_handle *h, evxgpu_block_desc *table_out, int16_t *records_out, uint32_t *n_noncopy);
/* Optional early copy of the NEXT frame: may be called while a frame is still in flight; the host->device copy
 * runs on a second stream under the current frame's kernels.  The following evxgpu_encode_submit takes rgb = NULL
 * and uses the uploaded frame.  rgb_host must stay unchanged until that submit's frame has been collected (or the
 * handle synchronised).  An upload whose submit never happened is replaced by the next upload. */
int evxgpu_encode_upload(evxgpu_handle *h, const uint8_t *rgb_host);
/* How many submitted, uncollected frames the handle accepts: 1 with table + records output, with bin-string output as
 * many as it has frame slots (evxgpu_config::frame_slots; 2 with EVXGPU_FRAME_OVERLAP=0).  A submit beyond it returns
 * status 8.  Frames in flight run concurrently on the device, each following its predecessor macroblock by macroblock. */
int evxgpu_encode_capacity(const evxgpu_handle *h);

/* The same slice as the string of bins serialize_slice feeds its arithmetic coder (serialize.cpp:156-340:
 * block table by field, then the Y, U, V residual blocks; raw bits and Exp-Golomb codes of golomb.cpp:8-91,
 * DC prediction of serialize.cpp:25-72 including the stale DC of copy blocks), binarised on the device.
 * Only encode_symbol (abac.cpp:97-121) is left for the host: it walks bin i = bit (i & 63) of word (i >> 6).
 * set_output: 0 = table + records (default), 1 = bins only, 2 = both (collect_bins first, then collect).
 * The device keeps the DC state of serialize.cpp:59-72 across frames, so the mode is chosen once per stream.
 * collect_bins: *bins points into the handle's pinned buffer, valid until the next-but-one submit.
 * With mode 1 TWO frames may be in flight: a second evxgpu_encode_submit is accepted while the first frame is
 * uncollected (its kernels are queued right behind the first frame's, so the device does not wait for the host);
 * collect_bins returns the frames in order.  The device-side string buffers are sized for the longest slice the
 * geometry can produce, so they cannot overflow.  Modes 0 and 2 take one frame at a time (status 8 otherwise). */
int evxgpu_set_output(evxgpu_handle *h, int mode);
int evxgpu_encode_collect_bins(evxgpu_handle *h, const uint64_t **bins, uint64_t *nbins, uint32_t *n_noncopy /* may be NULL */);

/* ---- decoder: replaces decode_slice + deblock_image_filter + convert_image ----
 * table: block_count descriptors as unserialize_slice leaves them; records: the
 * non-copy macroblocks' coefficients in raster order. */
int evxgpu_decode_submit(evxgpu_handle *h, const evxgpu_block_desc *table, const int16_t *records, uint32_t n_noncopy,
                         int frame_type, uint32_t frame_index);
int evxgpu_decode_collect(evxgpu_handle *h, uint8_t *rgb_out, int rgb_is_device);
/* The two halves of evxgpu_decode_collect: begin queues the copy-out of the oldest submitted frame's picture (on a copy
 * stream), end waits for it.  Between the two a SECOND frame may be submitted: its staging copies and kernels run under
 * the first frame's copy-out (two staging sets and two RGB buffers per handle).  Status 8 when a second frame is
 * submitted before the first one's copy-out has begun, or a third one at all. */
int evxgpu_decode_collect_begin(evxgpu_handle *h, uint8_t *rgb_out, int rgb_is_device);
int evxgpu_decode_collect_end(evxgpu_handle *h);

/* ---- single stages, for parity tests and profiling ---- */
int evxgpu_stage_convert_in(evxgpu_handle *h, const uint8_t *rgb_host);
int evxgpu_stage_inter_search(evxgpu_handle *h, uint32_t frame_index, int quality);
int evxgpu_stage_get_inter_result(evxgpu_handle *h, int offset, evxgpu_block_desc *desc_out, int32_t *sad_out);
int evxgpu_stage_deblock(evxgpu_handle *h, uint32_t frame_index);
int evxgpu_stage_set_block_table(evxgpu_handle *h, const evxgpu_block_desc *table);

/* which: 0 source YUV, 2 ring slot `slot`; comp 0 Y / 1 U / 2 V; tightly pitched int16,
 * aligned size (width,height rounded up to 16; chroma half of that). */
int evxgpu_peek_plane(evxgpu_handle *h, int which, int slot, int comp, int16_t *out_host);
int evxgpu_poke_plane(evxgpu_handle *h, int which, int slot, int comp, const int16_t *in_host);

/* debug views for evx1_encoder::peek (evx1enc.cpp:170-305): plane set `which` (0 source, 2 ring slot `slot`) through
 * the YUV->RGB kernel into rgb_out (host, width*height*3); the block table of the last frame (block_count entries) */
int evxgpu_peek_rgb(evxgpu_handle *h, int which, int slot, uint8_t *rgb_out_host);
int evxgpu_peek_table(evxgpu_handle *h, evxgpu_block_desc *table_out);

/* per-kernel device time of the last submitted frame, CUDA events on the handle's stream (ms) */
int evxgpu_get_timing(evxgpu_handle *h, float *ms_out /* [EVXGPU_T_COUNT] */);
int evxgpu_enable_timing(evxgpu_handle *h, int on);
/* the same, summed over the encoded frames since the last reset (no per-frame synchronisation by the caller) */
int evxgpu_get_timing_sum(evxgpu_handle *h, double *ms_out /* [EVXGPU_T_COUNT] */, int reset);
/* evaluated full-pel candidates / sub-pel tests since the last reset (SURVEY 8d roofline unit) */
int evxgpu_get_counters(evxgpu_handle *h, uint64_t *fullpel, uint64_t *subpel, int reset);
/* the same split by kernel: out4 = { inter full-pel, inter sub-pel, intra full-pel, intra sub-pel } */
int evxgpu_get_counters_split(evxgpu_handle *h, uint64_t *out4, int reset);
uint64_t evxgpu_launch_count(const evxgpu_handle *h);
/* Average duration (ms) of `reps` back-to-back launches of a streaming kernel on an idle handle: 0 = RGB->YUV, 1 = deblocking
 * (ring slot 0, in place), 2 = YUV->RGB.  A measurement aid (bench.py's HBM fractions); -1 on failure. */
double evxgpu_time_kernel(evxgpu_handle *h, int kind, int reps);
/* device-side clock of the pipeline (bench.py): mark() stamps now; every frame submitted afterwards records a CUDA event on
 * its own stream once its results have left the device, and last_done_ms() is that moment for the frame collected last, in
 * milliseconds since the mark (-1 before any mark, or when that frame was submitted before the mark) */
int evxgpu_timeline_mark(evxgpu_handle *h);
double evxgpu_last_done_ms(evxgpu_handle *h);
/* bytes copied device -> host for the last submitted frame, as queued by submit and the collect calls so far */
uint64_t evxgpu_d2h_bytes(const evxgpu_handle *h);
/* debug: per-row phase cycle sums of the encoder wavefront kernel, 6 x int64 per macroblock row */
int evxgpu_debug_profile(evxgpu_handle *h, int enable, long long *out_host);
/* debug: what a device-side wait that ran out of its time budget was waiting for (out4[0] = 0: none did).  Such a wait
 * traps, the launch fails and every later call reports status 5; EVXGPU_WAIT_BUDGET_MS sets the budget (default 4000, 0 = none). */
int evxgpu_debug_wait_diag(evxgpu_handle *h, unsigned int *out4);
/* tuning knob: number of persistent CTAs of the decoder's wavefront kernel (0 = default) */
int evxgpu_set_wave_grid(evxgpu_handle *h, int ctas);
/* tuning knob: persistent CTAs of the encoder's wavefront kernel (0 = default, ceil(mbw/3) + 4: the number of
 * macroblock rows that can be active at once, plus spares); any value >= 1 is correct */
int evxgpu_set_encode_grid(evxgpu_handle *h, int ctas);
/* tests: shrink the bin-string buffers so that the grow-and-emit-again path of collect_bins runs */
int evxgpu_debug_set_bins_capacity(evxgpu_handle *h, uint32_t bits);

/* integer-pipe micro-benchmark (the roofline denominator MEASURED_PEAKS.json lacks):
 * kind 0 IADD3, 1 VIADDMNMX.S16x2, 2 IDP.2A, 3 the 3:2 mix the search kernels issue.
 * Returns tera-instructions-lanes/s (lane-ops, i.e. one packed op counts once). */
double evxgpu_measure_int_peak(int device, int kind);

const char *evxgpu_last_error(void);

#ifdef __cplusplus
}
#endif
#endif
