/* include/evxgpu_records.h -- the two loops a reference-side binding needs around the device seam.
 *
 * evxgpu_encode_collect hands back the coefficient records of the NON-COPY macroblocks in raster order (384 int16
 * each: 16x16 luma row-major, then U 8x8, V 8x8); serialize_slice (serialize.cpp:125-154) reads the reference's
 * persistent coefficient planes context->cache_bank.output_cache, which copy blocks leave untouched
 * (encode.cpp:155-157) and whose stale values feed the DC prediction (serialize.cpp:59-72).  So the
 * binding scatters the records into those planes after every collect -- and gathers them out of
 * cache_bank.input_cache (what unserialize_slice writes, unserialize.cpp:321-341) before evxgpu_decode_submit.
 *
 * Planes are tightly pitched int16: Y aw x ah, U and V (aw/2) x (ah/2), aw/ah = frame size rounded up to 16
 * (evx1enc.cpp:79-80) -- image::query_data() of the R16S images of an image_set (imageset.cpp:45-47).
 * Plain C, no dependencies; returns the number of records moved. */
#ifndef EVXGPU_RECORDS_H
#define EVXGPU_RECORDS_H

#include <stdint.h>
#include <string.h>

#include "evxgpu.h"

#ifdef __cplusplus
extern "C" {
#endif

static inline uint32_t evxgpu_scatter_records(const evxgpu_block_desc *table, const int16_t *records, int aligned_width, int aligned_height,
                                              int16_t *y, int16_t *u, int16_t *v)
{
    const int mbw = aligned_width / 16, mbh = aligned_height / 16, cw = aligned_width / 2;
    uint32_t k = 0;
    for (int by = 0; by < mbh; ++by)
        for (int bx = 0; bx < mbw; ++bx)
        {
            if (table[by * mbw + bx].block_type & 4) continue;             /* copy block: the planes keep their stale values */
            const int16_t *r = records + (size_t) k * EVXGPU_MB_COEFFS;
            for (int j = 0; j < 16; ++j) memcpy(y + (size_t) (by * 16 + j) * aligned_width + bx * 16, r + j * 16, 32);
            for (int j = 0; j < 8; ++j)
            {
                memcpy(u + (size_t) (by * 8 + j) * cw + bx * 8, r + 256 + j * 8, 16);
                memcpy(v + (size_t) (by * 8 + j) * cw + bx * 8, r + 320 + j * 8, 16);
            }
            ++k;
        }
    return k;
}

static inline uint32_t evxgpu_gather_records(const evxgpu_block_desc *table, const int16_t *y, const int16_t *u, const int16_t *v,
                                             int aligned_width, int aligned_height, int16_t *records)
{
    const int mbw = aligned_width / 16, mbh = aligned_height / 16, cw = aligned_width / 2;
    uint32_t k = 0;
    for (int by = 0; by < mbh; ++by)
        for (int bx = 0; bx < mbw; ++bx)
        {
            if (table[by * mbw + bx].block_type & 4) continue;
            int16_t *r = records + (size_t) k * EVXGPU_MB_COEFFS;
            for (int j = 0; j < 16; ++j) memcpy(r + j * 16, y + (size_t) (by * 16 + j) * aligned_width + bx * 16, 32);
            for (int j = 0; j < 8; ++j)
            {
                memcpy(r + 256 + j * 8, u + (size_t) (by * 8 + j) * cw + bx * 8, 16);
                memcpy(r + 320 + j * 8, v + (size_t) (by * 8 + j) * cw + bx * 8, 16);
            }
            ++k;
        }
    return k;
}

#ifdef __cplusplus
}
#endif
#endif
