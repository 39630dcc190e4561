/* Force-included (-include) ahead of every reference translation unit by
 * oracle/Makefile.  Defining the reference config.h's include guard makes the
 * reference's own config.h expand to nothing, so this variant's switches win
 * without copying or editing any reference source.
 * Variant: r4_nodeblock  (ring slots R=4, linear quantisation=0, deblocking=0) */
#ifndef __EVX_CONFIG_H__
#define __EVX_CONFIG_H__
#define EVX_ALLOW_INTER_FRAMES            (1)
#define EVX_REFERENCE_FRAME_COUNT         (4)
#define EVX_DEFAULT_QUALITY_LEVEL         (8)
#define EVX_PERIODIC_INTRA_RATE           (3600)
#define EVX_ENABLE_CHROMA_SUPPORT         (1)
#define EVX_QUANTIZATION_ENABLED          (1)
#define EVX_ENABLE_LINEAR_QUANTIZATION    (0)
#define EVX_ROUNDED_QUANTIZATION          (1)
#define EVX_ADAPTIVE_QUANTIZATION         (1)
#define EVX_ENABLE_DEBLOCKING             (0)
#endif
