/*
 * kpeg_synth.h -- deterministic synthetic baseline-JPEG encoder (host only, no CUDA).
 *
 * The reference's own encoder is non-functional (README.md:23; src/Encoder.cpp writes
 * "output.jpg" with 12.6 dB PSNR), so benchmark and parity inputs come from this committed
 * generator instead (BASELINE.json north_star).  It has no counterpart in the reference; the
 * only things it takes from it are the facts about which container layout the reference's
 * parser accepts (src/Decoder.cpp:53-75): segments are written in exactly the
 * order of misc/images/lena.jpg -- SOI, APP0(JFIF 1.1), DQT id0, DQT id1, SOF0, DHT 0x00,
 * 0x10, 0x01, 0x11, [DRI], SOS, scan, EOI -- with the ITU-T T.81 Annex K quantisation tables
 * (scaled by the usual IJG quality rule) and Annex K Huffman tables.
 *
 * "Twin" streams: two calls that differ only in KPEG_SYNTH_EMIT_RESTART, or only in
 * `file_components` (1 vs 3 with KPEG_SYNTH_GRAY_CONTENT), carry IDENTICAL quantised
 * coefficients.  The reference cannot parse DRI/RSTn or 1-component files (SURVEY F2,F3), so
 * the 3-component, marker-free twin is what it decodes while the GPU path decodes the other.
 * KPEG_SYNTH_QUIRK_FREE nudges DC values so that no block has "DC difference == 0 and a
 * non-zero AC coefficient" in EITHER twin (SURVEY F1,F4): such streams decode identically
 * with and without the reference's AC-dropping quirk.
 */
#ifndef KPEG_SYNTH_H
#define KPEG_SYNTH_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define KPEG_SYNTH_EMIT_RESTART 1u /* write DRI + RSTn (otherwise restart_interval only shapes the nudge) */
#define KPEG_SYNTH_GRAY_CONTENT 2u /* luminance-only content; chroma blocks (if any) are all zero */
#define KPEG_SYNTH_QUIRK_FREE 4u   /* see above */
#define KPEG_SYNTH_NON_INTERLEAVED 8u /* three-component files: one scan per component (T.81 A.2.3) instead of one interleaved scan */

typedef struct kpeg_synth_params {
    int32_t width, height;       /* pixels; any size >= 1 (edges are replicated to whole blocks) */
    int32_t file_components;     /* 1 or 3 */
    int32_t quality;             /* IJG quality 1..100 */
    int32_t restart_interval;    /* in MCUs, 0 = none */
    uint32_t flags;              /* KPEG_SYNTH_* */
    uint64_t seed;               /* content seed */
    int32_t noise_amp;           /* noise amplitude, 0 = default (10, sigma ~ 5.8 grey levels) */
    int32_t threads;             /* worker threads, 0 = hardware concurrency */
} kpeg_synth_params;

/* Encodes one image.  On success returns 0 and hands back a malloc'd buffer the caller must
 * release with kpeg_synth_free.  Returns -1 on bad parameters, -2 on allocation failure. */
int kpeg_synth_encode(const kpeg_synth_params *p, uint8_t **out, size_t *out_len);
void kpeg_synth_free(uint8_t *buf);

/* The source pixels the encoder compressed (before the lossy steps), for PSNR reporting.
 * dst is [height][width][file_components==1 or GRAY_CONTENT ? 1 : 3]. */
int kpeg_synth_pixels(const kpeg_synth_params *p, uint8_t *dst);

#ifdef __cplusplus
}
#endif
#endif
