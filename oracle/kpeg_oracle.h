/*
 * kpeg_oracle.h -- CPU restatement of libKPEG's baseline-JPEG decode hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under libkpeg_b200/ (the product) may include, link or
 * call this.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
 * reference legs use it, and there only as the checker or the timed CPU arm.
 *
 * Parity status: PINNED.  tests/test_oracle_vs_reference.py checks this restatement
 * byte-for-byte against the compiled, unmodified reference (oracle/_ref/kpeg_ref_quiet, built
 * by oracle/Makefile from /root/reference) on misc/images/lena.jpg and on synthetic twins,
 * and tests/test_oracle_kats.py checks the known-answer vectors of the reference's dormant
 * self-tests (main.cpp:142-346).  Restart intervals, true 1-component files and ragged
 * (non multiple-of-8) sizes are extensions the reference cannot decode (SURVEY F2,F3,F6):
 * for those the restatement follows ITU-T T.81 and is pinned only through twin streams.
 *
 * All file:line citations are relative to /root/reference.
 */
#ifndef KPEG_ORACLE_H
#define KPEG_ORACLE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* flags for kpo_decode */
#define KPO_FLAG_REF_PARITY 1u /* reproduce MCU.cpp:97-104: DC diff == 0 drops the block's AC */

/* return codes */
#define KPO_OK 0
#define KPO_ERR_FORMAT (-1)      /* malformed container                  (reference: ResultCode::ERROR) */
#define KPO_ERR_UNSUPPORTED (-2) /* SOF1/SOF2/subsampling/16-bit DQT...  (reference: ResultCode::TERMINATE) */
#define KPO_ERR_STREAM (-3)      /* entropy-coded data ran out / bad code (reference: UB / hang) */
#define KPO_ERR_NOMEM (-4)

typedef struct kpo_image {
    int32_t width, height; /* SOF0 dimensions */
    int32_t ncomp;         /* 1 or 3 */
    int32_t mcus_x, mcus_y;
    int32_t restart_interval;
    int64_t nblocks;      /* mcus_x*mcus_y*ncomp */
    int64_t scan_bytes;   /* stuffed entropy-coded bytes (incl. RSTn markers) */
    int16_t *coef;        /* [nblocks][64], MCU-interleaved, zig-zag order, DC integrated (App. A.4) */
    uint8_t *pixels;      /* [height][width][ncomp] (RGB interleaved for ncomp==3, gray for 1) */
} kpo_image;

/* Whole-file decode.  Fills `img` (caller frees with kpo_free).  `want_pixels`==0 stops after
 * the coefficient stage (cheap; the IDCT restatement is deliberately slow). */
int kpo_decode(const uint8_t *file, size_t len, uint32_t flags, int want_pixels, kpo_image *img);
void kpo_free(kpo_image *img);

/* Worker threads used by the reconstruction stage (IDCT + colour). Default 1. */
void kpo_set_threads(int n);

/* Exact PPM header of Image.cpp:124-127.  Returns its length. */
int kpo_ppm_header(int width, int height, char *buf, size_t cap);

/* ---- building blocks exported for known-answer tests ---- */

/* HuffmanTree.cpp:106-157 (== T.81 Annex C): canonical codes.  counts[16], symbols in file
 * order; writes code value / length per symbol (same order).  Returns number of symbols. */
int kpo_huff_codes(const uint8_t counts[16], const uint8_t *symbols, uint16_t *codes, uint8_t *lens);
/* HuffmanTree.cpp:164-193 `contains`: returns symbol 0..255 if `bits` (ASCII '0'/'1') is a
 * complete code, -1 otherwise. */
int kpo_huff_lookup(const uint8_t counts[16], const uint8_t *symbols, const char *bits);
/* Image.cpp:285-302 bitStringtoValue (T.81 EXTEND). v = the n raw bits. */
int kpo_extend(int v, int n);
/* Transform.cpp:5-27 */
void kpo_zigzag_to_rc(int zz, int *row, int *col);
/* MCU.cpp:172-216 computeIDCT for one 8x8 block; F row-major ints; out row-major floats. */
void kpo_idct8x8(const int F[64], float out[64]);
/* MCU.cpp:228 performLevelShift on one value */
int kpo_level_shift(float v);
/* MCU.cpp:255-265 convertYCbCrToRGB for one pixel */
void kpo_ycbcr_to_rgb(int y, int cb, int cr, int rgb[3]);
/* Dequantise + de-zigzag + IDCT + level shift of one block of zig-zag coefficients
 * (MCU.cpp:110-120,172-245).  samples row-major, not clamped. */
void kpo_block_to_samples(const int16_t zz[64], const uint16_t qt[64], int samples[64]);

#ifdef __cplusplus
}
#endif
#endif
