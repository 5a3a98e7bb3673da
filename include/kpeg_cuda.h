/*
 * kpeg_cuda.h -- C ABI of the B200 (sm_100a) baseline-JPEG decode hot path.
 *
 * libKPEG has no plugin / FFI interface of its own: its hot path is reachable only through the
 * C++ class kpeg::JPEGDecoder (reference include/Decoder.hpp:26-128) and the `kpeg` CLI
 * (reference main.cpp:54-79).  This header is the thin boundary the drop-in JPEGDecoder in
 * libkpeg_b200/host/ calls; every entry point names the reference member(s) whose work it
 * replaces.  Plain pointers and sizes only, no C++ or torch types.  All file:line citations are
 * relative to /root/reference.
 *
 * There is NO CPU fallback behind this interface: every decode call runs the CUDA kernels in
 * libkpeg_b200/csrc/ and fails with KPEG_ERR_CUDA when no device is usable.
 */
#ifndef KPEG_CUDA_H
#define KPEG_CUDA_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ---- return codes ------------------------------------------------------------------------ */
#define KPEG_OK 0
#define KPEG_ERR_FORMAT (-1)      /* malformed container      -> JPEGDecoder::ResultCode::ERROR     (Decoder.cpp:126-132) */
#define KPEG_ERR_UNSUPPORTED (-2) /* SOF1/SOF2/subsampling... -> JPEGDecoder::ResultCode::TERMINATE (Decoder.cpp:65-66,351-356) */
#define KPEG_ERR_STREAM (-3)      /* corrupt entropy-coded data (the reference hangs / reads out of bounds, Decoder.cpp:706-803) */
#define KPEG_ERR_NOMEM (-4)
#define KPEG_ERR_CUDA (-5)        /* no device / CUDA runtime failure (see kpeg_cuda_last_error) */
#define KPEG_ERR_ARG (-6)
#define KPEG_ERR_NOT_CONVERGED (-7) /* internal: speculative decode needed more fix-up rounds (retried automatically) */

/* ---- plan flags -------------------------------------------------------------------------- */
/* Reproduce MCU.cpp:97-104: a block whose DC *difference* is 0 loses its AC coefficients
 * (SURVEY F1).  ON for bit-exact parity with the reference; OFF for ITU-T T.81 behaviour. */
#define KPEG_FLAG_REF_PARITY 1u

/* One Huffman table as it appears in a DHT segment (Types.hpp:116, Decoder.cpp:366-459). */
typedef struct kpeg_huff_spec {
    uint8_t counts[16];   /* number of codes of length 1..16 */
    uint8_t symbols[256]; /* symbols in order of increasing code length */
} kpeg_huff_spec;

/*
 * Everything the kernels need to know about one image, produced by the host-side container
 * parse (replaces the state JPEGDecoder accumulates in parseQuantizationTable / parseSOF0Segment /
 * parseHuffmanTable / parseSOSSegment, Decoder.cpp:230-530).  POD; caller-owned.
 */
typedef struct kpeg_plan {
    uint16_t width, height;     /* SOF0 X, Y (Decoder.cpp:301-364) */
    uint8_t ncomp;              /* 1 (T.81 extension) or 3; sampling is always 1x1 */
    uint8_t comp_tq[3];         /* quantisation-table id per component */
    uint8_t comp_td[3];         /* DC Huffman table id per component   */
    uint8_t comp_ta[3];         /* AC Huffman table id per component   */
    uint16_t restart_interval;  /* DRI value in MCUs, 0 = none (T.81 extension) */
    uint32_t flags;             /* KPEG_FLAG_* */
    uint8_t qt_present[4];
    uint8_t ht_present[2][4];
    uint16_t qt[4][64];         /* zig-zag (file) order, as Decoder.cpp:230-299 keeps them */
    kpeg_huff_spec ht[2][4];    /* [class 0=DC,1=AC][id] */
} kpeg_plan;

/* Stages timed with CUDA events on the decode stream while profiling is on. */
enum {
    KPEG_T_H2D = 0,        /* host -> device copy of the scan (host-pointer entry points only)            */
    KPEG_T_MEMSET,         /* zero-fill of bookkeeping, bit stream tail and coefficient buffer            */
    KPEG_T_UNSTUFF,        /* K0: unstuff_count + unstuff_scan + unstuff_write                            */
    KPEG_T_ENTROPY_COLD,   /* K1: speculative cold decode of every subsequence                            */
    KPEG_T_ENTROPY_RELAY,  /* K1: first relay round (every subsequence, emits symbol records)             */
    KPEG_T_ENTROPY_SCAN,   /* K1: segmented scan of slot counts and DC sums over the subsequences          */
    KPEG_T_ENTROPY_WRITE,  /* K2: record expansion + DC prediction -> coefficient tiles (or the Huffman final pass) */
    KPEG_T_DC_SCAN,        /* fallback only: DC prediction + tile conversion behind the Huffman final pass  */
    KPEG_T_IDCT,           /* K3: fused dequant + IDCT + colour + store, and the exact re-evaluation pass  */
    KPEG_T_D2H,            /* device -> host copy of the pixels (host-pointer entry points only)          */
    KPEG_T_RELAY_SPARSE,   /* K1: later relay rounds (work list only) up to the fixed point               */
    KPEG_T_IDCT_PATCH,     /* (not used: the exact pass is timed with KPEG_T_IDCT)                         */
    KPEG_T_COUNT
};

/* Per-decode statistics.  Stage times are filled only while profiling is enabled
 * (kpeg_cuda_set_profiling); they are CUDA-event times on the decode stream, in milliseconds. */
typedef struct kpeg_stats {
    uint32_t width, height, ncomp;
    uint64_t scan_bytes;        /* stuffed entropy-coded bytes handed in                  */
    uint64_t unstuffed_bytes;   /* after FF00 / RSTn removal                              */
    uint32_t segments;          /* restart intervals found (1 when there is no DRI)       */
    uint32_t subsequences;      /* speculative decode units                               */
    uint32_t sync_rounds;       /* fix-up rounds that ran until the relay reached a fixed point */
    uint32_t exact_samples;     /* IDCT samples re-evaluated on the exact (reference-order) path */
    float ms[KPEG_T_COUNT];     /* per-stage CUDA-event time, indexed by KPEG_T_* */
    float ms_total;             /* first to last event of the call */
    uint32_t kernel_launches;   /* kernels launched for this decode */
} kpeg_stats;

typedef struct kpeg_ctx kpeg_ctx;

/* ---- host-side container parse (no CUDA) --------------------------------------------------
 * Replaces JPEGDecoder::decodeImageFile's marker loop + parseSegmentInfo + the parse* members
 * (Decoder.cpp:53-75,90-152,164-530,579-619).  T.81-correct superset of what the reference
 * accepts: unknown APPn/COM are skipped by length, DRI and Nf==1 are honoured.  On success
 * [*scan_off, *scan_off + *scan_len) is the entropy-coded segment (what scanImageData,
 * Decoder.cpp:532-577, would read: everything after the SOS header up to the EOI marker). */
int kpeg_parse_jfif(const uint8_t *file, size_t len, kpeg_plan *plan, size_t *scan_off, size_t *scan_len);

/* A frame with one scan PER component (T.81 A.2.3; the reference reads such SOS headers, Decoder.cpp:461-530, and then
 * decodes the first scan as if it were interleaved).  One kpeg_scan per SOS: `plan` is everything needed to decode THAT
 * scan on its own -- ncomp = its number of components (1 for a non-interleaved scan, with that component's selectors in
 * slot 0), the tables and restart interval in force at its SOS (they may be redefined between scans, T.81 B.2.3) --
 * comp[] the frame component behind each slot, [off, off + len) its entropy-coded segment inside the file.
 * kpeg_parse_jfif_scans accepts what kpeg_parse_jfif accepts (then *nscans == 1 and scans[0].plan == *frame) plus
 * three-component frames coded as three single-component scans in any order; *frame then holds the dimensions and,
 * in slots 0..2, the quantisers of components 0..2 (frame->restart_interval is 0: it is a per-scan property). */
typedef struct kpeg_scan {
    kpeg_plan plan;
    uint8_t comp[3];
    size_t off, len;
} kpeg_scan;
#define KPEG_MAX_SCANS 3
int kpeg_parse_jfif_scans(const uint8_t *file, size_t len, kpeg_plan *frame, kpeg_scan *scans, int max_scans, int *nscans);

/* ---- context ------------------------------------------------------------------------------ */
/* One context per (thread, device): owns eight lanes (a CUDA stream with its own device scratch and pinned
 * bookkeeping each), all grown on demand and reused across decodes.  Not thread-safe; use one context per thread. */
int kpeg_cuda_create(int device, kpeg_ctx **out);
void kpeg_cuda_destroy(kpeg_ctx *ctx);
const char *kpeg_cuda_last_error(const kpeg_ctx *ctx); /* never NULL */
int kpeg_cuda_device_count(void);
int kpeg_cuda_set_profiling(kpeg_ctx *ctx, int on);
/* Tuning knobs of the speculative entropy decode (0 = leave unchanged): bits per subsequence (a power of two, 64..1024,
 * or $KPEG_SUB_BITS; -1 = back to the default: chosen per job, 1024 for streams with long restart segments AND long blocks --
 * large high-quality images without restart markers -- and 512 otherwise) and the number of relay rounds issued up front when the cooperative
 * loop is unavailable (>= 2; default 8 or $KPEG_RELAY_ROUNDS; more are added automatically when needed). */
int kpeg_cuda_set_tuning(kpeg_ctx *ctx, int sub_bits, int relay_rounds);
/* The CUDA stream (cudaStream_t) all of this context's work is issued on. */
void *kpeg_cuda_stream(kpeg_ctx *ctx);

/* A process-wide pool of idle contexts: acquire hands out one for `device` (creating it when the pool has none),
 * release puts it back with its scratch memory intact.  What kpeg::JPEGDecoder uses, so that decoding file after
 * file (reference main.cpp:54-79 constructs a decoder per file) does not pay context creation every time.
 * Thread-safe.  kpeg_cuda_pool_clear destroys the idle contexts. */
int kpeg_cuda_acquire(int device, kpeg_ctx **out);
void kpeg_cuda_release(int device, kpeg_ctx *ctx);
void kpeg_cuda_pool_clear(void);

/* Pinned host memory for scan / pixel buffers (plain malloc'd memory also works, slower); pinned for every device.
 * kpeg_cuda_host_register pins memory the caller already owns (a frame shared between processes, say). */
void *kpeg_cuda_host_alloc(size_t bytes);
void kpeg_cuda_host_free(void *p);
int kpeg_cuda_host_register(void *p, size_t bytes);
void kpeg_cuda_host_unregister(void *p);
/* Device memory helpers for callers that keep data resident (bench, pipelines). */
void *kpeg_cuda_device_alloc(kpeg_ctx *ctx, size_t bytes);
void kpeg_cuda_device_free(kpeg_ctx *ctx, void *p);
int kpeg_cuda_memcpy_h2d(kpeg_ctx *ctx, void *dst, const void *src, size_t bytes);
int kpeg_cuda_memcpy_d2h(kpeg_ctx *ctx, void *dst, const void *src, size_t bytes);

/* ---- the hot path ------------------------------------------------------------------------- */
/*
 * Decode one entropy-coded segment to pixels.  Replaces JPEGDecoder::byteStuffScanData,
 * decodeScanData (Decoder.cpp:621-855), MCU::constructMCU / computeIDCT / performLevelShift /
 * convertYCbCrToRGB (MCU.cpp:64-279) and Image::createImageFromMCUs (Image.cpp:20-86).
 *
 *  scan, scan_len : HOST pointer to the stuffed entropy-coded bytes (may contain RSTn markers).
 *  pixels_out     : HOST buffer of width*height*ncomp bytes; interleaved R,G,B rows top-down for
 *                   ncomp==3 (the payload Image::dumpRawData writes, Image.cpp:129-135), one gray
 *                   byte per pixel for ncomp==1.
 * Synchronous: returns after the pixels are in pixels_out.  An image of 64 MB of pixels or more whose restart interval is
 * a whole number of MCU rows is cut into bands that run on the context's lanes with their copies overlapping
 * ($KPEG_BANDS, default 4; 1 = one job).
 */
int kpeg_cuda_decode(kpeg_ctx *ctx, const kpeg_plan *plan, const uint8_t *scan, size_t scan_len,
                     uint8_t *pixels_out, kpeg_stats *stats);

/* Same, but `d_scan` and `d_pixels_out` are DEVICE pointers on the context's device; work is
 * enqueued on the context's stream and the call returns after the stream has drained and the
 * device-side status word has been checked. */
int kpeg_cuda_decode_device(kpeg_ctx *ctx, const kpeg_plan *plan, const uint8_t *d_scan, size_t scan_len,
                            uint8_t *d_pixels_out, kpeg_stats *stats);

/*
 * Batch of n images that share ONE plan (same dimensions, tables and restart interval -- the
 * case of BASELINE.json configs 4 and of repeated frames).  The scans are decoded as one
 * concatenated stream whose image boundaries are treated like restart boundaries, so the whole
 * batch costs one kernel sequence.  Host-pointer variant copies in/out through pinned staging.
 */
int kpeg_cuda_decode_batch(kpeg_ctx *ctx, const kpeg_plan *plan, int n, const uint8_t *const *scans,
                           const size_t *scan_lens, uint8_t *const *pixels_out, kpeg_stats *stats);
/* Device-resident batch: d_scans is ONE device buffer holding the n scans back to back,
 * scan_offsets[n+1] (host array) delimit them; d_pixels_out receives n images back to back. */
int kpeg_cuda_decode_batch_device(kpeg_ctx *ctx, const kpeg_plan *plan, int n, const uint8_t *d_scans,
                                  const uint64_t *scan_offsets, uint8_t *d_pixels_out, kpeg_stats *stats);
/* "Packed" batch stream = the n stuffed scans back to back, each FOLLOWED by one 2-byte RSTn marker
 * (FF D0+(i&7)).  kpeg_batch_pack builds it on the host; the *_packed_device entry decodes such a
 * stream that is already resident in device memory (no copies at all inside the call). */
size_t kpeg_batch_packed_size(int n, const size_t *scan_lens);
int kpeg_batch_pack(int n, const uint8_t *const *scans, const size_t *scan_lens, uint8_t *dst, size_t cap);
int kpeg_cuda_decode_batch_packed_device(kpeg_ctx *ctx, const kpeg_plan *plan, int n, const uint8_t *d_packed,
                                         size_t packed_len, uint8_t *d_pixels_out, kpeg_stats *stats);
/* Same, given the offsets of the n scans inside the packed stream (packed_offsets[n] = its length): the
 * batch is cut into up to four parts that run as concurrent jobs on the context's lanes (streams). */
int kpeg_cuda_decode_batch_packed_device_split(kpeg_ctx *ctx, const kpeg_plan *plan, int n, const uint8_t *d_packed,
                                               const uint64_t *packed_offsets, uint8_t *d_pixels_out, kpeg_stats *stats);

/* Deferred form: enqueue the batch and return at once; kpeg_cuda_wait completes everything submitted since the
 * last wait and returns the first failure (stats: accumulated over those batches).  Consecutive submissions
 * overlap on the device.  d_packed / d_pixels_out must stay valid until kpeg_cuda_wait returns.  Any other
 * decode call on the context completes pending submissions first. */
int kpeg_cuda_submit_batch_packed_device(kpeg_ctx *ctx, const kpeg_plan *plan, int n, const uint8_t *d_packed,
                                         const uint64_t *packed_offsets, uint8_t *d_pixels_out);
int kpeg_cuda_wait(kpeg_ctx *ctx, kpeg_stats *stats);
/* Deferred form of kpeg_cuda_decode_batch (host buffers, ideally pinned, in and out): returns once the copies
 * and kernels are enqueued; the pixels are valid after kpeg_cuda_wait.  scans[i] / pixels_out[i] must stay
 * valid and untouched until then (the pointer arrays themselves are read before the call returns). */
int kpeg_cuda_submit_batch(kpeg_ctx *ctx, const kpeg_plan *plan, int n, const uint8_t *const *scans,
                           const size_t *scan_lens, uint8_t *const *pixels_out);

/* The scans kpeg_parse_jfif_scans found, decoded to one frame (HOST buffers in and out, synchronous).  One interleaved
 * scan is kpeg_cuda_decode; three single-component scans are entropy-decoded one after another, each with its own
 * tables, into per-component coefficient tiles, which one small kernel interleaves into the MCU order K3 consumes --
 * from there on (dequantisation, IDCT, colour, store: MCU.cpp:110-279, Image.cpp:51-70) the path is the same.
 * frame->flags selects the parity mode for every scan. */
int kpeg_cuda_decode_scans(kpeg_ctx *ctx, const kpeg_plan *frame, const kpeg_scan *scans, int nscans, const uint8_t *file,
                           size_t file_len, uint8_t *pixels_out, kpeg_stats *stats);

/* Whole-file convenience used by the JPEGDecoder drop-in: parse (kpeg_parse_jfif_scans) + decode (kpeg_cuda_decode_scans).
 * pixels_out must hold width*height*ncomp bytes (query with kpeg_parse_jfif_scans first) -- cap is checked. */
int kpeg_cuda_decode_file(kpeg_ctx *ctx, const uint8_t *file, size_t len, uint32_t flags, uint8_t *pixels_out,
                          size_t cap, kpeg_plan *plan_out, kpeg_stats *stats);

/* n whole files of ANY mix of dimensions, tables and codings (what "decode this directory" needs; the reference
 * constructs one JPEGDecoder per file, main.cpp:54-79).  Parsed on the host; files whose plans are identical are decoded
 * together as one batch (kpeg_cuda_submit_batch: one kernel sequence per group), files coded one scan per component go
 * through kpeg_cuda_decode_scans.  pixels_out[i] must hold caps[i] >= width*height*ncomp bytes of file i (query with
 * kpeg_parse_jfif_scans).  results[i] (may be NULL) receives file i's own return code, plans_out[i] (may be NULL) its
 * frame plan; a bad file does not stop the others.  Returns KPEG_OK when every file decoded, else the first failure
 * (text: kpeg_cuda_last_error).  Host buffers (ideally pinned) in and out; synchronous. */
int kpeg_cuda_decode_files(kpeg_ctx *ctx, int n, const uint8_t *const *files, const size_t *lens, uint32_t flags,
                           uint8_t *const *pixels_out, const size_t *caps, kpeg_plan *plans_out, int *results, kpeg_stats *stats);

/* Parity hook: the quantised coefficients of the LAST decode on this context, as the reference
 * holds them transiently inside MCU::constructMCU (MCU.cpp:93-108): [block][64] int16, blocks
 * MCU-interleaved (Y,Cb,Cr per MCU), zig-zag order, DC prediction already integrated.
 * `cap` is in int16 elements.  (After a decode that ran as several jobs -- a large restart-marked image cut into bands,
 * a batch cut into chunks -- these are the coefficients of the job that finished last.) */
int kpeg_cuda_read_coefficients(kpeg_ctx *ctx, int16_t *out, size_t cap);

/* Host-only: cut one restart-marked scan into `parts` bands of whole MCU rows; band b is the byte
 * range scan[out_begin[b], out_end[b]) and decodes as an image of the same width and tables whose
 * MCU rows are [out_row[b], out_row[b+1]).  This is how restart-interval tiles of a very large image
 * are spread over several GPUs (one context per GPU, no device-to-device traffic; the reference has
 * no tiling).  Needs a restart interval that is a whole number of MCU rows or divides one.
 * out_begin / out_end have `parts` entries, out_row has parts + 1. */
int kpeg_split_restart_bands(const uint8_t *scan, size_t len, const kpeg_plan *plan, int parts, uint64_t *out_begin,
                             uint64_t *out_end, uint32_t *out_row);

/* Host-only: the same kind of bands, cut at BYTE positions instead of given rows: band b begins at the first RSTn marker
 * at or after b/parts of the scan, and its rows follow from the number of restart markers it holds (counted by one host
 * thread per band) -- no single walk over all markers of a large scan; the bands' heights depend on the data.  Needs a
 * restart interval of whole MCU rows and at least two intervals per band (else KPEG_ERR_UNSUPPORTED: use
 * kpeg_split_restart_bands).  Same outputs as kpeg_split_restart_bands. */
int kpeg_split_restart_bands_by_bytes(const uint8_t *scan, size_t len, const kpeg_plan *plan, int parts, uint64_t *out_begin,
                                      uint64_t *out_end, uint32_t *out_row);

/* ---- one image over several GPUs (kpeg_tiled.cpp) ---------------------------------------------------------
 * The restart-interval tiles of one very large image, one band of whole MCU rows per listed device (a device may be
 * listed more than once), each band decoded by its own GPU straight into its rows of the caller's frame
 * pixels_out (ideally pinned: kpeg_cuda_host_alloc) -- the frame Image::createImageFromMCUs assembles
 * (Image.cpp:51-70), gathered on the host with no device-to-device traffic.  One host thread and one pooled context
 * per device.  An image whose restart interval does not line up with MCU rows (or has none) does not shard: it is
 * decoded whole on devices[0].  Errors: the return code; text in kpeg_tiled_last_error (per calling thread). */
int kpeg_cuda_decode_tiled(const int *devices, int ndev, const kpeg_plan *plan, const uint8_t *scan, size_t scan_len,
                           uint8_t *pixels_out, kpeg_stats *stats);
int kpeg_cuda_decode_file_tiled(const int *devices, int ndev, const uint8_t *file, size_t len, uint32_t flags,
                                uint8_t *pixels_out, size_t cap, kpeg_plan *plan_out, kpeg_stats *stats);
/* Same, the frame assembled in DEVICE memory of dst_device (d_frame: height*width*ncomp bytes there): every band
 * travels by one peer copy over NVLink (staged by the driver where there is no peer access), so a consumer on
 * dst_device never pays PCIe. */
int kpeg_cuda_decode_tiled_device(const int *devices, int ndev, const kpeg_plan *plan, const uint8_t *scan, size_t scan_len,
                                  int dst_device, uint8_t *d_frame, kpeg_stats *stats);
const char *kpeg_tiled_last_error(void);
/* One band of such a decode: host scan in, pixels to device memory d_dst of dst_device (this context's own device:
 * written in place by the kernels). */
int kpeg_cuda_decode_to_peer(kpeg_ctx *ctx, const kpeg_plan *plan, const uint8_t *scan, size_t scan_len, int dst_device,
                             uint8_t *d_dst, kpeg_stats *stats);

/* ---- output formats written on the GPU ------------------------------------------------------------------------
 * The file Image::dumpRawData writes (Image.cpp:108-140: P6 header, then R,G,B rows) assembled in device memory:
 * d_out[*ppm_off, *ppm_off + *ppm_len) holds it (ppm_off < 16 keeps the payload aligned for the kernels' vector
 * stores; d_out must be 16-byte aligned, cap >= 176 + 3*width*height).  One-component images come out as R = G = B. */
int kpeg_cuda_decode_ppm_device(kpeg_ctx *ctx, const kpeg_plan *plan, const uint8_t *d_scan, size_t scan_len, uint8_t *d_out,
                                size_t cap, size_t *ppm_off, size_t *ppm_len, kpeg_stats *stats);
/* Interleaved R,G,B (the decode output) -> three planes [3][npixels], both in device memory. */
int kpeg_cuda_interleaved_to_planar(kpeg_ctx *ctx, const uint8_t *d_rgb, uint8_t *d_planes, size_t npixels);

/* Exact bytes of the PPM header Image::dumpRawData writes (Image.cpp:124-127); returns length. */
int kpeg_ppm_header(int width, int height, char *buf, size_t cap);

#ifdef __cplusplus
}
#endif
#endif
