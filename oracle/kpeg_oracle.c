/*
 * kpeg_oracle.c -- CPU restatement of libKPEG's baseline-JPEG decode hot path.
 *
 * TEST INFRASTRUCTURE ONLY (see kpeg_oracle.h).  Parity status: PINNED against the compiled
 * reference (byte-exact PPM on lena.jpg and synthetic twins) and the reference's own KATs.
 *
 * This is a restatement of WHAT the reference computes, written from the behavioural
 * specification in SURVEY.md Appendix A, not a transcription of its code: the reference
 * walks a '0'/'1' std::string and a shared_ptr tree; this file uses a byte-wise bit reader and
 * canonical-code tables that produce the same symbols.  The only place where the reference's
 * exact instruction sequence matters is the floating-point IDCT / level shift / colour
 * conversion, whose operation order and operand types are reproduced literally because the
 * rounding of x.5 ties depends on them.
 *
 * Build: oracle/Makefile (gcc -O2 -ffp-contract=off, x86-64 SSE2, no fast-math, like the
 * reference's plain -O2 in CMakeLists.txt:18).
 *
 * All file:line citations are relative to /root/reference.
 */
#define _GNU_SOURCE
#include "kpeg_oracle.h"

#include <math.h>
#include <pthread.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

/* ------------------------------------------------------------------------------------------
 * Zig-zag map.  Transform.cpp:5-27 (zzOrderToMatIndices): index i of the zig-zag scan ->
 * (row, col).  Generated here by walking the anti-diagonals instead of a literal table.
 * ---------------------------------------------------------------------------------------- */
static int g_zz_row[64], g_zz_col[64];
static pthread_once_t g_zz_once = PTHREAD_ONCE_INIT;

static void zz_init(void)
{
    int i = 0;
    for (int d = 0; d < 15; ++d) {
        /* even diagonals run bottom-left -> top-right, odd ones top-right -> bottom-left */
        int lo = d < 8 ? 0 : d - 7, hi = d < 8 ? d : 7;
        for (int k = lo; k <= hi; ++k) {
            int r = (d & 1) ? k : (hi + lo - k);
            g_zz_row[i] = r;
            g_zz_col[i] = d - r;
            ++i;
        }
    }
}

void kpo_zigzag_to_rc(int zz, int *row, int *col)
{
    pthread_once(&g_zz_once, zz_init);
    *row = g_zz_row[zz & 63];
    *col = g_zz_col[zz & 63];
}

/* ------------------------------------------------------------------------------------------
 * Canonical Huffman codes.  HuffmanTree.cpp:106-157 builds the tree level by level, shortest
 * codes first, left to right -- exactly T.81 Annex C: code=0; for each length, each symbol
 * gets `code++`; then code <<= 1.
 * ---------------------------------------------------------------------------------------- */
typedef struct {
    uint8_t present;
    uint8_t counts[16];
    uint8_t symbols[256];
    int nsym;
    /* per length L (1..16): first code value, index of first symbol, count */
    int32_t first_code[17];
    int32_t first_idx[17];
} huff_t;

static int huff_prepare(huff_t *h)
{
    int code = 0, idx = 0;
    for (int L = 1; L <= 16; ++L) {
        h->first_code[L] = code;
        h->first_idx[L] = idx;
        code += h->counts[L - 1];
        idx += h->counts[L - 1];
        if (code > (1 << L))
            return -1; /* over-subscribed */
        code <<= 1;
    }
    h->nsym = idx;
    return idx <= 256 ? 0 : -1;
}

int kpo_huff_codes(const uint8_t counts[16], const uint8_t *symbols, uint16_t *codes, uint8_t *lens)
{
    (void)symbols;
    int code = 0, idx = 0;
    for (int L = 1; L <= 16; ++L) {
        for (int k = 0; k < counts[L - 1]; ++k) {
            codes[idx] = (uint16_t)code++;
            lens[idx] = (uint8_t)L;
            ++idx;
        }
        code <<= 1;
    }
    return idx;
}

int kpo_huff_lookup(const uint8_t counts[16], const uint8_t *symbols, const char *bits)
{
    /* HuffmanTree.cpp:164-193: a bit string is "contained" iff it is exactly a leaf's code. */
    size_t n = strlen(bits);
    if (n < 1 || n > 16)
        return -1;
    int v = 0;
    for (size_t i = 0; i < n; ++i)
        v = (v << 1) | (bits[i] == '1');
    int code = 0, idx = 0;
    for (int L = 1; L <= 16; ++L) {
        if ((size_t)L == n) {
            if (v >= code && v < code + counts[L - 1])
                return symbols[idx + (v - code)];
            return -1;
        }
        code = (code + counts[L - 1]) << 1;
        idx += counts[L - 1];
    }
    return -1;
}

/* Image.cpp:285-302 bitStringtoValue: leading '1' -> plain binary; leading '0' -> minus the
 * bitwise complement; empty -> 0.  (T.81 F.2.2.1 EXTEND.) */
int kpo_extend(int v, int n)
{
    if (n == 0)
        return 0;
    if (v >> (n - 1))
        return v;
    return v - ((1 << n) - 1);
}

/* ------------------------------------------------------------------------------------------
 * Container parse (cold path).  Decoder.cpp:53-75,164-530,579-619 on the subset the reference
 * accepts; T.81 B.2 beyond it (DRI, Nf==1, APPn/COM skipped by length, SOF/SOS selectors
 * honoured -- the reference hard-wires Y->0, Cb/Cr->1 (Decoder.cpp:704, MCU.cpp:110), which
 * coincides with the selectors of every file it decodes correctly).
 * ---------------------------------------------------------------------------------------- */
typedef struct {
    int width, height, ncomp;
    int comp_id[3], comp_tq[3], comp_td[3], comp_ta[3];
    int restart_interval;
    uint16_t qt[4][64];
    int qt_present[4];
    huff_t ht[2][4];
    size_t scan_off, scan_len; /* the first scan (all of it for the single interleaved scan the reference decodes) */
    /* Every scan of the frame, each with the tables and restart interval in force at its SOS (T.81 B.2.3: tables may
     * be redefined between scans).  The reference reads Ns = 1..4 (Decoder.cpp:461-530) but then decodes the first scan
     * as if it were interleaved; here a frame is either ONE interleaved scan of all components or one scan PER
     * component (T.81 A.2.3: the MCU of a non-interleaved scan is one block, blocks in raster order). */
    int nscans;
    struct scan_t {
        int ns, comp[3], td[3], ta[3];
        int restart_interval;
        huff_t ht[2][4];
        size_t off, len;
    } scans[3];
} plan_t;

static int rd16(const uint8_t *p) { return (p[0] << 8) | p[1]; }

static int parse_container(const uint8_t *f, size_t n, plan_t *pl)
{
    memset(pl, 0, sizeof *pl);
    if (n < 4 || f[0] != 0xFF || f[1] != 0xD8)
        return KPO_ERR_FORMAT;
    size_t i = 2;
    int have_sof = 0;
    for (;;) {
        if (i + 4 > n)
            return KPO_ERR_FORMAT;
        if (f[i] != 0xFF)
            return KPO_ERR_FORMAT; /* Decoder.cpp:126-132 "[ FATAL ] Invalid JFIF file" */
        while (i < n && f[i] == 0xFF && i + 1 < n && f[i + 1] == 0xFF)
            ++i; /* fill bytes */
        int m = f[i + 1];
        i += 2;
        if (m == 0xD8 || (m >= 0xD0 && m <= 0xD7) || m == 0x01)
            continue;
        if (m == 0xD9)
            return KPO_ERR_FORMAT; /* EOI before SOS */
        if (i + 2 > n)
            return KPO_ERR_FORMAT;
        int L = rd16(f + i);
        if (L < 2 || i + (size_t)L > n)
            return KPO_ERR_FORMAT;
        const uint8_t *p = f + i + 2;
        int plen = L - 2;
        switch (m) {
        case 0xC0: { /* SOF0, Decoder.cpp:301-364 */
            if (plen < 6)
                return KPO_ERR_FORMAT;
            if (p[0] != 8)
                return KPO_ERR_UNSUPPORTED;
            pl->height = rd16(p + 1);
            pl->width = rd16(p + 3);
            pl->ncomp = p[5];
            if (pl->ncomp != 1 && pl->ncomp != 3)
                return KPO_ERR_UNSUPPORTED;
            if (plen < 6 + 3 * pl->ncomp)
                return KPO_ERR_FORMAT;
            for (int c = 0; c < pl->ncomp; ++c) {
                pl->comp_id[c] = p[6 + 3 * c];
                if (pl->ncomp == 3 && p[7 + 3 * c] != 0x11)
                    return KPO_ERR_UNSUPPORTED; /* "Chroma subsampling not yet supported!" :351-356 */
                pl->comp_tq[c] = p[8 + 3 * c] & 3;
            }
            if (pl->width == 0 || pl->height == 0)
                return KPO_ERR_UNSUPPORTED;
            have_sof = 1;
            break;
        }
        case 0xC1: case 0xC2: case 0xC3: case 0xC5: case 0xC6: case 0xC7:
        case 0xC9: case 0xCA: case 0xCB: case 0xCD: case 0xCE: case 0xCF:
            return KPO_ERR_UNSUPPORTED; /* Decoder.cpp:65-66 */
        case 0xDB: { /* DQT, Decoder.cpp:230-299: (len-2)/65 tables, 8-bit entries, zig-zag order */
            int k = 0;
            while (k < plen) {
                int pq = p[k] >> 4, tq = p[k] & 15;
                if (pq != 0 || tq > 3)
                    return KPO_ERR_UNSUPPORTED;
                if (k + 65 > plen)
                    return KPO_ERR_FORMAT;
                for (int q = 0; q < 64; ++q)
                    pl->qt[tq][q] = p[k + 1 + q];
                pl->qt_present[tq] = 1;
                k += 65;
            }
            break;
        }
        case 0xC4: { /* DHT, Decoder.cpp:366-459 */
            int k = 0;
            while (k < plen) {
                if (k + 17 > plen)
                    return KPO_ERR_FORMAT;
                int tc = (p[k] >> 4) & 1, th = p[k] & 15;
                if ((p[k] >> 5) || th > 3)
                    return KPO_ERR_FORMAT;
                huff_t *h = &pl->ht[tc][th];
                memset(h, 0, sizeof *h);
                int tot = 0;
                for (int q = 0; q < 16; ++q) {
                    h->counts[q] = p[k + 1 + q];
                    tot += h->counts[q];
                }
                if (tot > 256 || k + 17 + tot > plen)
                    return KPO_ERR_FORMAT;
                memcpy(h->symbols, p + k + 17, (size_t)tot);
                if (huff_prepare(h))
                    return KPO_ERR_FORMAT;
                h->present = 1;
                k += 17 + tot;
            }
            break;
        }
        case 0xDD: /* DRI -- unhandled by the reference; T.81 B.2.4.4 */
            if (plen < 2)
                return KPO_ERR_FORMAT;
            pl->restart_interval = rd16(p);
            break;
        case 0xDA: { /* SOS, Decoder.cpp:461-530 */
            if (!have_sof || plen < 1)
                return KPO_ERR_FORMAT;
            int ns = p[0];
            if ((ns != pl->ncomp && ns != 1) || plen < 1 + 2 * ns + 3 || pl->nscans >= 3)
                return KPO_ERR_UNSUPPORTED; /* two-component scans, more scans than components */
            struct scan_t *sc = &pl->scans[pl->nscans];
            sc->ns = ns;
            for (int s = 0; s < ns; ++s) {
                int cid = p[1 + 2 * s], c;
                for (c = 0; c < pl->ncomp; ++c)
                    if (pl->comp_id[c] == cid)
                        break;
                if (c == pl->ncomp)
                    return KPO_ERR_FORMAT;
                if (ns > 1 && c != s)
                    return KPO_ERR_UNSUPPORTED; /* interleaved: scan order must equal frame order */
                for (int k = 0; k < pl->nscans; ++k)
                    for (int j = 0; j < pl->scans[k].ns; ++j)
                        if (pl->scans[k].comp[j] == c)
                            return KPO_ERR_UNSUPPORTED; /* a component coded twice (not baseline sequential) */
                sc->comp[s] = c;
                sc->td[s] = p[2 + 2 * s] >> 4;
                sc->ta[s] = p[2 + 2 * s] & 15;
                if (sc->td[s] > 3 || sc->ta[s] > 3)
                    return KPO_ERR_FORMAT;
                if (!pl->qt_present[pl->comp_tq[c]])
                    return KPO_ERR_FORMAT;
                if (!pl->ht[0][sc->td[s]].present || !pl->ht[1][sc->ta[s]].present)
                    return KPO_ERR_FORMAT;
                pl->comp_td[c] = sc->td[s];
                pl->comp_ta[c] = sc->ta[s];
            }
            sc->restart_interval = pl->restart_interval;
            memcpy(sc->ht, pl->ht, sizeof sc->ht);
            /* entropy-coded segment: everything up to the first marker that is neither a
             * stuffed FF00, an RSTn nor an FF fill byte.  scanImageData (Decoder.cpp:532-577)
             * stops only at FFD9; identical on well-formed single-scan files. */
            size_t s0 = i + (size_t)L, e = s0;
            while (e + 1 < n) {
                if (f[e] == 0xFF) {
                    int b = f[e + 1];
                    if (b == 0x00 || (b >= 0xD0 && b <= 0xD7) || b == 0xFF) {
                        e += (b == 0xFF) ? 1 : 2;
                        continue;
                    }
                    break;
                }
                ++e;
            }
            if (e + 1 >= n)
                e = n; /* truncated file: take what is there */
            sc->off = s0;
            sc->len = e - s0;
            if (pl->nscans == 0) {
                pl->scan_off = s0;
                pl->scan_len = e - s0;
            }
            ++pl->nscans;
            int covered = 0;
            for (int k = 0; k < pl->nscans; ++k)
                covered += pl->scans[k].ns;
            if (covered == pl->ncomp)
                return KPO_OK; /* every component has its scan: whatever follows (EOI, trailing bytes) is not looked at */
            if (e >= n)
                return KPO_ERR_FORMAT; /* the file ends before the remaining components' scans */
            i = e; /* on the marker that ended the segment */
            continue;
        }
        default: /* APPn, COM, DNL...: skipped by length (the reference FATALs on most, F7) */
            break;
        }
        i += (size_t)L;
    }
}

/* ------------------------------------------------------------------------------------------
 * Entropy decode.  Decoder.cpp:532-577 (bytes -> bits), :621-653 (drop the 00 of FF00),
 * :655-855 (symbol loop), Image.cpp:285-302 (EXTEND), MCU.cpp:88-108 (RLE expansion with the
 * F1 quirk, DC prediction).
 * ---------------------------------------------------------------------------------------- */
typedef struct {
    const uint8_t *p;
    size_t n, pos;
    uint32_t acc;
    int nbits;
    int overrun;
} bitrd_t;

static int br_bit(bitrd_t *b)
{
    if (b->nbits == 0) {
        uint8_t v;
        if (b->pos >= b->n) {
            b->overrun = 1;
            v = 0xFF;
        } else {
            v = b->p[b->pos++];
            if (v == 0xFF && b->pos < b->n && b->p[b->pos] == 0x00)
                b->pos++; /* byte stuffing */
        }
        b->acc = v;
        b->nbits = 8;
    }
    b->nbits--;
    return (int)((b->acc >> b->nbits) & 1u);
}

static int br_bits(bitrd_t *b, int n)
{
    int v = 0;
    while (n--)
        v = (v << 1) | br_bit(b);
    return v;
}

static int huff_decode(bitrd_t *b, const huff_t *h)
{
    int code = 0;
    for (int L = 1; L <= 16; ++L) {
        code = (code << 1) | br_bit(b);
        int off = code - h->first_code[L];
        if (off >= 0 && off < h->counts[L - 1])
            return h->symbols[h->first_idx[L] + off];
    }
    return -1; /* the reference would keep appending bits forever (Decoder.cpp:706-748) */
}

static int decode_coefficients(const plan_t *pl, const struct scan_t *sc, const uint8_t *scan, size_t slen, uint32_t flags,
                               int16_t *coef)
{
    const int nc = pl->ncomp;
    const int mx = (pl->width + 7) / 8, my = (pl->height + 7) / 8;
    const long nmcu = (long)mx * my; /* 1x1 sampling: a non-interleaved scan has as many MCUs (blocks) as an interleaved one */
    bitrd_t br = {scan, slen, 0, 0, 0, 0};
    int pred[3] = {0, 0, 0}; /* MCU.cpp:53 DCDiff[]; per-image here */
    const int ri = sc->restart_interval;

    for (long m = 0; m < nmcu; ++m) {
        if (ri && m && (m % ri) == 0) {
            /* T.81 F.2.2.4 / E.2.4: byte-align, consume RSTn, reset predictors */
            br.nbits = 0;
            while (br.pos + 1 < br.n && !(br.p[br.pos] == 0xFF && br.p[br.pos + 1] >= 0xD0 &&
                                          br.p[br.pos + 1] <= 0xD7))
                br.pos++; /* tolerate stray bytes before the marker */
            if (br.pos + 1 >= br.n)
                return KPO_ERR_STREAM;
            br.pos += 2;
            pred[0] = pred[1] = pred[2] = 0;
        }
        for (int si = 0; si < sc->ns; ++si) {
            const int c = sc->comp[si];
            const huff_t *hd = &sc->ht[0][sc->td[si]], *ha = &sc->ht[1][sc->ta[si]];
            int16_t *zz = coef + ((size_t)m * nc + c) * 64;
            memset(zz, 0, 64 * sizeof(int16_t));

            /* DC: Decoder.cpp:706-748 */
            int s = huff_decode(&br, hd);
            if (s < 0)
                return KPO_ERR_STREAM;
            int cat = s & 15;
            int diff = kpo_extend(br_bits(&br, cat), cat);

            /* AC: Decoder.cpp:755-803.  `count` is checked BEFORE each symbol (:759) and may
             * overshoot 63 (a run that crosses the end of the block). */
            int count = 0, j = 0;
            int drop_ac = (flags & KPO_FLAG_REF_PARITY) && diff == 0; /* MCU.cpp:99-100 (F1) */
            while (count != 63) {
                s = huff_decode(&br, ha);
                if (s < 0)
                    return KPO_ERR_STREAM;
                if (s == 0x00)
                    break; /* EOB */
                int run = s >> 4;
                cat = s & 15;
                int val = kpo_extend(br_bits(&br, cat), cat);
                count += run + 1;
                j += run + 1;
                if (count > 63)
                    break; /* malformed: the reference would write past zzOrder[63] (MCU.cpp:103) */
                if (!drop_ac)
                    zz[j] = (int16_t)val; /* MCU.cpp:102-103 */
            }
            /* MCU.cpp:107-108 */
            pred[c] += diff;
            zz[0] = (int16_t)pred[c];
            if (br.overrun)
                return KPO_ERR_STREAM;
        }
    }
    return KPO_OK;
}

/* ------------------------------------------------------------------------------------------
 * Reconstruction.  MCU.cpp:110-120 (dequantise, de-zigzag), :172-216 (IDCT), :218-245 (level
 * shift), :247-279 (colour), Image.cpp:51-70 (placement).
 * ---------------------------------------------------------------------------------------- */
static double g_cos[8][8]; /* g_cos[x][u] = cos((2x+1) u pi / 16), the expression of MCU.cpp:193 */
static float g_cc[8][8];   /* (float)Cu * (float)Cv */
static pthread_once_t g_cos_once = PTHREAD_ONCE_INIT;

static void cos_init(void)
{
    for (int x = 0; x < 8; ++x)
        for (int u = 0; u < 8; ++u)
            g_cos[x][u] = cos((2 * x + 1) * u * M_PI / 16.0);
    for (int u = 0; u < 8; ++u)
        for (int v = 0; v < 8; ++v) {
            float Cu = u == 0 ? 1.0 / sqrt(2.0) : 1.0; /* MCU.cpp:190 */
            float Cv = v == 0 ? 1.0 / sqrt(2.0) : 1.0; /* MCU.cpp:191 */
            g_cc[u][v] = Cu * Cv;
        }
}

void kpo_idct8x8(const int F[64], float out[64])
{
    pthread_once(&g_cos_once, cos_init);
    /* MCU.cpp:178-199.  x indexes output ROWS here, y output COLUMNS (the reference's names
     * are swapped, its result is the ordinary IDCT).  Types: the running sum is float; each
     * term is (float*float*float-from-int) promoted to double, times two double cosines, and
     * the += rounds back to float every term. */
    for (int x = 0; x < 8; ++x) {
        for (int y = 0; y < 8; ++y) {
            float sum = 0.0f;
            for (int u = 0; u < 8; ++u) {
                for (int v = 0; v < 8; ++v) {
                    float t = g_cc[u][v] * (float)F[u * 8 + v];
                    double d = (double)t * g_cos[x][u] * g_cos[y][v];
                    sum = (float)((double)sum + d);
                }
            }
            out[x * 8 + y] = (float)(0.25 * (double)sum); /* MCU.cpp:198 */
        }
    }
}

int kpo_level_shift(float v)
{
    /* MCU.cpp:228: roundl() = half away from zero on the float value; no clamp. */
    return (int)(roundl((long double)v) + 128);
}

void kpo_ycbcr_to_rgb(int yi, int cbi, int cri, int rgb[3])
{
    /* MCU.cpp:255-265: float inputs (exact ints), double arithmetic, floor, clamp. */
    float Y = (float)yi, Cb = (float)cbi, Cr = (float)cri;
    int R = (int)floor(Y + 1.402 * (1.0 * Cr - 128.0));
    int G = (int)floor(Y - 0.344136 * (1.0 * Cb - 128.0) - 0.714136 * (1.0 * Cr - 128.0));
    int B = (int)floor(Y + 1.772 * (1.0 * Cb - 128.0));
    rgb[0] = R < 0 ? 0 : (R > 255 ? 255 : R);
    rgb[1] = G < 0 ? 0 : (G > 255 ? 255 : G);
    rgb[2] = B < 0 ? 0 : (B > 255 ? 255 : B);
}

void kpo_block_to_samples(const int16_t zz[64], const uint16_t qt[64], int samples[64])
{
    pthread_once(&g_zz_once, zz_init);
    int F[64];
    float o[64];
    for (int i = 0; i < 64; ++i)
        F[g_zz_row[i] * 8 + g_zz_col[i]] = (int)zz[i] * (int)qt[i]; /* MCU.cpp:110-120 */
    kpo_idct8x8(F, o);
    for (int i = 0; i < 64; ++i)
        samples[i] = kpo_level_shift(o[i]);
}

static int g_threads = 1;
void kpo_set_threads(int n) { g_threads = n < 1 ? 1 : (n > 256 ? 256 : n); }

typedef struct {
    const plan_t *pl;
    const int16_t *coef;
    uint8_t *pix;
    int mx, my;
    int row0, row1; /* MCU rows */
} recon_job_t;

static void *recon_worker(void *arg)
{
    recon_job_t *j = (recon_job_t *)arg;
    const plan_t *pl = j->pl;
    const int nc = pl->ncomp, W = pl->width, H = pl->height;
    int smp[3][64];
    for (int by = j->row0; by < j->row1; ++by) {
        for (int bx = 0; bx < j->mx; ++bx) {
            size_t m = (size_t)by * j->mx + bx; /* Image.cpp:51-70: raster MCU order */
            for (int c = 0; c < nc; ++c)
                kpo_block_to_samples(j->coef + (m * nc + c) * 64, pl->qt[pl->comp_tq[c]], smp[c]);
            for (int r = 0; r < 8; ++r) {
                int yy = by * 8 + r;
                if (yy >= H)
                    break; /* Image.cpp:72-83 crop */
                for (int q = 0; q < 8; ++q) {
                    int xx = bx * 8 + q;
                    if (xx >= W)
                        break;
                    if (nc == 3) {
                        int rgb[3];
                        kpo_ycbcr_to_rgb(smp[0][r * 8 + q], smp[1][r * 8 + q], smp[2][r * 8 + q], rgb);
                        uint8_t *o = j->pix + ((size_t)yy * W + xx) * 3;
                        o[0] = (uint8_t)rgb[0];
                        o[1] = (uint8_t)rgb[1];
                        o[2] = (uint8_t)rgb[2];
                    } else {
                        /* gray-as-YCbCr twin (Cb=Cr=128) gives R=G=B=clamp(Y): SURVEY A.8 */
                        int v = smp[0][r * 8 + q];
                        j->pix[(size_t)yy * W + xx] = (uint8_t)(v < 0 ? 0 : (v > 255 ? 255 : v));
                    }
                }
            }
        }
    }
    return NULL;
}

static void reconstruct(const plan_t *pl, const int16_t *coef, uint8_t *pix)
{
    pthread_once(&g_zz_once, zz_init);
    pthread_once(&g_cos_once, cos_init);
    const int mx = (pl->width + 7) / 8, my = (pl->height + 7) / 8;
    int nt = g_threads > my ? my : g_threads;
    recon_job_t jobs[256];
    pthread_t th[256];
    for (int t = 0; t < nt; ++t) {
        jobs[t] = (recon_job_t){pl, coef, pix, mx, my, (int)((long)my * t / nt), (int)((long)my * (t + 1) / nt)};
        if (t + 1 == nt)
            recon_worker(&jobs[t]);
        else
            pthread_create(&th[t], NULL, recon_worker, &jobs[t]);
    }
    for (int t = 0; t + 1 < nt; ++t)
        pthread_join(th[t], NULL);
}

/* ---------------------------------------------------------------------------------------- */

int kpo_decode(const uint8_t *file, size_t len, uint32_t flags, int want_pixels, kpo_image *img)
{
    plan_t *pl = (plan_t *)malloc(sizeof(plan_t));
    if (!pl)
        return KPO_ERR_NOMEM;
    memset(img, 0, sizeof *img);
    int rc = parse_container(file, len, pl);
    if (rc != KPO_OK) {
        free(pl);
        return rc;
    }
    img->width = pl->width;
    img->height = pl->height;
    img->ncomp = pl->ncomp;
    img->mcus_x = (pl->width + 7) / 8;
    img->mcus_y = (pl->height + 7) / 8;
    img->restart_interval = pl->restart_interval;
    img->nblocks = (int64_t)img->mcus_x * img->mcus_y * pl->ncomp;
    img->scan_bytes = (int64_t)pl->scan_len;
    img->coef = (int16_t *)malloc((size_t)img->nblocks * 64 * sizeof(int16_t));
    if (!img->coef) {
        free(pl);
        return KPO_ERR_NOMEM;
    }
    img->scan_bytes = 0;
    for (int k = 0; k < pl->nscans && rc == KPO_OK; ++k) {
        img->scan_bytes += (int64_t)pl->scans[k].len;
        rc = decode_coefficients(pl, &pl->scans[k], file + pl->scans[k].off, pl->scans[k].len, flags, img->coef);
    }
    if (rc == KPO_OK && want_pixels) {
        img->pixels = (uint8_t *)malloc((size_t)pl->width * pl->height * pl->ncomp);
        if (!img->pixels)
            rc = KPO_ERR_NOMEM;
        else
            reconstruct(pl, img->coef, img->pixels);
    }
    free(pl);
    if (rc != KPO_OK)
        kpo_free(img);
    return rc;
}

void kpo_free(kpo_image *img)
{
    free(img->coef);
    free(img->pixels);
    img->coef = NULL;
    img->pixels = NULL;
}

int kpo_ppm_header(int width, int height, char *buf, size_t cap)
{
    /* Image.cpp:124-127 */
    return snprintf(buf, cap,
                    "P6\n# PPM dump created using libKPEG: https://github.com/TheIllusionistMirage/libKPEG\n"
                    "%d %d\n255\n",
                    width, height);
}
