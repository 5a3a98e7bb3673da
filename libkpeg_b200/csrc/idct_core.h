// idct_core.h -- reconstruction arithmetic of the B200 decode path (host + device inline).
//
// Replaces MCU::constructMCU's dequantise / de-zigzag, MCU::computeIDCT, performLevelShift and
// convertYCbCrToRGB (reference src/MCU.cpp:110-120, 172-216, 218-245, 247-279).
//
// Parity strategy.  The reference evaluates the 2-D IDCT as a 64-term sum with a
// *float* accumulator and *double* products, then rounds half away from zero.  About 0.7 % of all
// samples of a natural image are exact x.5 ties in real arithmetic and the reference resolves them
// by its own rounding noise, so no "accurate" IDCT reproduces its integers.  We therefore run
//   (1) a fast separable fp32 IDCT (AAN factorisation: 25 adds, 1 multiply, 4 fused multiply-adds per 8 points) and
//       accept its rounding wherever the value is provably further from a tie than the combined
//       error bound of both evaluations, and
//   (2) for the few samples inside that band, the reference's own operation sequence
//       (exact_sample below: same operand types, same order, no FMA contraction),
// which makes the integer samples -- and hence the RGB bytes -- identical to the reference's.
// The colour conversion uses the same two-tier scheme (fp32 with a provable margin, the
// reference's double expression otherwise).
#ifndef KPEG_IDCT_CORE_H
#define KPEG_IDCT_CORE_H

#include <math.h>
#include <string.h>

#include "kpeg_common.h"

namespace kpeg {

// ---- explicitly rounded primitives (never contracted into FMAs) ------------------------------
KPEG_HD float mul_f32(float a, float b)
{
#if defined(__CUDA_ARCH__)
    return __fmul_rn(a, b);
#else
    return a * b;
#endif
}
KPEG_HD double mul_f64(double a, double b)
{
#if defined(__CUDA_ARCH__)
    return __dmul_rn(a, b);
#else
    return a * b;
#endif
}
KPEG_HD double add_f64(double a, double b)
{
#if defined(__CUDA_ARCH__)
    return __dadd_rn(a, b);
#else
    return a + b;
#endif
}
KPEG_HD double sub_f64(double a, double b)
{
#if defined(__CUDA_ARCH__)
    return __dsub_rn(a, b);
#else
    return a - b;
#endif
}

// zig-zag index -> natural (row*8+col) index; same map as Transform.cpp:5-27
// (zzOrderToMatIndices), produced by walking the anti-diagonals.  constexpr so that fully
// unrolled kernels turn it into register renaming.
KPEG_HD constexpr int zigzag_to_natural(int i)
{
    int idx = 0;
    for (int d = 0; d < 15; ++d) {
        const int lo = d < 8 ? 0 : d - 7, hi = d < 8 ? d : 7;
        for (int k = lo; k <= hi; ++k) {
            const int r = (d & 1) ? k : (hi + lo - k);
            if (idx == i)
                return r * 8 + (d - r);
            ++idx;
        }
    }
    return 63;
}

struct ZigZagTables {
    unsigned char zz2nat[64];
    unsigned char nat2zz[64];
};
KPEG_HD constexpr ZigZagTables make_zigzag_tables()
{
    ZigZagTables t{};
    for (int i = 0; i < 64; ++i) {
        t.zz2nat[i] = (unsigned char)zigzag_to_natural(i);
    }
    for (int i = 0; i < 64; ++i)
        t.nat2zz[t.zz2nat[i]] = (unsigned char)i;
    return t;
}

// AAN prescale factors: s[0] = 1, s[k] = sqrt(2) * cos(k*pi/16).
KPEG_HD constexpr double aan_scale_c(int k)
{
    return k == 0 ? 1.0 : k == 1 ? 1.387039845322148 : k == 2 ? 1.306562964876377 : k == 3 ? 1.175875602419359
         : k == 4 ? 1.0 : k == 5 ? 0.785694958387102 : k == 6 ? 0.541196100146197 : 0.275899379282943;
}

// 1 / (prescale of the coefficient at natural index nat): |prescaled dequantised coefficient| * this = |c| * q
KPEG_HD constexpr float aan_unscale(int nat) { return (float)(8.0 / (aan_scale_c(nat >> 3) * aan_scale_c(nat & 7))); }

// Pair layout of a block going into the packed transform (kernels.cu): pair p = rp * 8 + c holds column c of two
// rows side by side -- rows (0,1), (4,7), (2,5), (6,3) for rp = 0..3.  The row pass transforms two rows per
// instruction; in the column pass these are exactly the operand pairs of the first two butterfly stages
// (v0 +- v4 beside v1 +- v7, v2 +- v6 beside v5 +- v3), so it needs no register transposition.
KPEG_HD constexpr int pair_row(int rp, int half) { return rp == 0 ? (half ? 1 : 0) : rp == 1 ? (half ? 7 : 4) : rp == 2 ? (half ? 5 : 2) : (half ? 3 : 6); }
KPEG_HD constexpr int pair_nat(int p, int half) { return pair_row(p >> 3, half) * 8 + (p & 7); }

inline double aan_scale(int k)
{
    const double s[8] = {1.0, 1.387039845322148, 1.306562964876377, 1.175875602419359,
                         1.0, 0.785694958387102, 0.541196100146197, 0.275899379282943};
    return s[k];
}

// Lane operations of the fast IDCT.  The transform is written once (idct8_aan below) over a lane type V:
// `float` here (host single-stepper, scalar device code) and the packed two-lane type of kernels.cu (sm_100a
// FADD2 / FMUL2 / FFMA2: two independent IEEE fp32 lanes per instruction).  Every fused multiply-add is explicit,
// so all instantiations round identically, lane for lane.
KPEG_HD float lane_add(float a, float b)
{
#if defined(__CUDA_ARCH__)
    return __fadd_rn(a, b);
#else
    return a + b;
#endif
}
KPEG_HD float lane_sub(float a, float b)
{
#if defined(__CUDA_ARCH__)
    return __fsub_rn(a, b);
#else
    return a - b;
#endif
}
KPEG_HD float lane_mul_k(float a, float k) { return mul_f32(a, k); }
KPEG_HD float lane_fma_k(float a, float k, float c) { return fmaf(a, k, c); } // a * k + c, one rounding

// One 8-point AAN inverse DCT on prescaled inputs, in place: 25 adds, 1 multiply, 4 fused multiply-adds.
template <class V>
KPEG_HD void idct8_aan(V &v0, V &v1, V &v2, V &v3, V &v4, V &v5, V &v6, V &v7)
{
    // even part
    const V t10 = lane_add(v0, v4), t11 = lane_sub(v0, v4);
    const V t13 = lane_add(v2, v6);
    const V n12 = lane_fma_k(lane_sub(v2, v6), -1.414213562373095f, t13); // -(t12)
    const V e0 = lane_add(t10, t13), e3 = lane_sub(t10, t13), e1 = lane_sub(t11, n12), e2 = lane_add(t11, n12);
    // odd part
    const V z13 = lane_add(v5, v3), z10 = lane_sub(v5, v3), z11 = lane_add(v1, v7), z12 = lane_sub(v1, v7);
    const V o7 = lane_add(z11, z13);
    const V z5 = lane_mul_k(lane_add(z10, z12), 1.847759065022573f);
    const V t20 = lane_fma_k(z12, -1.082392200292394f, z5);
    const V t22 = lane_fma_k(z10, -2.613125929752753f, z5);
    const V o6 = lane_sub(t22, o7);
    const V n5 = lane_fma_k(lane_sub(z11, z13), -1.414213562373095f, o6); // -(o5)
    const V o4 = lane_add(t20, n5);
    v0 = lane_add(e0, o7);
    v7 = lane_sub(e0, o7);
    v1 = lane_add(e1, o6);
    v6 = lane_sub(e1, o6);
    v2 = lane_sub(e2, n5);
    v5 = lane_add(e2, n5);
    v3 = lane_add(e3, o4);
    v4 = lane_sub(e3, o4);
}

// Fast 2-D IDCT of one block held as 64 prescaled floats in NATURAL order (f[row*8+col]);
// result overwrites f: f[row*8+col] = sample before the +128 level shift.  Rows first, then columns -- the
// order of the packed device version (kernels.cu idct8x8_packed), which this scalar form matches bit for bit.
KPEG_HD void idct8x8_fast(float f[64])
{
#pragma unroll
    for (int r = 0; r < 8; ++r) // rows
        idct8_aan(f[8 * r], f[8 * r + 1], f[8 * r + 2], f[8 * r + 3], f[8 * r + 4], f[8 * r + 5], f[8 * r + 6],
                  f[8 * r + 7]);
#pragma unroll
    for (int c = 0; c < 8; ++c) // columns
        idct8_aan(f[c], f[8 + c], f[16 + c], f[24 + c], f[32 + c], f[40 + c], f[48 + c], f[56 + c]);
}

// |fast - reference| <= TIE_REL * A + TIE_ABS with A = sum |dequantised coefficient| = sum |c_i| * q_i
// (derivation in DESIGN.md "K3 rounding"; margin checked by tests/test_emu_logic.py).  The kernel
// accumulates A with one FFMA per coefficient while dequantising.
constexpr float TIE_REL = 1.5e-6f;
constexpr float TIE_ABS = 2.0e-5f;

KPEG_HD float tie_band(float A) { return TIE_REL * A + TIE_ABS; }

// round-half-away-from-zero of a float, as roundl() does on the promoted value (MCU.cpp:228).
KPEG_HD int round_half_away(float out)
{
    const float a = fabsf(out);
    float r = floorf(a);
    if (a - r >= 0.5f)
        r += 1.0f;
    const int ri = (int)r;
    return out < 0.0f ? -ri : ri;
}

// The reference's own evaluation of one output sample (row x, column y) of one block:
// src/MCU.cpp:178-199 + :228.  coef_at(i) = quantised coefficient at zig-zag index i (DC already
// integrated), q = quantiser in zig-zag order.  Returns the UNSHIFTED rounded sample
// (the reference then adds 128).  Terms with a zero coefficient add +-0.0 to the float
// accumulator, which leaves it unchanged, so they are skipped.
template <class CoefAt>
KPEG_HD int exact_sample(const CoefAt &coef_at, const int32_t *q, const double (*cosd)[8], const float (*cc)[8],
                         const unsigned char *nat2zz, int x, int y)
{
    float sum = 0.0f;
    for (int nat = 0; nat < 64; ++nat) { // u outer, v inner == ascending natural index (MCU.cpp:184-187)
        const int zi = nat2zz[nat];
        const int cq = coef_at(zi);
        if (cq == 0)
            continue;
        const int u = nat >> 3, v = nat & 7;
        const int F = cq * q[zi];                      // MCU.cpp:110-112 (dequantise), :115-120 (de-zigzag)
        const float t = mul_f32(cc[u][v], (float)F);   // Cu * Cv * block : float
        const double d = mul_f64(mul_f64((double)t, cosd[x][u]), cosd[y][v]);
        sum = (float)add_f64((double)sum, d);          // float accumulator, rounded every term
    }
    const float out = (float)mul_f64(0.25, (double)sum); // MCU.cpp:198
    return round_half_away(out);
}

// ---- colour ------------------------------------------------------------------------------------
// MCU.cpp:255-265 on UNSHIFTED samples y, cb, cr (reference value = sample + 128):
//   R = floor(Y + 1.402 (Cr-128)), G = floor(Y - 0.344136 (Cb-128) - 0.714136 (Cr-128)),
//   B = floor(Y + 1.772 (Cb-128)), all in double, then clamped to [0,255].
KPEG_HD void ycc_to_rgb_exact(int y, int cb, int cr, int &R, int &G, int &B)
{
    const double Y = (double)(y + 128), db = (double)cb, dr = (double)cr;
    const double r = add_f64(Y, mul_f64(1.402, dr));
    const double g = sub_f64(sub_f64(Y, mul_f64(0.344136, db)), mul_f64(0.714136, dr));
    const double b = add_f64(Y, mul_f64(1.772, db));
    R = (int)floor(r);
    G = (int)floor(g);
    B = (int)floor(b);
    R = R < 0 ? 0 : (R > 255 ? 255 : R);
    G = G < 0 ? 0 : (G > 255 ? 255 : G);
    B = B < 0 ? 0 : (B > 255 ? 255 : B);
}

// fp32 evaluation, valid when |y|,|cb|,|cr| <= COLOUR_FAST_RANGE (all integer-valued):
//  * 1.402 d = 701 d / 500 and 1.772 d = 443 d / 250 are either integers or at least 0.002 away
//    from one; with a +0.001 bias and < 5e-4 of accumulated fp32 error the floor cannot flip, and
//    for the integer case the reference's double expression yields that integer too (checked
//    exhaustively by tests/test_emu_logic.py);
//  * G's fraction is a multiple of 8e-6: if the fp32 value is closer than COLOUR_G_BAND to an
//    integer the caller must use ycc_to_rgb_exact (returns false);
//  * floor(v) is taken as rint(v - 0.5) through the 1.5*2^23 magic-number add, which is exact for
//    any v that is not within the error bound of an integer -- guaranteed by the two points above.
// Outputs are unclamped integers.
constexpr float COLOUR_FAST_RANGE = 1023.0f;
constexpr float COLOUR_G_BAND = 1.0e-3f;
constexpr float RINT_MAGIC = 12582912.0f; // 1.5 * 2^23 == 0x4B400000
constexpr int32_t RINT_MAGIC_BITS = 0x4B400000;

KPEG_HD int32_t float_bits(float x)
{
#if defined(__CUDA_ARCH__)
    return __float_as_int(x);
#else
    int32_t i;
    memcpy(&i, &x, 4);
    return i;
#endif
}

// (double)f for f zero or normal, by integer arithmetic: sign | (exponent + 896) << 20 | mantissa >> 3 in the high
// word, the low three mantissa bits on top of the low word.  The hardware conversion (F2F.F64.F32) runs on the
// quarter-rate XU pipe, and the exact re-evaluation of a sample (idct_patch_kernel, k3_fused.cu) does two of them per
// term of its 64-term chain: 128 of its 193 conversions per sample.  Denormal inputs cannot occur there: a term is
// cc * F with F an integer (|t| >= 0.49 or 0), and the float accumulator holds 0 or a residue of such terms (>= 2^-53
// times their size), never a value below 2^-126.  Checked against the cast in tests/test_emu_logic.py.
KPEG_HD double widen_f32(float f)
{
    const uint32_t b = (uint32_t)float_bits(f), a = b & 0x7FFFFFFFu;
    const uint32_t hi = (b & 0x80000000u) | (a ? (a >> 3) + 0x38000000u : 0u), lo = b << 29;
#if defined(__CUDA_ARCH__)
    return __hiloint2double((int)hi, (int)lo);
#else
    const uint64_t u = ((uint64_t)hi << 32) | lo;
    double d;
    memcpy(&d, &u, 8);
    return d;
#endif
}

KPEG_HD bool ycc_to_rgb_fast(float y, float cb, float cr, int &R, int &G, int &B)
{
    const float yr = y + 127.501f; // +128 level shift, +0.001 bias, -0.5 (floor by rint)
    const float yg = y + 127.5f;
    const float tr = fmaf(1.402f, cr, yr) + RINT_MAGIC;
    const float tb = fmaf(1.772f, cb, yr) + RINT_MAGIC;
    const float g = fmaf(-0.714136f, cr, fmaf(-0.344136f, cb, yg));
    const float tg = g + RINT_MAGIC;
    const float dg = g - (tg - RINT_MAGIC); // |dg| close to 0.5 <=> the true G is close to an integer
    R = float_bits(tr) - RINT_MAGIC_BITS;
    B = float_bits(tb) - RINT_MAGIC_BITS;
    G = float_bits(tg) - RINT_MAGIC_BITS;
    // cb == cr == 0 (flat chroma, gray-as-YCbCr files): every channel is exactly y + 128
    const bool flat = (cb == 0.0f) && (cr == 0.0f);
    G = flat ? R : G;
    const float m = fmaxf(fmaxf(fabsf(y), fabsf(cb)), fabsf(cr));
    return ((fabsf(dg) < 0.5f - COLOUR_G_BAND) || flat) && (m <= COLOUR_FAST_RANGE);
}

KPEG_HD int clamp_u8(int v) { return v < 0 ? 0 : (v > 255 ? 255 : v); }

} // namespace kpeg
#endif
