"""CPU single-stepping of the kernel logic (tests/emu: the same entropy_core.h / idct_core.h /
unstuff_core.h the sm_100a kernels compile) against the oracle: speculative decode + relay fixed
point + offset scan, restart / image boundaries, the two-tier IDCT and colour arithmetic."""
import ctypes as C

import numpy as np
import pytest
from hypothesis import example, given, settings
from hypothesis import strategies as st

import helpers as H
import libkpeg_b200 as K
from libkpeg_b200.synth import EMIT_RESTART, GRAY_CONTENT, QUIRK_FREE, SynthParams, synth_encode


def agree(jpg, parity=True, **kw):
    o = H.oracle_decode(jpg, parity=parity)
    e = H.emu_decode(jpg, flags=1 if parity else 0, **kw)
    assert e["status"] == 0
    assert e["records_ok"], "record final pass differs from the Huffman final pass"
    assert e["max_records"] <= kw.get("sub_bits", 512) * 3 // 8 + 1  # what rec_kmax provides (kpeg_cuda.cu) for Annex-K-like tables
    assert np.array_equal(o["coef"], e["coef"]), "coefficients"
    assert np.array_equal(o["pixels"], e["pixels"]), "pixels"
    return o, e


@pytest.mark.parametrize("sub_bits", [64, 96, 512, 1024, 8192])
def test_lena_all_subsequence_sizes(lena_jpg, sub_bits):
    o, e = agree(lena_jpg, sub_bits=sub_bits)
    assert e["final_slot"] == 512 * 512 // 64 * 3 * 64


def test_lena_t81_mode(lena_jpg):
    agree(lena_jpg, parity=False)


@pytest.mark.parametrize("w,h,q,ri", [(64, 48, 90, 0), (256, 64, 90, 16), (200, 120, 75, 7), (96, 96, 50, 1),
                                       (320, 240, 95, 40), (64, 64, 100, 3)])
def test_restart_twins(w, h, q, ri):
    base = dict(width=w, height=h, quality=q, restart_interval=ri, seed=w * 31 + h, noise_amp=30 if q == 100 else 0)
    plain = synth_encode(SynthParams(**base, flags=QUIRK_FREE)).tobytes()
    rst = synth_encode(SynthParams(**base, flags=QUIRK_FREE | EMIT_RESTART)).tobytes()
    _, a = agree(plain)
    _, b = agree(rst, sub_bits=128)
    assert np.array_equal(a["pixels"], b["pixels"])
    assert np.array_equal(a["coef"], b["coef"])


@pytest.mark.parametrize("w,h", [(96, 96), (40, 24)])
def test_gray_twins(w, h):
    base = dict(width=w, height=h, quality=90, seed=w + h)
    g1 = synth_encode(SynthParams(**base, file_components=1, flags=QUIRK_FREE | GRAY_CONTENT)).tobytes()
    g3 = synth_encode(SynthParams(**base, file_components=3, flags=QUIRK_FREE | GRAY_CONTENT)).tobytes()
    _, a = agree(g1)
    _, b = agree(g3)
    assert np.array_equal(a["pixels"], b["pixels"][..., 1])
    assert b["colour_exact"] == 0  # flat chroma stays on the fast colour path


@pytest.mark.parametrize("w,h", [(60, 45), (17, 9), (8, 8), (1, 1), (1000, 3)])
def test_ragged(w, h):
    agree(synth_encode(SynthParams(w, h, quality=85, seed=w * h)).tobytes())


def test_batch_as_one_stream():
    jpgs = [synth_encode(SynthParams(64, 40, quality=92, restart_interval=5, flags=QUIRK_FREE | EMIT_RESTART, seed=7 + i))
            for i in range(6)]
    parsed = [K.parse_jfif(j) for j in jpgs]
    scans = [j[o:o + n] for j, (_, o, n) in zip(jpgs, parsed)]
    e = H.emu_decode(None, scans=scans, plan=parsed[0][0], sub_bits=256)
    assert e["status"] == 0
    for i, j in enumerate(jpgs):
        o = H.oracle_decode(j.tobytes())
        assert np.array_equal(o["pixels"], e["pixels"][i])


def test_corrupt_stream_flags_an_error(lena_jpg):
    buf = np.frombuffer(lena_jpg, dtype=np.uint8).copy()
    plan, off, n = K.parse_jfif(buf)
    e = H.emu_decode(buf[: off + n // 2].tobytes() + b"\xff\xd9", want_pixels=False)
    assert e["status"] != 0  # truncated: the stream ends before the last MCU
    assert e["records_ok"], "the record flavour (what the product runs) must reject a truncated stream as well"
    rng = np.random.default_rng(3)
    bad = buf.copy()
    idx = rng.integers(off + 100, off + n - 100, size=200)
    bad[idx] = rng.integers(1, 255, size=200).astype(np.uint8)
    H.emu_decode(bad, want_pixels=False)  # must terminate; status is content dependent


@settings(max_examples=25, deadline=None)
@given(w=st.integers(1, 96), h=st.integers(1, 64), q=st.integers(5, 100), ri=st.integers(0, 9),
       nc=st.sampled_from([1, 3]), seed=st.integers(0, 2 ** 31), sb=st.sampled_from([64, 128, 512]))
# a subsequence entered in the padding bits BEFORE a restart boundary whose slot is also where a strip's predictors restart
@example(w=57, h=33, q=5, ri=7, nc=1, seed=5, sb=64)
def test_property_random_streams(w, h, q, ri, nc, seed, sb):
    flags = QUIRK_FREE | (EMIT_RESTART if ri else 0) | (GRAY_CONTENT if nc == 1 else 0)
    jpg = synth_encode(SynthParams(w, h, file_components=nc, quality=q, restart_interval=ri, flags=flags, seed=seed,
                                   noise_amp=(seed % 60) + 1)).tobytes()
    agree(jpg, sub_bits=sb)


def test_fast_idct_error_bound():
    """|fast fp32 IDCT - reference evaluation| stays inside the tie band the kernel uses (idct_core.h
    TIE_REL / TIE_ABS), on random sparse and dense blocks up to the extremes of 8-bit JPEG."""
    rng = np.random.default_rng(11)
    emu, orc = H.emu(), H.oracle()
    qt = np.ones(64, dtype=np.uint16)
    worst = 0.0
    for trial in range(3000):
        dens = rng.choice([1, 3, 8, 20, 64])
        amp = rng.choice([4, 40, 400, 2000])
        zz = np.zeros(64, dtype=np.int16)
        idx = rng.choice(64, size=dens, replace=False)
        zz[idx] = rng.integers(-amp, amp + 1, size=dens)
        qt[:] = rng.integers(1, 256)
        fast = np.zeros(64, dtype=np.float32)
        band = emu.emu_idct_fast(zz.ctypes.data, qt.ctypes.data, fast.ctypes.data)
        # reference-order evaluation (float accumulator) before rounding
        F = np.zeros(64, dtype=np.int32)
        r, c = C.c_int(), C.c_int()
        for i in range(64):
            orc.kpo_zigzag_to_rc(i, C.byref(r), C.byref(c))
            F[r.value * 8 + c.value] = int(zz[i]) * int(qt[i])
        ref = np.zeros(64, dtype=np.float32)
        orc.kpo_idct8x8(F.ctypes.data, ref.ctypes.data)
        err = float(np.abs(fast.astype(np.float64) - ref.astype(np.float64)).max())
        assert err <= band, (trial, err, band)
        worst = max(worst, err / band)
    assert worst < 0.9  # keep some margin


def test_colour_two_tier_exhaustive_slices():
    """Fast fp32 colour path + double fallback == the reference's double expression, on dense slices of
    the (Y, Cb, Cr) cube including the lattice points where G is an exact integer."""
    emu, orc = H.emu(), H.oracle()
    got, want = (C.c_int * 3)(), (C.c_int * 3)()
    n_exact = 0
    for y in (-300, -129, -128, -1, 0, 1, 77, 127, 128, 400):
        for cb in range(-260, 261, 3):
            for cr in range(-260, 261, 5):
                n_exact += emu.emu_colour(y, cb, cr, got)
                orc.kpo_ycbcr_to_rgb(y + 128, cb + 128, cr + 128, want)
                assert list(got) == list(want), (y, cb, cr)
    # G = Y exactly when 43017*cb + 89267*cr == 0 (mod 125000): all such points near the origin
    for cb in range(-600, 601):
        for cr in range(-600, 601):
            if (43017 * cb + 89267 * cr) % 125000 == 0:
                for y in (-140, -3, 0, 5, 131):
                    emu.emu_colour(y, cb, cr, got)
                    orc.kpo_ycbcr_to_rgb(y + 128, cb + 128, cr + 128, want)
                    assert list(got) == list(want), (y, cb, cr)
    assert n_exact < 2000  # the fallback is rare


def test_colour_rb_integer_points():
    """1.402 d and 1.772 d are exact integers for d = 500 k / 250 k: the fast path must agree there."""
    emu, orc = H.emu(), H.oracle()
    got, want = (C.c_int * 3)(), (C.c_int * 3)()
    for d in (-1000, -750, -500, -250, 250, 500, 750, 1000):
        for y in (-128, 0, 100):
            emu.emu_colour(y, d, d, got)
            orc.kpo_ycbcr_to_rgb(y + 128, d + 128, d + 128, want)
            assert list(got) == list(want)


def test_byte_parallel_unstuff_classification_matches_bytewise():
    """K0's sixteen-bytes-at-a-time classification (classify16_swar, what the count kernel runs) against the
    byte-wise classify16 on adversarial byte soup: dense FF / 00 / RSTn / fill / stray markers, chunk-edge cases."""
    import ctypes as C
    lib = H.emu()
    rng = np.random.default_rng(0xFF00)
    alphabet = np.array([0xFF, 0xFF, 0xFF, 0x00, 0x00, 0xD0, 0xD3, 0xD7, 0xD9, 0xC4, 0x01, 0x7F, 0x80, 0xFE, 0xD8, 0xCF],
                        dtype=np.uint8)
    for trial in range(40):
        n = int(rng.integers(16, 4096))
        if trial % 3 == 0:
            buf = rng.integers(0, 256, size=n, dtype=np.uint8)          # FF is rare
        else:
            buf = alphabet[rng.integers(0, alphabet.size, size=n)]      # FF / 00 / markers everywhere
        buf = np.ascontiguousarray(buf)
        bad = C.c_int(0)
        mism = lib.emu_classify_compare(buf.ctypes.data, n, C.byref(bad))
        assert mism == 0, f"trial {trial}: {mism} chunks classified differently"
        assert bad.value == 0, f"trial {trial}: unexpected-marker verdict differs"


def test_pair_order_quantiser_tables(lena_jpg):
    """The tables K3 stages by bulk copy (host_tables.h): `qpair` is `qscale` permuted into the pair order of the
    packed transform, `qdc` keeps only the DC entry; the pair order is a permutation of the 64 positions that puts
    rows (0,1), (4,7), (2,5), (6,3) side by side at every column -- the operand pairs of the first two butterfly
    stages of the column pass (idct_core.h pair_nat, kernels.cu idct8_column)."""
    plan, _, _ = K.parse_jfif(np.frombuffer(lena_jpg, dtype=np.uint8))
    emu = H.emu()
    qscale = np.zeros((3, 64), dtype=np.float32)
    qpair = np.zeros((3, 64), dtype=np.float32)
    qdc = np.zeros((3, 64), dtype=np.float32)
    pn = np.zeros(64, dtype=np.int32)
    assert emu.emu_pair_tables(C.addressof(plan), qscale.ctypes.data, qpair.ctypes.data, qdc.ctypes.data, pn.ctypes.data) == 0
    assert sorted(pn.tolist()) == list(range(64))
    rows = {(int(pn[2 * p]) // 8, int(pn[2 * p + 1]) // 8) for p in range(32)}
    assert rows == {(0, 1), (4, 7), (2, 5), (6, 3)}
    for p in range(32):
        assert pn[2 * p] % 8 == pn[2 * p + 1] % 8 == p % 8  # both halves of a pair sit in the same column
    nat2zz = {emu.emu_zigzag(i): i for i in range(64)}  # emu_zigzag: zig-zag index -> natural position
    for c in range(plan.ncomp):
        assert (qscale[c] > 0).all()
        for j in range(64):
            assert qpair[c, j] == qscale[c, nat2zz[int(pn[j])]]
            assert qdc[c, j] == (qpair[c, j] if pn[j] == 0 else 0.0)


def test_float_to_double_by_integer_arithmetic():
    """widen_f32 (idct_core.h), which the exact sample evaluation can use instead of the hardware conversion
    (-DKPEG_EXACT_WIDEN_INT=1): equal to the cast for +-0 and every normal float -- random bit patterns, the extremes,
    integers times the reference's float scale factors (the only operands it ever sees)."""
    import ctypes as C
    lib = H.emu()
    lib.emu_widen_f32.restype = None
    lib.emu_widen_f32.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t]
    rng = np.random.default_rng(9)
    bits = rng.integers(0, 2 ** 32, size=2_000_000, dtype=np.uint64).astype(np.uint32)
    exp = (bits >> 23) & 0xFF
    bits = bits[(exp != 0) & (exp != 255)]  # normal floats
    f = np.concatenate([bits.view(np.float32),
                        np.array([0.0, -0.0, np.finfo(np.float32).tiny, -np.finfo(np.float32).tiny, np.finfo(np.float32).max,
                                  -np.finfo(np.float32).max, 1.0, -1.0, 0.49999997, 0.70710677], dtype=np.float32),
                        (np.arange(-70000, 70000, dtype=np.float32) * np.float32(0.70710677)),
                        (np.arange(-70000, 70000, dtype=np.float32) * np.float32(0.49999997))])
    f = np.ascontiguousarray(f, dtype=np.float32)
    out = np.empty(f.size, dtype=np.float64)
    lib.emu_widen_f32(f.ctypes.data, out.ctypes.data, f.size)
    assert np.array_equal(out.view(np.uint64), f.astype(np.float64).view(np.uint64))
