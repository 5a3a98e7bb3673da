"""The committed deterministic synthetic encoder (libkpeg_b200/host/synth_encoder.cpp): determinism,
twin-stream property (identical quantised coefficients with/without DRI, 1 vs 3 components), the
quirk-free guarantee and independent decodability (PIL / libjpeg)."""
import hashlib
import io

import numpy as np
import pytest

import helpers as H
from libkpeg_b200.synth import EMIT_RESTART, GRAY_CONTENT, QUIRK_FREE, SynthParams, synth_encode, synth_pixels


def test_deterministic_and_thread_independent():
    a = synth_encode(SynthParams(200, 120, quality=90, seed=5, threads=1))
    b = synth_encode(SynthParams(200, 120, quality=90, seed=5, threads=7))
    c = synth_encode(SynthParams(200, 120, quality=90, seed=6, threads=7))
    assert np.array_equal(a, b) and not np.array_equal(a, c)


def test_segment_order_is_lenas(lena_jpg):
    """SOI, APP0, DQT, DQT, SOF0, DHT x4, SOS -- the layout the reference accepts."""
    def markers(buf):
        out, i = [], 2
        while buf[i] == 0xFF:
            m = buf[i + 1]
            out.append(m)
            if m == 0xDA:
                break
            i += 2 + ((buf[i + 2] << 8) | buf[i + 3])
        return out
    j = synth_encode(SynthParams(64, 64)).tobytes()
    assert markers(j) == markers(lena_jpg) == [0xE0, 0xDB, 0xDB, 0xC0, 0xC4, 0xC4, 0xC4, 0xC4, 0xDA]
    assert j[-2:] == b"\xff\xd9"


@pytest.mark.parametrize("ri", [1, 5, 16])
def test_restart_twins_share_coefficients(ri):
    base = dict(width=120, height=72, quality=88, restart_interval=ri, seed=ri)
    a = H.oracle_decode(synth_encode(SynthParams(**base, flags=QUIRK_FREE)).tobytes(), parity=False, want_pixels=False)
    b = H.oracle_decode(synth_encode(SynthParams(**base, flags=QUIRK_FREE | EMIT_RESTART)).tobytes(), parity=False,
                        want_pixels=False)
    assert b["restart_interval"] == ri and a["restart_interval"] == 0
    assert np.array_equal(a["coef"], b["coef"])


def test_quirk_free_streams_are_insensitive_to_f1():
    for flags in (QUIRK_FREE, QUIRK_FREE | EMIT_RESTART):
        j = synth_encode(SynthParams(160, 96, quality=90, restart_interval=7, seed=3, flags=flags)).tobytes()
        a = H.oracle_decode(j, parity=True, want_pixels=False)["coef"]
        b = H.oracle_decode(j, parity=False, want_pixels=False)["coef"]
        assert np.array_equal(a, b)
    # without the nudge the quirk fires on some block
    j = synth_encode(SynthParams(256, 256, quality=50, seed=3, flags=0)).tobytes()
    assert not np.array_equal(H.oracle_decode(j, parity=True, want_pixels=False)["coef"],
                              H.oracle_decode(j, parity=False, want_pixels=False)["coef"])


def test_gray_twins_share_luma():
    base = dict(width=72, height=40, quality=90, seed=9)
    g1 = H.oracle_decode(synth_encode(SynthParams(**base, file_components=1, flags=QUIRK_FREE | GRAY_CONTENT)).tobytes(),
                         want_pixels=False)
    g3 = H.oracle_decode(synth_encode(SynthParams(**base, file_components=3, flags=QUIRK_FREE | GRAY_CONTENT)).tobytes(),
                         want_pixels=False)
    assert np.array_equal(g1["coef"], g3["coef"][0::3])
    assert not g3["coef"][1::3].any() and not g3["coef"][2::3].any()


def test_decodable_by_libjpeg_and_close_to_source():
    PIL = pytest.importorskip("PIL.Image")
    p = SynthParams(256, 192, quality=95, seed=2)
    j = synth_encode(p).tobytes()
    ext = np.asarray(PIL.open(io.BytesIO(j)).convert("RGB")).astype(np.float64)
    src = synth_pixels(p).astype(np.float64)
    mse = ((ext - src) ** 2).mean()
    assert 10 * np.log10(255 ** 2 / mse) > 35.0
    ours = H.oracle_decode(j, parity=False)["pixels"].astype(np.float64)
    assert np.abs(ours - ext).max() <= 6  # float IDCT + floor vs libjpeg's integer pipeline


def test_bits_per_pixel_of_the_benchmark_content():
    j = synth_encode(SynthParams(512, 512, quality=95, seed=1))
    bpp = j.size * 8 / (512 * 512)
    assert 4.5 < bpp < 8.0  # ~6.5 bit/px at q95
