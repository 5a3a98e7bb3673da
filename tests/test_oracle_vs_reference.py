"""Pins the oracle (oracle/kpeg_oracle.c) to the reference: byte-exact against the committed golden
output of the compiled reference, and -- when the reference binary is present (build container, or
shipped to the GPU box in oracle/_ref/) -- against live runs on synthetic streams."""
import hashlib
import json

import numpy as np
import pytest

import helpers as H
import libkpeg_b200 as K
from libkpeg_b200.synth import QUIRK_FREE, SynthParams, synth_encode


def test_oracle_matches_golden_lena(lena_jpg):
    g = np.load(H.GOLDEN / "lena_ref.npz")
    assert hashlib.sha256(lena_jpg).hexdigest() == str(g["jpg_sha256"])
    o = H.oracle_decode(lena_jpg, parity=True)
    assert np.array_equal(o["pixels"], g["payload"])
    ppm = bytes(g["header"].tobytes()) + o["pixels"].tobytes()
    assert hashlib.sha256(ppm).hexdigest() == str(g["ppm_sha256"])
    assert K.ppm_header(512, 512) == g["header"].tobytes()


def test_oracle_f1_quirk_matters_on_lena(lena_jpg):
    """without the DC-difference rule the result is far from the reference's."""
    g = np.load(H.GOLDEN / "lena_ref.npz")
    o = H.oracle_decode(lena_jpg, parity=False)
    d = np.abs(o["pixels"].astype(int) - g["payload"].astype(int))
    assert d.max() > 50
    c0 = H.oracle_decode(lena_jpg, parity=True, want_pixels=False)["coef"]
    assert (c0 != o["coef"]).any(axis=1).sum() == 502  # blocks hit by the quirk


def test_oracle_matches_golden_twins():
    twins = json.loads((H.GOLDEN / "twins.json").read_text())
    for name, t in twins.items():
        p = SynthParams(**{"flags": QUIRK_FREE, **t["params"]})
        jpg = synth_encode(p).tobytes()
        assert hashlib.sha256(jpg).hexdigest() == t["jpg_sha256"], name
        o = H.oracle_decode(jpg)
        ppm = K.ppm_header(p.width, p.height) + o["pixels"].tobytes()
        assert hashlib.sha256(ppm).hexdigest() == t["ppm_sha256"], name


@pytest.mark.skipif(not H.have_reference_binary(), reason="compiled reference not present")
@pytest.mark.parametrize("w,h,q,seed", [(64, 64, 90, 1), (96, 48, 95, 2), (48, 120, 35, 3), (128, 128, 100, 4)])
def test_oracle_vs_live_reference(w, h, q, seed):
    jpg = synth_encode(SynthParams(w, h, quality=q, seed=seed, noise_amp=20 if q == 100 else 0)).tobytes()
    hdr, payload = H.split_ppm(H.reference_decode(jpg))
    o = H.oracle_decode(jpg)
    assert hdr == K.ppm_header(w, h)
    assert np.array_equal(o["pixels"], payload)


@pytest.mark.skipif(not H.have_reference_binary(), reason="compiled reference not present")
def test_live_reference_lena(lena_jpg):
    g = np.load(H.GOLDEN / "lena_ref.npz")
    ppm = H.reference_decode(lena_jpg)
    assert hashlib.sha256(ppm).hexdigest() == str(g["ppm_sha256"])
