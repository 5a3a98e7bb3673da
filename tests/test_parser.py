"""Host-side container parse (kpeg_parse_jfif) on real-world variants, with PIL as the producer.
The reference's own parser handles only SOI/APP0/COM/DQT/SOF0/DHT/SOS; the
drop-in is a T.81-correct superset that must agree with the oracle on everything it accepts."""
import io

import numpy as np
import pytest

import helpers as H
import libkpeg_b200 as K
from libkpeg_b200 import api

PIL = pytest.importorskip("PIL.Image")


def pil_jpeg(w=64, h=48, mode="RGB", **kw):
    rng = np.random.default_rng(w * h)
    base = rng.integers(0, 255, size=(h // 8 + 1, w // 8 + 1, 3)).astype(np.uint8)
    arr = np.kron(base, np.ones((8, 8, 1), dtype=np.uint8))[:h, :w]
    img = PIL.fromarray(arr, "RGB").convert(mode)
    buf = io.BytesIO()
    img.save(buf, "JPEG", **kw)
    return buf.getvalue()


def test_pil_444_standard_and_optimised_tables():
    for kw in (dict(quality=90, subsampling=0), dict(quality=90, subsampling=0, optimize=True),
               dict(quality=75, subsampling=0, comment=b"hello")):
        jpg = pil_jpeg(**kw)
        plan, off, n = K.parse_jfif(jpg)
        assert (plan.width, plan.height, plan.ncomp) == (64, 48, 3)
        o = H.oracle_decode(jpg)
        e = H.emu_decode(jpg)
        assert e["status"] == 0 and np.array_equal(o["coef"], e["coef"]) and np.array_equal(o["pixels"], e["pixels"])


def test_pil_restart_markers_and_exif():
    jpg = pil_jpeg(quality=85, subsampling=0, restart_marker_blocks=4)
    plan, off, n = K.parse_jfif(jpg)
    assert plan.restart_interval == 4
    o, e = H.oracle_decode(jpg), H.emu_decode(jpg, sub_bits=128)
    assert e["status"] == 0 and np.array_equal(o["pixels"], e["pixels"])
    # APP1 (EXIF-like) segment: the reference FATALs, T.81 says skip by length
    base = pil_jpeg(quality=85, subsampling=0)
    app1 = b"\xff\xe1" + (2 + 10).to_bytes(2, "big") + b"Exif\0\0abcd"
    withapp = base[:2] + app1 + base[2:]
    p2, _, _ = K.parse_jfif(withapp)
    assert (p2.width, p2.height) == (plan.width, plan.height)
    assert np.array_equal(H.emu_decode(withapp)["pixels"], H.emu_decode(base)["pixels"])


def test_true_grayscale():
    jpg = pil_jpeg(mode="L", quality=90)
    plan, _, _ = K.parse_jfif(jpg)
    assert plan.ncomp == 1
    o, e = H.oracle_decode(jpg), H.emu_decode(jpg)
    assert np.array_equal(o["pixels"], e["pixels"])


@pytest.mark.parametrize("kw,code", [(dict(quality=90, subsampling=2), api.KPEG_ERR_UNSUPPORTED),  # 4:2:0 -> TERMINATE
                                     (dict(quality=90, subsampling=0, progressive=True), api.KPEG_ERR_UNSUPPORTED)])
def test_unsupported_codings(kw, code):
    with pytest.raises(K.KpegError) as e:
        K.parse_jfif(pil_jpeg(**kw))
    assert e.value.code == code


def test_malformed_containers(lena_jpg):
    for bad in (b"", b"\xff\xd8", lena_jpg[:300], b"\x00" + lena_jpg, lena_jpg[:2] + b"\x12\x34" + lena_jpg[2:]):
        with pytest.raises(K.KpegError) as e:
            K.parse_jfif(bad)
        assert e.value.code in (api.KPEG_ERR_FORMAT, api.KPEG_ERR_UNSUPPORTED)
    # trailing bytes after EOI: the scan still ends at the marker
    plan, off, n = K.parse_jfif(lena_jpg + b"\x00\x01\x02")
    assert n == 104113
