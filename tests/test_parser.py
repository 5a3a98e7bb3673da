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


def test_restart_bands_balanced_in_cut_units():
    """DRI spanning more than one MCU row, parts close to the number of cut units, my % row_step != 0:
    every band must begin on a distinct cut point and the bands must reassemble the image."""
    from libkpeg_b200.shard import split_restart_bands
    from libkpeg_b200.synth import EMIT_RESTART, QUIRK_FREE, SynthParams, synth_encode
    # 16x40: mx = 2, my = 5; DRI = 4 MCUs = 2 MCU rows -> row_step = 2, units = 3
    jpg = synth_encode(SynthParams(16, 40, quality=90, restart_interval=4, flags=QUIRK_FREE | EMIT_RESTART, seed=5,
                                   file_components=1))
    plan, off, n = K.parse_jfif(jpg)
    scan = jpg[off:off + n]
    whole = H.emu_decode(jpg)["pixels"]
    for parts in (1, 2, 3, 4, 7):
        bands = split_restart_bands(plan, scan, parts)
        rows = [b.row0 for b in bands if b.rows]
        assert rows == sorted(set(rows)) and rows[0] == 0 and all(r % 16 == 0 for r in rows), (parts, rows)
        assert sum(b.rows for b in bands) == 40
        for b in bands:
            if not b.rows:
                assert b.scan.size == 0
                continue
            got = H.emu_decode(None, scans=[b.scan], plan=b.plan)["pixels"][0]
            assert np.array_equal(got, whole[b.row0:b.row0 + b.rows]), (parts, b.row0)


def test_scan_ends_at_first_eoi(lena_jpg):
    """Files with data after the first EOI that also END in FF D9 (concatenated JPEGs, MPO): the scan stops at the
    first EOI, as the reference's scanImageData does (Decoder.cpp:546-557)."""
    plan, off, n = K.parse_jfif(lena_jpg + lena_jpg)
    assert n == 104113
    plan, off, n = K.parse_jfif(lena_jpg + b"\x00\x01\xff\xd9")
    assert n == 104113


def test_distinct_tables_per_component_cpu():
    """Three quantiser tables and a Huffman table pair of its own per component, table / component ids unlike the
    0/1/1 mapping the reference hard-wires: kernel logic (CPU single-stepper) == oracle, and the result
    differs from a decode that ignores the selectors."""
    import jpeg_writer as JW
    from libkpeg_b200.synth import QUIRK_FREE, SynthParams, synth_encode
    base = synth_encode(SynthParams(96, 64, quality=88, seed=21, flags=QUIRK_FREE)).tobytes()
    jpg = JW.distinct_tables_variant(base, H.oracle_decode, K.parse_jfif)
    plan, off, n = K.parse_jfif(jpg)
    assert tuple(plan.comp_tq) == (2, 0, 1) and tuple(plan.comp_td) == (1, 0, 2) and tuple(plan.comp_ta) == (2, 0, 1)
    o = H.oracle_decode(jpg)
    assert np.array_equal(o["coef"], H.oracle_decode(base)["coef"])  # same coefficients, other tables
    assert not np.array_equal(o["pixels"], H.oracle_decode(base)["pixels"])
    for sb in (64, 512):
        e = H.emu_decode(jpg, sub_bits=sb)
        assert e["status"] == 0 and e["records_ok"]
        assert np.array_equal(o["coef"], e["coef"]) and np.array_equal(o["pixels"], e["pixels"])
