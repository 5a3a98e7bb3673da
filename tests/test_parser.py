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


# ---- frames coded one scan per component (T.81 A.2.3) ------------------------------------------------
def _noninterleaved(w=72, h=40, ri=0, q=80, seed=5, flags=0):
    from libkpeg_b200.synth import EMIT_RESTART, NON_INTERLEAVED, QUIRK_FREE, SynthParams, synth_encode
    fl = QUIRK_FREE | flags | (EMIT_RESTART if ri else 0)
    base = dict(width=w, height=h, quality=q, restart_interval=ri, seed=seed)
    return (synth_encode(SynthParams(**base, flags=fl)).tobytes(),
            synth_encode(SynthParams(**base, flags=fl | NON_INTERLEAVED)).tobytes())


def test_one_scan_per_component_is_parsed_scan_by_scan():
    inter, split = _noninterleaved(ri=4)
    with pytest.raises(K.KpegError) as e:
        K.parse_jfif(split)  # the single-scan entry point does not take it
    assert e.value.code == api.KPEG_ERR_UNSUPPORTED
    frame, scans = api.parse_jfif_scans(split)
    assert (frame.width, frame.height, frame.ncomp, len(scans)) == (72, 40, 3, 3)
    fi, si = api.parse_jfif_scans(inter)
    assert len(si) == 1 and si[0].plan.ncomp == 3 and bytes(si[0].plan) == bytes(fi)
    for k, sc in enumerate(scans):
        assert sc.plan.ncomp == 1 and sc.comp[0] == k and sc.plan.restart_interval == 4
        assert (sc.plan.width, sc.plan.height) == (72, 40)
        assert sc.plan.comp_td[0] == sc.plan.comp_ta[0] == (0 if k == 0 else 1)
        assert split[sc.off - 10:sc.off - 8] == b"\xff\xda" and sc.len > 0
        # the quantiser of component k sits in slot k of the frame plan
        src = fi.qt[0 if k == 0 else 1]
        assert list(frame.qt[k]) == list(src) and frame.comp_tq[k] == k


def test_one_scan_per_component_equals_the_interleaved_twin_in_the_oracle():
    """Twin streams: the same quantised coefficients as one interleaved scan (which the reference decodes and the oracle
    is pinned on) and as three scans -- the oracle must give the same coefficients and pixels for both."""
    for kw in (dict(), dict(w=57, h=33, q=40), dict(w=120, h=80, ri=5, q=75), dict(w=33, h=17, ri=3, q=20)):
        inter, split = _noninterleaved(**kw)
        a, b = H.oracle_decode(inter, parity=True), H.oracle_decode(split, parity=True)
        assert np.array_equal(a["coef"], b["coef"]) and np.array_equal(a["pixels"], b["pixels"])
    pil = np.asarray(PIL.open(io.BytesIO(split)).convert("RGB")).astype(int)  # an independent decoder accepts the file
    assert np.abs(pil - b["pixels"].astype(int)).max() <= 4


def test_multi_scan_files_the_path_does_not_take():
    _, split = _noninterleaved()
    frame, scans = api.parse_jfif_scans(split)
    # the file ends after the second scan
    cut = split[:scans[2].off - 10] + b"\xff\xd9"
    with pytest.raises(K.KpegError) as e:
        api.parse_jfif_scans(cut)
    assert e.value.code == api.KPEG_ERR_FORMAT
    # the second scan codes component 1 again ... (component id lives 5 bytes into the SOS payload)
    twice = bytearray(split)
    twice[scans[2].off - 5] = 2
    with pytest.raises(K.KpegError) as e:
        api.parse_jfif_scans(bytes(twice))
    assert e.value.code == api.KPEG_ERR_UNSUPPORTED
    # a two-component scan
    import jpeg_writer as JW
    two = bytearray(split)
    sos = scans[0].off - 10
    two[sos:scans[0].off] = JW._seg(0xDA, bytes([2, 1, 0x00, 2, 0x11, 0, 63, 0]))
    with pytest.raises(K.KpegError) as e:
        api.parse_jfif_scans(bytes(two))
    assert e.value.code == api.KPEG_ERR_UNSUPPORTED


def test_end_of_a_large_scan_is_found_by_the_parallel_walk():
    """Scans of 4 MB and more are walked by several host threads, each over its own stretch: the end must be where the
    sequential walk puts it -- the first FF that is not followed by 00, D0..D7 or FF -- wherever it falls (stretch
    boundaries included), with stuffed FFs, restart markers and fill bytes on both sides of every boundary."""
    from libkpeg_b200.synth import QUIRK_FREE, SynthParams, synth_encode
    small = synth_encode(SynthParams(16, 16, quality=50, seed=1, flags=QUIRK_FREE)).tobytes()
    _, off, n = K.parse_jfif(small)
    head = small[:off]
    rng = np.random.default_rng(5)
    size = 6 * 1024 * 1024 + 123
    body = rng.integers(0, 255, size=size, dtype=np.uint8)  # no FF at all
    pos = rng.choice(size - 4, size=40000, replace=False)
    pos.sort()
    pos = pos[np.diff(pos, prepend=-10) > 3]
    kind = rng.integers(0, 3, size=pos.size)
    for p, k in zip(pos, kind):
        body[p] = 0xFF
        body[p + 1] = (0x00, 0xD0 + (int(p) & 7), 0xFF)[int(k)]
        if k == 2:
            body[p + 2] = 0x00  # FF FF 00: a fill byte, then a stuffed FF
    nthreads = 8
    span = (size + nthreads - 1) // nthreads
    ends = [size - 2, size // 2 + 17, 5] + [t * span + d for t in range(1, nthreads) for d in (-2, -1, 0, 1)]
    for e in ends:
        buf = body.copy()
        buf[e] = 0xFF
        buf[e + 1] = 0xD9
        if buf[e - 1] == 0xFF:  # do not turn the byte before into the first half of another pair
            buf[e - 1] = 0x01
        file = head + buf.tobytes()
        _, o2, n2 = K.parse_jfif(file)
        # the sequential definition, restricted to the bytes before e (nothing there may terminate the scan)
        assert o2 == off and n2 == e, f"end at {e}: found {n2}"


def test_bands_cut_at_byte_positions_decode_to_the_same_frame():
    """kpeg_split_restart_bands_by_bytes: every band begins right after a restart marker, its rows are (markers inside
    it + 1) x rows per interval, the bands tile the frame -- and, decoded one by one (by the CPU single-stepper here, by
    the GPUs in tests/test_gpu_multi.py), they give the rows of the whole image."""
    from libkpeg_b200.shard import split_restart_bands
    from libkpeg_b200.synth import EMIT_RESTART, QUIRK_FREE, SynthParams, synth_encode
    for (w, h, ri_rows, parts) in [(256, 200, 1, 3), (128, 320, 2, 4), (64, 64, 1, 2)]:
        jpg = synth_encode(SynthParams(w, h, quality=85, restart_interval=(w // 8) * ri_rows, flags=QUIRK_FREE | EMIT_RESTART, seed=w + h)).tobytes()
        plan, off, n = K.parse_jfif(jpg)
        scan = np.frombuffer(jpg, dtype=np.uint8)[off:off + n]
        bands = split_restart_bands(plan, scan, parts, by_bytes=True)
        assert len(bands) == parts and bands[0].row0 == 0 and sum(b.rows for b in bands) == h
        whole = H.oracle_decode(jpg, parity=True)["pixels"]
        at = off
        for k, b in enumerate(bands):
            start = b.scan.ctypes.data - scan.ctypes.data
            if k:
                assert scan[start - 2] == 0xFF and 0xD0 <= scan[start - 1] <= 0xD7  # begins right after a marker
                assert start >= n * k // parts                                          # ... at or after k / parts of the scan
            body = bytes(b.scan)
            markers = sum(1 for i in range(len(body) - 1) if body[i] == 0xFF and 0xD0 <= body[i + 1] <= 0xD7)
            assert b.row0 % 8 == 0 and (b.rows == (markers + 1) * 8 * ri_rows or k == parts - 1)
            plan_b = b.plan
            plan_b.flags = 1
            e = H.emu_decode(None, scans=[np.frombuffer(body, dtype=np.uint8)], plan=plan_b)
            assert e["status"] == 0 and np.array_equal(e["pixels"][0], whole[b.row0:b.row0 + b.rows]), f"band {k}"
    # a restart interval that is not a whole number of MCU rows, or fewer than two intervals per band: not this splitter's case
    odd = synth_encode(SynthParams(128, 128, quality=70, restart_interval=5, flags=QUIRK_FREE | EMIT_RESTART, seed=3)).tobytes()
    plan, off, n = K.parse_jfif(odd)
    with pytest.raises(K.KpegError) as ex:
        split_restart_bands(plan, np.frombuffer(odd, dtype=np.uint8)[off:off + n], 2, by_bytes=True)
    assert ex.value.code == api.KPEG_ERR_UNSUPPORTED
