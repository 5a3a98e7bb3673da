"""GPU tests of the multi-device entry points, the GPU-side output writers and the host-buffer batch path.

The tiled decode takes a device LIST; listing device 0 several times runs the whole multi-band machinery (band split,
one host thread and one pooled context per band, host gather / peer gather) on a single-GPU box, which is what the
driver's test run has.  With several GPUs visible the same tests also spread the bands over all of them."""
import ctypes as C

import numpy as np
import pytest

import helpers as H
import libkpeg_b200 as K
from libkpeg_b200 import api
from libkpeg_b200.synth import EMIT_RESTART, GRAY_CONTENT, QUIRK_FREE, SynthParams, synth_encode

pytestmark = pytest.mark.gpu


def _ndev():
    return api.load_cuda_library().kpeg_cuda_device_count()


def _banded_jpg(w=512, h=384, ri_rows=1, q=90, seed=3):
    return synth_encode(SynthParams(width=w, height=h, quality=q, restart_interval=(w // 8) * ri_rows,
                                    flags=QUIRK_FREE | EMIT_RESTART, seed=seed)).tobytes()


@pytest.mark.parametrize("devices", [[0], [0, 0], [0, 0, 0], "all"])
def test_tiled_decode_equals_single_device(decoder, devices):
    """Row J of the scope table: restart-interval bands of one image, one per listed device, gathered into ONE host
    frame (placement of reference src/Image.cpp:51-70) == the whole image decoded by one GPU == the oracle."""
    if devices == "all":
        devices = list(range(_ndev())) * (2 if _ndev() == 1 else 1)
    jpg = _banded_jpg(w=640, h=424)  # 53 MCU rows: bands of unequal height
    whole = decoder.decode_file(jpg)
    frame = api.PinnedArray(whole.nbytes)
    got, st = api.decode_file_tiled(devices, jpg, out=frame.array)
    assert np.array_equal(got, whole)
    ref = H.oracle_decode(jpg, parity=True)
    assert np.array_equal(got, ref["pixels"])
    assert st.segments >= 53 and st.kernel_launches >= 10 * min(len(devices), 53)
    frame.free()


def test_tiled_decode_without_restart_markers_does_not_shard(decoder):
    jpg = synth_encode(SynthParams(width=256, height=128, quality=90, seed=9, flags=QUIRK_FREE)).tobytes()
    got, st = api.decode_file_tiled([0, 0], jpg)
    assert np.array_equal(got, decoder.decode_file(jpg))
    assert st.kernel_launches <= 12  # one band: one kernel sequence


def test_tiled_decode_more_devices_than_rows(decoder):
    jpg = _banded_jpg(w=64, h=16)  # two MCU rows
    got, _ = api.decode_file_tiled([0] * 5, jpg)
    assert np.array_equal(got, decoder.decode_file(jpg))


def test_tiled_decode_reports_corrupt_band(decoder):
    jpg = bytearray(_banded_jpg(w=256, h=256))
    _, off, n = K.parse_jfif(bytes(jpg))
    jpg[off + n // 2 + 5] = 0xFF
    jpg[off + n // 2 + 6] = 0xC4  # a marker that may not appear inside a scan
    with pytest.raises(K.KpegError):
        api.decode_file_tiled([0, 0], bytes(jpg))


def test_tiled_decode_into_device_frame_peer_gather(decoder):
    """The bands gathered in DEVICE memory of one GPU (peer copies; in place for the band that GPU decodes itself)."""
    lib = api.load_cuda_library()
    nd = _ndev()
    devices = [0, 0, 0] if nd == 1 else list(range(nd))
    jpg = _banded_jpg(w=512, h=256)
    whole = decoder.decode_file(jpg)
    plan, off, n = K.parse_jfif(jpg)
    plan.flags = K.KPEG_FLAG_REF_PARITY
    scan = np.frombuffer(jpg, dtype=np.uint8)[off:off + n].copy()
    d_frame = decoder.device_alloc(whole.nbytes + 64)
    dv = (C.c_int * len(devices))(*devices)
    st = api.Stats()
    rc = lib.kpeg_cuda_decode_tiled_device(dv, len(devices), C.byref(plan), scan.ctypes.data, scan.size, 0, d_frame, C.byref(st))
    assert rc == 0, lib.kpeg_tiled_last_error()
    got = np.empty_like(whole)
    decoder.d2h(got, d_frame)
    assert np.array_equal(got, whole)
    decoder.device_free(d_frame)


def test_context_pool_reuses_contexts():
    lib = api.load_cuda_library()
    lib.kpeg_cuda_pool_clear()  # contexts the tiled decodes of other tests left idle
    a, b = C.c_void_p(), C.c_void_p()
    assert lib.kpeg_cuda_acquire(0, C.byref(a)) == 0
    lib.kpeg_cuda_release(0, a)
    assert lib.kpeg_cuda_acquire(0, C.byref(b)) == 0
    assert a.value == b.value  # the same context came back, with its scratch memory
    lib.kpeg_cuda_release(0, b)
    lib.kpeg_cuda_pool_clear()


@pytest.mark.parametrize("gray", [False, True])
def test_ppm_written_on_the_gpu(decoder, gray):
    """GPU-side PPM writer: the bytes Image::dumpRawData writes (reference src/Image.cpp:108-140), assembled in HBM."""
    kw = dict(file_components=1, flags=QUIRK_FREE | GRAY_CONTENT) if gray else dict(flags=QUIRK_FREE)
    jpg = synth_encode(SynthParams(width=200, height=120, quality=90, seed=21, **kw)).tobytes()
    plan, off, n = K.parse_jfif(jpg)
    plan.flags = K.KPEG_FLAG_REF_PARITY
    scan = np.frombuffer(jpg, dtype=np.uint8)[off:off + n].copy()
    d_scan = decoder.device_alloc(scan.size + 64)
    cap = 200 * 120 * 3 + 256
    d_out = decoder.device_alloc(cap)
    decoder.h2d(d_scan, scan)
    o, ln = decoder.decode_ppm_device(plan, d_scan, scan.size, d_out, cap)
    buf = np.empty(cap, dtype=np.uint8)
    decoder.d2h(buf, d_out)
    px = decoder.decode_file(jpg)
    rgb = np.repeat(px[:, :, None], 3, axis=2) if gray else px
    expect = K.ppm_header(200, 120) + rgb.tobytes()
    assert o < 16 and ln == len(expect)
    assert buf[o:o + ln].tobytes() == expect
    decoder.device_free(d_scan)
    decoder.device_free(d_out)


def test_planar_writer(decoder):
    rng = np.random.default_rng(5)
    for n in (4096, 1001):
        rgb = rng.integers(0, 256, size=(n, 3), dtype=np.uint8)
        d_in, d_out = decoder.device_alloc(n * 3 + 64), decoder.device_alloc(n * 3 + 64)
        decoder.h2d(d_in, rgb)
        decoder.interleaved_to_planar(d_in, d_out, n)
        got = np.empty((3, n), dtype=np.uint8)
        decoder.d2h(got, d_out)
        assert np.array_equal(got, rgb.T)
        decoder.device_free(d_in)
        decoder.device_free(d_out)


def test_host_batch_small_and_large_scans_contiguous_outputs(decoder):
    """kpeg_cuda_submit_batch: small scans travel through pinned staging as one copy, large ones straight from the
    caller's buffer, the separators are written by a kernel, contiguous outputs leave as one copy per chunk."""
    w, h = 128, 128
    jpgs = [synth_encode(SynthParams(width=w, height=h, quality=90, seed=200 + i, noise_amp=(0, 0, 200, 0, 200, 200, 0)[i % 7],
                                     flags=QUIRK_FREE)).tobytes() for i in range(23)]
    parsed = [K.parse_jfif(j) for j in jpgs]
    plan = parsed[0][0]
    plan.flags = K.KPEG_FLAG_REF_PARITY
    scans = [np.frombuffer(j, dtype=np.uint8)[o:o + n].copy() for j, (_, o, n) in zip(jpgs, parsed)]
    sizes = sorted(s.size for s in scans)
    assert sizes[0] < 24 * 1024 < sizes[-1], sizes  # both sides of the staging threshold
    want = [decoder.decode_file(j) for j in jpgs]
    frame = api.PinnedArray(len(jpgs) * w * h * 3)
    outs = [frame.array[i * w * h * 3:(i + 1) * w * h * 3].reshape(h, w, 3) for i in range(len(jpgs))]
    decoder.decode_batch(plan, scans, outs)
    for i, (a, b) in enumerate(zip(outs, want)):
        assert np.array_equal(a, b), f"image {i} (contiguous outputs)"
    # scans back to back in ONE host arena: a single copy in, repacked on the device
    arena = api.PinnedArray(sum(sc.size for sc in scans))
    views, o = [], 0
    for sc in scans:
        arena.array[o:o + sc.size] = sc
        views.append(arena.array[o:o + sc.size])
        o += sc.size
    frame.array[:] = 0
    handle = decoder.prepare_batch(views, outs)
    decoder.submit_prepared(plan, handle)
    decoder.wait()
    for i, (a, b) in enumerate(zip(outs, want)):
        assert np.array_equal(a, b), f"image {i} (contiguous inputs)"
    arena.free()
    outs2 = [np.empty((h, w, 3), dtype=np.uint8) for _ in jpgs]  # separate buffers: one copy per image
    decoder.decode_batch(plan, scans, outs2)
    for i, (a, b) in enumerate(zip(outs2, want)):
        assert np.array_equal(a, b), f"image {i} (separate outputs)"
    frame.free()


def test_large_restart_image_goes_through_the_lanes_in_bands(monkeypatch):
    """kpeg_cuda_decode on a large (>= 64 MB of pixels) restart-marked image from host memory cuts it into four bands that
    run on the context's lanes with their copies overlapping (decode_banded: cut at byte positions, the restart markers
    of every band counted on the GPU to learn its rows); the frame must equal the one a context with KPEG_BANDS=1 (whole
    image, one lane) produces, a corrupt band must surface as an error, and an image whose restart interval is not a whole
    number of MCU rows must still decode (whole)."""
    w, h = 8192, 2744  # 67 MB of pixels; 343 MCU rows: bands of unequal height
    jpg = _banded_jpg(w=w, h=h, ri_rows=1, q=60, seed=21)
    banded = K.Decoder(device=0)
    monkeypatch.setenv("KPEG_BANDS", "1")
    whole = K.Decoder(device=0)
    monkeypatch.delenv("KPEG_BANDS")
    try:
        a = banded.decode_file(jpg)
        launches_banded = banded.last_stats.kernel_launches
        b = whole.decode_file(jpg)
        assert np.array_equal(a, b)
        assert launches_banded >= 3 * whole.last_stats.kernel_launches  # four kernel sequences, not one
        assert banded.last_stats.height == h and banded.last_stats.segments >= 343
        # (content against the oracle: test_config4_full_size_restart_bands and the small-image tests; the oracle on an
        # image of this size takes minutes)
        bad = bytearray(jpg)
        _, off, n = K.parse_jfif(bytes(bad))
        bad[off + n // 2: off + n // 2 + 64] = bytes(64)  # a run of zero bytes inside the third band
        with pytest.raises(K.KpegError):
            banded.decode_file(bytes(bad))
        assert np.array_equal(banded.decode_file(jpg), b)  # the context is usable afterwards
        # two restart intervals per MCU row, and rows of unequal band sizes with a two-row interval
        two = _banded_jpg(w=w, h=h, ri_rows=2, q=40, seed=8)
        assert np.array_equal(banded.decode_file(two), whole.decode_file(two))
        # restart interval of 5 MCUs: not a whole number of rows -> decoded whole
        odd = synth_encode(SynthParams(width=w, height=h, quality=30, restart_interval=5, flags=QUIRK_FREE | EMIT_RESTART, seed=4)).tobytes()
        assert np.array_equal(banded.decode_file(odd), whole.decode_file(odd))
    finally:
        banded.close()
        whole.close()


def test_decode_files_takes_any_mix_of_files(decoder):
    """kpeg_cuda_decode_files: files of different sizes, qualities (tables), component counts and codings in ONE call;
    files with identical plans share a batch; a truncated file and a non-JPEG get their own error codes and do not stop
    the others.  Every decoded image equals the oracle's."""
    from libkpeg_b200.synth import NON_INTERLEAVED
    lena = (H.ROOT / "tests" / "golden" / "lena.jpg").read_bytes() if hasattr(H, "ROOT") else None
    files = []
    for seed in (1, 2, 3):  # one plan, three files
        files.append(synth_encode(SynthParams(64, 48, quality=80, seed=seed)).tobytes())
    files.append(synth_encode(SynthParams(120, 80, quality=50, seed=4)).tobytes())
    files.append(synth_encode(SynthParams(120, 80, quality=95, restart_interval=7, flags=QUIRK_FREE | EMIT_RESTART, seed=5)).tobytes())
    files.append(synth_encode(SynthParams(97, 33, file_components=1, quality=70, flags=QUIRK_FREE | GRAY_CONTENT, seed=6)).tobytes())
    files.append(synth_encode(SynthParams(88, 56, quality=70, seed=7, flags=QUIRK_FREE | NON_INTERLEAVED)).tobytes())
    if lena:
        files.append(lena)
    good = len(files)
    _, off, n = K.parse_jfif(files[0])
    files.append(files[0][:off + n // 2] + b"\xff\xd9")  # same plan as the first three, truncated
    files.append(b"not a jpeg at all")
    imgs, codes = decoder.decode_files(files)
    assert codes[:good] == [0] * good and codes[good] == api.KPEG_ERR_STREAM and codes[good + 1] == api.KPEG_ERR_FORMAT
    assert imgs[good] is None and imgs[good + 1] is None
    for i in range(good):
        ref = H.oracle_decode(files[i], parity=True)["pixels"]
        assert np.array_equal(imgs[i], ref), f"file {i}"
    # all good: one kernel sequence per distinct plan (5 plans + lena), not one per file
    imgs2, codes2 = decoder.decode_files(files[:good])
    assert codes2 == [0] * good and all(np.array_equal(a, b) for a, b in zip(imgs2, imgs[:good]))


@pytest.mark.parametrize("ndev,ri_rows", [(2, 1), (3, 1), (5, 2)])
def test_tiled_decode_of_a_large_scan_cuts_at_byte_positions(decoder, ndev, ri_rows):
    """Scans of 1 MB and more with a restart interval of whole MCU rows are cut at byte positions and the rows of every band
    follow from its marker count (counted by one host thread per band): bands of unequal, data-dependent height."""
    jpg = _banded_jpg(w=2048, h=1400, ri_rows=ri_rows, q=92, seed=12)
    _, off, n = K.parse_jfif(jpg)
    assert n >= 1 << 20
    whole = decoder.decode_file(jpg)
    got, st = api.decode_file_tiled([0] * ndev, jpg)
    assert np.array_equal(got, whole)
    assert st.kernel_launches >= 10 * ndev
    # a band that lost a restart marker: the rows no longer add up -> an error, not a shifted image
    bad = bytearray(jpg)
    k = jpg.index(b"\xff\xd3", off + n // 3)
    bad[k:k + 2] = b"\x00\x00"
    with pytest.raises(K.KpegError):
        api.decode_file_tiled([0] * ndev, bytes(bad))


def test_decode_files_many_files_few_plans(decoder):
    """120 files in six plans (sizes x qualities), shuffled: six batches, every image equal to its single-file decode."""
    rng = np.random.default_rng(77)
    specs = [(64, 48, 80), (64, 48, 60), (96, 64, 80), (40, 40, 90), (128, 24, 70), (57, 33, 50)]
    files = []
    for i in range(120):
        w, h, q = specs[int(rng.integers(0, len(specs)))]
        files.append(synth_encode(SynthParams(w, h, quality=q, seed=1000 + i)).tobytes())
    imgs, codes = decoder.decode_files(files)
    assert codes == [0] * len(files)
    assert decoder.last_stats.kernel_launches <= 12 * len(specs) + 12  # one kernel sequence per plan, not per file
    for i in (0, 17, 63, 119):
        assert np.array_equal(imgs[i], decoder.decode_file(files[i])), f"file {i}"
    ref = H.oracle_decode(files[5], parity=True)["pixels"]
    assert np.array_equal(imgs[5], ref)
