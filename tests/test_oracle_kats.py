"""Known-answer tests from the reference's dormant self-tests (reference main.cpp:142-346; their
calls are commented out at main.cpp:120-122) against the oracle AND against the host+device
inline code the kernels are built from (through tests/emu)."""
import ctypes as C

import numpy as np

import helpers as H

# huffmanTreeTest(), main.cpp:142-203: hand-written table, codes of length 1..16
KAT_COUNTS = [0, 2, 1, 3, 3, 1, 0, 0, 0, 3, 2, 0, 1, 0, 2, 1]
KAT_SYMBOLS = [0x01, 0x02, 0x03, 0x11, 0x04, 0x00, 0x05, 0x21, 0x12, 0x07, 0xA0, 0xA1, 0xA3, 0xC3, 0x14, 0x27, 0x3A,
               0x4A, 0x56]
# canonical codes the reference tree assigns (verified by running HuffmanTree on this table)
KAT_CODES = {0x01: "00", 0x02: "01", 0x03: "100", 0x11: "1010", 0x04: "1011", 0x00: "1100", 0x05: "11010",
             0x21: "11011", 0x12: "11100", 0x07: "111010", 0xA0: "1110110000", 0xA1: "1110110001", 0xA3: "1110110010",
             0xC3: "11101100110", 0x14: "11101100111", 0x27: "1110110100000", 0x3A: "111011010000100",
             0x4A: "111011010000101", 0x56: "1110110100001100"}
# contains() queries of main.cpp:196-202 -> symbol, or None for "" (not a code)
KAT_QUERIES = [("100", 3), ("1100", 0), ("101", None), ("1" * 16, None), ("111010", 7), ("111011010000101", 74),
               ("1110110100001100", 86)]


def _tbl():
    return (np.array(KAT_COUNTS, dtype=np.uint8), np.array(KAT_SYMBOLS + [0] * (256 - len(KAT_SYMBOLS)), dtype=np.uint8))


def test_huffman_canonical_codes_oracle():
    counts, syms = _tbl()
    codes = np.zeros(256, dtype=np.uint16)
    lens = np.zeros(256, dtype=np.uint8)
    n = H.oracle().kpo_huff_codes(counts.ctypes.data, syms.ctypes.data, codes.ctypes.data, lens.ctypes.data)
    assert n == len(KAT_SYMBOLS)
    for i, s in enumerate(KAT_SYMBOLS):
        assert format(int(codes[i]), "b").zfill(int(lens[i])) == KAT_CODES[s]


def test_huffman_contains_oracle():
    counts, syms = _tbl()
    for bits, want in KAT_QUERIES:
        got = H.oracle().kpo_huff_lookup(counts.ctypes.data, syms.ctypes.data, bits.encode())
        assert got == (-1 if want is None else want), bits


def _window(bits: str) -> int:
    # left-align the code in a 32-bit window and pad with the complement of its last bit so that a
    # longer code cannot complete by accident where the KAT expects "not a code"
    return int(bits.ljust(32, "0"), 2)


def test_huffman_lut_kernel_code():
    """The two-level LUT the kernels use resolves every code of the KAT table to (length, symbol)."""
    counts, syms = _tbl()
    emu = H.emu()
    for s, code in KAT_CODES.items():
        for pad in ("0", "1"):
            win = int(code.ljust(32, pad), 2)
            for is_ac in (0, 1):
                e = emu.emu_huff_lookup(counts.ctypes.data, syms.ctypes.data, is_ac, win, 0)
                ec = emu.emu_huff_lookup(counts.ctypes.data, syms.ctypes.data, is_ac, win, 1)
                assert e == ec, "LUT path and canonical search disagree"
                assert ((e >> 5) & 15) == (s & 15)
                assert (e & 31) - ((e >> 5) & 15) == len(code)  # bits 0-4: code length + magnitude bits
                adv = e >> 9
                assert adv == (1 if not is_ac else (64 if s == 0 else (s >> 4) + 1))
    # "1111111111111111" is no code: flagged as needing more than 16 bits
    e = emu.emu_huff_lookup(counts.ctypes.data, syms.ctypes.data, 1, 0xFFFFFFFF, 0)
    assert (e & 31) == 17


def test_huffman_lut_exhaustive_annex_k():
    """Every 16-bit window: LUT path == canonical search, for the four Annex K tables (as found in
    lena.jpg's DHT segments)."""
    import libkpeg_b200 as K
    data = (H.GOLDEN / "lena.jpg").read_bytes()
    plan, _, _ = K.parse_jfif(data)
    emu = H.emu()
    rng = np.random.default_rng(1)
    for tc in (0, 1):
        for th in (0, 1):
            spec = plan.ht[tc][th]
            counts = np.ctypeslib.as_array(spec.counts).copy()
            syms = np.ctypeslib.as_array(spec.symbols).copy()
            for w16 in range(0, 65536, 7):
                win = (w16 << 16) | int(rng.integers(0, 65536))
                assert emu.emu_huff_lookup(counts.ctypes.data, syms.ctypes.data, tc, win, 0) == \
                    emu.emu_huff_lookup(counts.ctypes.data, syms.ctypes.data, tc, win, 1)


def test_extend():
    # bitStringtoValue, Image.cpp:285-302: "" -> 0; leading 1 -> value; leading 0 -> -(complement)
    o = H.oracle()
    assert o.kpo_extend(0, 0) == 0
    assert o.kpo_extend(1, 1) == 1 and o.kpo_extend(0, 1) == -1
    assert o.kpo_extend(0b101, 3) == 5 and o.kpo_extend(0b010, 3) == -5
    assert o.kpo_extend(0b0000000000, 10) == -1023 and o.kpo_extend(0b1111111111, 10) == 1023


# transformTest(), main.cpp:205-250: the Wikipedia JPEG example block
WIKI_BLOCK = np.array([[52, 55, 61, 66, 70, 61, 64, 73], [63, 59, 55, 90, 109, 85, 69, 72],
                       [62, 59, 68, 113, 144, 104, 66, 73], [63, 58, 71, 122, 154, 106, 70, 69],
                       [67, 61, 68, 104, 126, 88, 68, 70], [79, 65, 60, 70, 77, 68, 58, 75],
                       [85, 71, 64, 59, 55, 61, 65, 83], [87, 79, 69, 68, 65, 76, 78, 94]], dtype=np.int32)


def _fdct(block):
    x = np.arange(8)
    cs = np.cos((2 * x[:, None] + 1) * x[None, :] * np.pi / 16.0)  # [x][u]
    cu = np.where(x == 0, 1 / np.sqrt(2), 1.0)
    return 0.25 * cu[:, None] * cu[None, :] * (cs.T @ (block - 128.0) @ cs)


def test_idct_wikipedia_block():
    """DCTTest -> IDCTTest round trip (main.cpp:252-326): FDCT rounded to 0.01 starts
    -415.38 -30.19 -61.20 27.24 56.12 -20.10 -2.39 0.46 / 4.47 -21.86 ...; an integer-rounded
    version of it pushed through the reference's computeIDCT restatement gives back the block."""
    F = _fdct(WIKI_BLOCK)
    assert abs(F[0, 0] - (-415.375)) < 1e-6
    np.testing.assert_allclose(np.round(F[0], 2), [-415.38, -30.19, -61.20, 27.24, 56.12, -20.10, -2.39, 0.46], atol=0.011)
    np.testing.assert_allclose(np.round(F[1, :2], 2), [4.47, -21.86], atol=0.011)
    Fi = np.rint(F).astype(np.int32)
    out = np.zeros(64, dtype=np.float32)
    H.oracle().kpo_idct8x8(np.ascontiguousarray(Fi.reshape(-1)).ctypes.data, out.ctypes.data)
    rec = np.array([H.oracle().kpo_level_shift(float(v)) for v in out]).reshape(8, 8)
    assert np.abs(rec - WIKI_BLOCK).max() <= 1  # coefficients were rounded to integers


def test_colour_kat():
    # colorTest(), main.cpp:328-346: (Y,Cb,Cr) = (383,128,128) -> (255,255,255): floor, THEN clamp,
    # and Y may exceed 255 before the clamp
    rgb = (C.c_int * 3)()
    H.oracle().kpo_ycbcr_to_rgb(383, 128, 128, rgb)
    assert list(rgb) == [255, 255, 255]
    H.emu().emu_colour(383 - 128, 0, 0, rgb)
    assert list(rgb) == [255, 255, 255]


def test_zigzag_table():
    # Transform.cpp:5-27 zzOrderToMatIndices: first entries (0,0) (0,1) (1,0) (2,0) (1,1) (0,2) ... last (7,7)
    want = [(0, 0), (0, 1), (1, 0), (2, 0), (1, 1), (0, 2), (0, 3), (1, 2), (2, 1), (3, 0), (4, 0)]
    r, c = C.c_int(), C.c_int()
    seen = set()
    for i in range(64):
        H.oracle().kpo_zigzag_to_rc(i, C.byref(r), C.byref(c))
        if i < len(want):
            assert (r.value, c.value) == want[i]
        assert H.emu().emu_zigzag(i) == r.value * 8 + c.value
        seen.add((r.value, c.value))
    assert len(seen) == 64 and (r.value, c.value) == (7, 7)


def test_ppm_header_bytes():
    import libkpeg_b200 as K
    want = b"P6\n# PPM dump created using libKPEG: https://github.com/TheIllusionistMirage/libKPEG\n512 512\n255\n"
    assert len(want) == 97
    assert K.ppm_header(512, 512) == want
    buf = C.create_string_buffer(200)
    n = H.oracle().kpo_ppm_header(512, 512, buf, 200)
    assert buf.raw[:n] == want
