"""Regenerates the committed golden fixtures.  Run in the build container, where /root/reference
and the compiled unmodified reference (oracle/_ref/kpeg_ref_quiet, `make -C oracle`) exist:

    python tests/golden/make_golden.py

Outputs (all small, all committed):
  lena.jpg            the reference's only JPEG fixture (misc/images/lena.jpg), the config-1 input
  lena_ref.npz        pixel payload of the PPM the compiled reference writes for it + PPM header + sha256
  twins.json          for a handful of synthetic streams from the committed encoder: sha256 of the
                      reference's PPM for the reference-decodable twin (3 components, no DRI)
"""
import hashlib
import json
import shutil
import sys
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
ROOT = HERE.parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))

import helpers as H  # noqa: E402
from libkpeg_b200.synth import EMIT_RESTART, GRAY_CONTENT, QUIRK_FREE, SynthParams, synth_encode  # noqa: E402

REF = Path("/root/reference")

TWINS = [
    # name, params of the reference-decodable twin (3 components, no DRI emitted)
    ("rgb_64x48_q90", dict(width=64, height=48, quality=90, seed=1)),
    ("rgb_128x128_q95", dict(width=128, height=128, quality=95, seed=2)),
    ("rgb_160x120_q50", dict(width=160, height=120, quality=50, seed=3)),
    ("rgb_256x64_q90_ri16", dict(width=256, height=64, quality=90, seed=4, restart_interval=16)),
    ("gray_96x96_q90", dict(width=96, height=96, quality=90, seed=5, flags=QUIRK_FREE | GRAY_CONTENT)),
    ("rgb_64x64_q100", dict(width=64, height=64, quality=100, seed=6, noise_amp=40)),
    ("rgb_64x64_q90_noquirkfix", dict(width=64, height=64, quality=90, seed=7, flags=0)),
]


def main():
    assert H.have_reference_binary(), "build the reference first: make -C oracle"
    shutil.copyfile(REF / "misc/images/lena.jpg", HERE / "lena.jpg")
    lena = (HERE / "lena.jpg").read_bytes()
    ppm = H.reference_decode(lena)
    header, payload = H.split_ppm(ppm)
    np.savez_compressed(HERE / "lena_ref.npz", payload=payload, header=np.frombuffer(header, dtype=np.uint8),
                        ppm_sha256=np.array(hashlib.sha256(ppm).hexdigest()),
                        jpg_sha256=np.array(hashlib.sha256(lena).hexdigest()))
    print("lena.ppm sha256", hashlib.sha256(ppm).hexdigest(), len(ppm))

    twins = {}
    for name, kw in TWINS:
        p = SynthParams(**{"flags": QUIRK_FREE, **kw})
        jpg = synth_encode(p)
        ppm = H.reference_decode(jpg.tobytes())
        twins[name] = dict(params=kw, jpg_sha256=hashlib.sha256(jpg.tobytes()).hexdigest(), jpg_bytes=int(jpg.size),
                           ppm_sha256=hashlib.sha256(ppm).hexdigest(), ppm_bytes=len(ppm))
        print(name, twins[name]["ppm_sha256"][:16], jpg.size)
    (HERE / "twins.json").write_text(json.dumps(twins, indent=1, sort_keys=True) + "\n")


if __name__ == "__main__":
    main()
