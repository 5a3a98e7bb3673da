#!/usr/bin/env python3
"""Device-resident decode rate of the BASELINE.json parity configurations (informational; bench.py measures
config 3, the one the metric is quoted on).  Run on a GPU box:  python tools/bench_configs.py"""
import sys, time
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import numpy as np
import libkpeg_b200 as K
from libkpeg_b200.api import pack_batch, packed_offsets
from libkpeg_b200.synth import EMIT_RESTART, GRAY_CONTENT, QUIRK_FREE, SynthParams, synth_encode


def run(name, params, nb, steps=30):
    jpgs = [synth_encode(SynthParams(**{**params, "seed": params.get("seed", 1) + i})) for i in range(min(nb, 16))]
    jpgs = [jpgs[i % len(jpgs)] for i in range(nb)]
    parsed = [K.parse_jfif(j) for j in jpgs]
    plan = parsed[0][0]
    plan.flags = K.KPEG_FLAG_REF_PARITY
    scans = [j[o:o + n] for j, (_, o, n) in zip(jpgs, parsed)]
    packed, off = pack_batch(scans), packed_offsets(scans)
    dec = K.Decoder(0)
    npix = plan.width * plan.height
    d_in = dec.device_alloc(packed.size + 64)
    d_out = [dec.device_alloc(nb * npix * plan.ncomp + 64) for _ in range(2)]
    dec.h2d(d_in, packed)
    for i in range(10):
        dec.submit_batch_packed_device(plan, nb, d_in, off, d_out[i & 1])
    dec.wait()
    t0 = time.perf_counter()
    for i in range(steps):
        dec.submit_batch_packed_device(plan, nb, d_in, off, d_out[i & 1])
    dec.wait()
    dt = time.perf_counter() - t0
    st = dec.last_stats
    print(f"{name}: {nb} x {plan.width}x{plan.height}x{plan.ncomp}, scan {packed.size / nb / 1e6:.2f} MB/image, "
          f"{steps * nb * npix / dt / 1e9:.1f} Gpixel/s, {packed.size * steps / dt / 1e9:.1f} GB/s of scan bytes, "
          f"relay rounds {st.sync_rounds}")
    dec.close()


if __name__ == "__main__":
    run("config 1 (1080p q90, DRI 16)", dict(width=1920, height=1080, quality=90, restart_interval=16,
                                              flags=QUIRK_FREE | EMIT_RESTART), nb=32)
    run("config 2 (4K q95, no DRI)", dict(width=3840, height=2160, quality=95, flags=QUIRK_FREE), nb=8)
    run("config 3 (512x512 gray q90)", dict(width=512, height=512, file_components=1, quality=90,
                                            flags=QUIRK_FREE | GRAY_CONTENT), nb=1024)
    run("config 4 (16384x16384 q90, DRI = MCU row)", dict(width=16384, height=16384, quality=90, restart_interval=2048,
                                                        flags=QUIRK_FREE | EMIT_RESTART), nb=1, steps=8)
