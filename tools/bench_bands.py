#!/usr/bin/env python3
"""BASELINE.json configs[4]: one 16384x16384 RGB 4:4:4 image with a restart interval per MCU row, cut at its RSTn
markers into bands of whole MCU rows (kpeg_split_restart_bands), one band per GPU, no collective in the data path.
Run on a GPU box, one process per GPU:
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 tools/bench_bands.py
(or plain `python tools/bench_bands.py` for N = 1).  Rank 0 prints one JSON line: device-resident Gpixel/s of the
whole image (band scans already in HBM, pixels left in HBM), the max over ranks of the CUDA-side time, and the same end
to end with pinned host buffers (each rank copies its band's pixels into its rows of a host frame)."""
import json, os, sys, time
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import numpy as np
import torch
import libkpeg_b200 as K
from libkpeg_b200.api import pack_batch, packed_offsets
from libkpeg_b200.shard import split_restart_bands
from libkpeg_b200.synth import EMIT_RESTART, QUIRK_FREE, SynthParams, synth_encode


def main():
    rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    side = int(os.environ.get("KPEG_BANDS_SIDE", 16384))
    steps = int(os.environ.get("KPEG_BANDS_STEPS", 10))
    torch.cuda.set_device(local)
    if world > 1:
        torch.distributed.init_process_group("nccl")  # barrier and max-over-ranks only
    jpg = synth_encode(SynthParams(width=side, height=side, quality=90, restart_interval=side // 8,
                                   flags=QUIRK_FREE | EMIT_RESTART, seed=5))
    plan, off, n = K.parse_jfif(jpg)
    plan.flags = K.KPEG_FLAG_REF_PARITY
    band = split_restart_bands(plan, jpg[off:off + n], world)[rank]
    dec = K.Decoder(local)
    nbytes = plan.width * band.rows * plan.ncomp
    packed, offs = pack_batch([band.scan]), packed_offsets([band.scan])  # the band as a batch of one image
    d_in = dec.device_alloc(packed.size + 64)
    d_out = [dec.device_alloc(nbytes + 64) for _ in range(2)]
    dec.h2d(d_in, packed)

    def run(k):
        for i in range(k):
            dec.submit_batch_packed_device(band.plan, 1, d_in, offs, d_out[i & 1])
        dec.wait()

    def sync():
        torch.cuda.synchronize()
        if world > 1:
            torch.distributed.barrier()

    run(10)  # every lane of the context has its scratch after this
    sync()
    t0 = time.perf_counter()
    run(steps)
    torch.cuda.synchronize()
    dt = torch.tensor([time.perf_counter() - t0], device="cuda")
    if world > 1:
        torch.distributed.all_reduce(dt, op=torch.distributed.ReduceOp.MAX)
    # end to end: pinned host scan in, the band's rows of a pinned host frame out
    h_in = torch.from_numpy(np.ascontiguousarray(band.scan)).pin_memory()
    h_out = torch.empty(nbytes, dtype=torch.uint8).pin_memory()
    scan_np, out_np = h_in.numpy(), h_out.numpy()
    dec.decode_scan(band.plan, scan_np, out=out_np)
    sync()
    t0 = time.perf_counter()
    for _ in range(steps):
        dec.decode_scan(band.plan, scan_np, out=out_np)
    de = torch.tensor([time.perf_counter() - t0], device="cuda")
    if world > 1:
        torch.distributed.all_reduce(de, op=torch.distributed.ReduceOp.MAX)
    if rank == 0:
        npix = side * side
        print(json.dumps({"workload": f"{side}x{side} RGB 4:4:4 q90, restart interval = one MCU row, {world} band(s)",
                          "n_gpus": world, "steps": steps, "band_rows": band.rows, "band_scan_mb": round(band.scan.size / 1e6, 2),
                          "device_resident_gpixel_per_s": round(npix * steps / dt.item() / 1e9, 2),
                          "e2e_gpixel_per_s": round(npix * steps / de.item() / 1e9, 2),
                          "timer": "host clock around submit..wait after a device sync and barrier, max over ranks"}))
    dec.close()
    if world > 1:
        torch.distributed.destroy_process_group()


if __name__ == "__main__":
    main()
