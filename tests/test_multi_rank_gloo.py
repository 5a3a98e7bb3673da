"""world_size-2 (and 3) CPU runs of the multi-GPU host logic over torch.distributed/gloo: image
batches sharded by rank and restart-interval bands of one large image, each rank decoding its share
(with the CPU single-stepper standing in for the GPU), results gathered on rank 0 and compared with
the oracle's decode of the whole thing.  No collective touches the data path: the gather is the
host-side assembly the north-star describes."""
import os
import socket
import sys
from pathlib import Path

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, mode, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import helpers as H
    import libkpeg_b200 as K
    from libkpeg_b200.shard import shard_range, split_restart_bands
    from libkpeg_b200.synth import EMIT_RESTART, GRAY_CONTENT, QUIRK_FREE, SynthParams, synth_encode
    try:
        if mode == "batch":
            n, w, h = 7, 48, 40
            jpgs = [synth_encode(SynthParams(w, h, file_components=1, quality=90, flags=QUIRK_FREE | GRAY_CONTENT, seed=50 + i))
                    for i in range(n)]
            mine = shard_range(n, rank, world)
            out = np.zeros((n, h, w), dtype=np.uint8)
            if len(mine):
                parsed = [K.parse_jfif(jpgs[i]) for i in mine]
                scans = [jpgs[i][o:o + ln] for i, (_, o, ln) in zip(mine, parsed)]
                e = H.emu_decode(None, scans=scans, plan=parsed[0][0], sub_bits=128)
                assert e["status"] == 0
                out[mine.start:mine.stop] = e["pixels"]
            t = torch.from_numpy(out)
            dist.reduce(t, dst=0, op=dist.ReduceOp.SUM)  # disjoint slices: SUM == gather
            if rank == 0:
                for i in range(n):
                    assert np.array_equal(t.numpy()[i], H.oracle_decode(jpgs[i].tobytes())["pixels"]), i
        else:
            w, h = 96, 136  # 12 x 17 MCUs; one restart interval per MCU row
            jpg = synth_encode(SynthParams(w, h, quality=92, restart_interval=12, flags=QUIRK_FREE | EMIT_RESTART, seed=77))
            plan, off, ln = K.parse_jfif(jpg)
            # "bands": cut at balanced rows; "bands_by_bytes": cut at byte positions, rows from each band's marker count
            bands = split_restart_bands(plan, jpg[off:off + ln], world, by_bytes=(mode == "bands_by_bytes"))
            assert sum(b.rows for b in bands) == h and [b.row0 for b in bands] == sorted(b.row0 for b in bands)
            b = bands[rank]
            out = np.zeros((h, w, 3), dtype=np.uint8)
            if b.rows:
                e = H.emu_decode(None, scans=[b.scan], plan=b.plan, sub_bits=128)
                assert e["status"] == 0
                out[b.row0:b.row0 + b.rows] = e["pixels"][0]
            t = torch.from_numpy(out)
            dist.reduce(t, dst=0, op=dist.ReduceOp.SUM)
            if rank == 0:
                assert np.array_equal(t.numpy(), H.oracle_decode(jpg.tobytes())["pixels"])
        # the bench's timing reduction: max over ranks
        tm = torch.tensor([float(rank + 1)], dtype=torch.float64)
        dist.all_reduce(tm, op=dist.ReduceOp.MAX)
        assert float(tm) == float(world)
        ret[rank] = "ok"
    except Exception as ex:  # pragma: no cover
        ret[rank] = f"{type(ex).__name__}: {ex}"
        os._exit(1)  # do not wait in a collective the failed rank will never reach
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("mode,world", [("batch", 2), ("bands", 2), ("bands", 3), ("bands_by_bytes", 2), ("bands_by_bytes", 3)])
def test_sharded_decode_over_gloo(mode, world):
    port = _free_port()
    mgr = mp.Manager()
    ret = mgr.dict()
    try:
        mp.spawn(_worker, args=(world, port, mode, ret), nprocs=world, join=True)
    except Exception:
        pass
    assert dict(ret) == {r: "ok" for r in range(world)}


def test_shard_range_partition():
    from libkpeg_b200.shard import shard_range
    for n in (0, 1, 7, 4096):
        for world in (1, 2, 3, 8):
            parts = [shard_range(n, r, world) for r in range(world)]
            assert sum(len(p) for p in parts) == n
            assert [i for p in parts for i in p] == list(range(n))
            assert max(len(p) for p in parts) - min(len(p) for p in parts) <= 1
