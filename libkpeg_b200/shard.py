"""Work partitioning for the multi-GPU paths (one process per GPU, no collective in the data path):

  * batches of independent images: contiguous index ranges per rank (BASELINE.json configs[3]);
  * restart-interval tiles of one very large image: bands of whole MCU rows per rank, each band being
    a complete smaller image of the same width and tables (BASELINE.json configs[4]).

The band cutter itself is native (kpeg_split_restart_bands in libkpeg_cuda.so, host only); this
module is the thin Python face bench.py and the tests use."""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass

import numpy as np

from .api import KPEG_OK, KpegError, Plan, load_cuda_library


def shard_range(n: int, rank: int, world: int) -> range:
    """Contiguous, balanced split of n items; ranks beyond n get an empty range."""
    lo = (n * rank) // world
    hi = (n * (rank + 1)) // world
    return range(lo, hi)


@dataclass
class Band:
    plan: Plan          # same tables / width / restart interval, height = rows of this band
    scan: np.ndarray    # view of the parent's entropy-coded bytes
    row0: int           # first pixel row of the band in the parent image
    rows: int           # pixel rows in the band


def split_restart_bands(plan: Plan, scan: np.ndarray, parts: int, by_bytes: bool = False) -> list[Band]:
    """by_bytes: cut at byte positions (kpeg_split_restart_bands_by_bytes: the bands' heights follow from their restart
    marker counts) instead of at balanced rows (kpeg_split_restart_bands: one walk over all markers of the scan)."""
    lib = load_cuda_library()
    scan = np.ascontiguousarray(scan, dtype=np.uint8)
    ob = (C.c_uint64 * parts)()
    oe = (C.c_uint64 * parts)()
    orow = (C.c_uint32 * (parts + 1))()
    fn = lib.kpeg_split_restart_bands_by_bytes if by_bytes else lib.kpeg_split_restart_bands
    rc = fn(scan.ctypes.data, scan.size, C.byref(plan), parts, ob, oe, orow)
    if rc != KPEG_OK:
        raise KpegError(rc, "kpeg_split_restart_bands_by_bytes" if by_bytes else "kpeg_split_restart_bands")
    bands = []
    for b in range(parts):
        r0, r1 = min(orow[b] * 8, plan.height), min(orow[b + 1] * 8, plan.height)
        p = Plan.from_buffer_copy(bytes(plan))
        p.height = max(r1 - r0, 0)
        bands.append(Band(plan=p, scan=scan[ob[b]:oe[b]], row0=r0, rows=max(r1 - r0, 0)))
    return bands
