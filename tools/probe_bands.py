"""GPU box: time kpeg_cuda_decode on the 16384x16384 restart-marked image (BASELINE configs[4]) from host memory into
(a) a pinned buffer, (b) a registered /dev/shm frame, with the band pipeline on (KPEG_BANDS=4) and off (=1)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import libkpeg_b200 as K
from libkpeg_b200 import api
from libkpeg_b200.synth import EMIT_RESTART, QUIRK_FREE, SynthParams, synth_encode
side = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
jpg = synth_encode(SynthParams(width=side, height=side, quality=90, restart_interval=side // 8, flags=QUIRK_FREE | EMIT_RESTART, seed=5))
plan, off, n = K.parse_jfif(jpg)
plan.flags = K.KPEG_FLAG_REF_PARITY
h_in = api.PinnedArray(n)
h_in.array[:] = jpg[off:off + n]
lib = api.load_cuda_library()
nb = side * side * 3
pin = api.PinnedArray(nb)
path = "/dev/shm/kpeg_probe_frame"
with open(path, "wb") as f:
    f.truncate(nb)
shm = np.memmap(path, dtype=np.uint8, mode="r+", shape=(nb,))
print("register rc", lib.kpeg_cuda_host_register(shm.ctypes.data, nb))
for bands, chain in (("1", "1"), ("2", "1"), ("4", "1"), ("8", "1")):
    os.environ["KPEG_BANDS"] = bands
    os.environ["KPEG_D2H_CHAIN"] = chain
    dec = K.Decoder(device=0)
    for name, out in (("pinned", pin.array), ("shm-registered", shm)):
        dec.decode_scan(plan, h_in.array, out=out)
        t0 = time.perf_counter()
        for _ in range(5):
            dec.decode_scan(plan, h_in.array, out=out)
        dt = (time.perf_counter() - t0) / 5
        print(f"KPEG_BANDS={bands} chain={chain} {name}: {dt*1e3:.1f} ms per decode = {side*side/dt/1e9:.2f} Gpixel/s ({nb/dt/1e9:.1f} GB/s out), launches {dec.last_stats.kernel_launches}")
    dec.set_profiling(True)
    dec.decode_scan(plan, h_in.array, out=pin.array)
    print("   stage ms:", {k: round(v, 2) for k, v in dec.last_stats.stage_ms().items() if v})
    dec.close()
os.unlink(path)
