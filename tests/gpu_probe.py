"""Quick per-stage timing probe (not the bench): python tests/gpu_probe.py [shape] [MB]"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
from phyngsc_b200 import api, synth  # noqa: E402

shape = sys.argv[1] if len(sys.argv) > 1 else "36bp"
mb = int(sys.argv[2]) if len(sys.argv) > 2 else 256
t = time.time()
data = synth.fastq(shape, 2, target_bytes=mb * 1_000_000)
print(f"generated {data.size} bytes in {time.time() - t:.1f}s")
ctx = api.Context(0, max_batch_bytes=data.size + (1 << 20), max_subblocks=max(192, data.size // (6 << 20)))
prm = api.region_params(data.size, 1, 0)
ctx.upload(data)
for i in range(3):
    d, res = ctx.compress_resident(data.size, prm)
    print(f"resident run {i}: {res.kernel_ms:.3f} ms  {res.bytes_in / res.kernel_ms / 1e6:.1f} GB/s in, out/in {res.bytes_out / res.bytes_in:.3f}, "
          f"{res.n_subblocks} subblocks, {res.kernel_launches} launches")
ctx.profile(True)
for i in range(3):
    ctx.compress_resident(data.size, prm)
pr = ctx.profile_read()
tot = sum(pr.values())
for k, v in pr.items():
    print(f"   {k:14s} {v:8.3f} ms  {100 * v / tot:5.1f}%")
print(f"   {'total':14s} {tot:8.3f} ms")
ctx.profile(False)
pin = api.pinned_array(data.size); pin.array[:] = data
pout = api.pinned_array(data.size // 2 + (1 << 20))
for i in range(2):
    t = time.perf_counter()
    d, o, res = ctx.compress_region(pin.array, prm, out=pout.array)
    dt = time.perf_counter() - t
    print(f"e2e pinned run {i}: {dt * 1e3:.2f} ms wall  {data.size / dt / 1e9:.2f} GB/s  (h2d {res.h2d_ms:.2f} k {res.kernel_ms:.2f} d2h {res.d2h_ms:.2f})")

if os.environ.get("PROBE_E2E_SWEEP"):
    for bmb in (32, 64, 128, 256):
        cx = api.Context(0, max_batch_bytes=bmb << 20, max_subblocks=(bmb << 20) // (4 << 20) + 16)
        best = 1e9
        for i in range(4):
            t = time.perf_counter()
            d, o, res = cx.compress_region(pin.array, prm, out=pout.array)
            best = min(best, time.perf_counter() - t)
        print(f"e2e batch {bmb:4d} MiB: best {best * 1e3:.2f} ms  {data.size / best / 1e9:.2f} GB/s  batches {res.n_batches} (h2d {res.h2d_ms:.2f} k {res.kernel_ms:.2f} d2h {res.d2h_ms:.2f})")
        cx.close()
