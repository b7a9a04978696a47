"""Kernel-only time of one resident region under several batch sizes: python tests/gpu_batch_probe.py [shape] [MB] [batch MiB ...]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from phyngsc_b200 import api, synth  # noqa: E402

shape = sys.argv[1] if len(sys.argv) > 1 else "100bp"
mb = int(sys.argv[2]) if len(sys.argv) > 2 else 1000
sizes = [int(x) for x in sys.argv[3:]] or [1040, 512, 256]
data = synth.fastq(shape, 2, target_bytes=mb * 1_000_000)
prm = api.region_params(data.size, 1, 0)
for bm in sizes:
    ctx = api.Context(0, max_batch_bytes=bm << 20, max_subblocks=max(64, (bm << 20) // (6 << 20) + 16))
    ctx.upload(data)
    ms = [ctx.compress_resident(data.size, prm, max_descs=data.size // (4 << 20) + 64)[1] for _ in range(6)]
    best = min(r.kernel_ms for r in ms[2:])
    print(f"{shape} {mb} MB, batches of {bm} MiB: {best:.3f} ms  {data.size / best / 1e6:.1f} GB/s  ({ms[-1].n_batches} batches, {ms[-1].kernel_launches} launches)", flush=True)
    ctx.close()
