"""Target program for ncu captures: uploads one synthetic shard and runs the resident kernel sequence N times.
    python tests/gpu_prof_target.py [shape] [MB] [runs]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from phyngsc_b200 import api, synth  # noqa: E402

shape = sys.argv[1] if len(sys.argv) > 1 else "36bp"
mb = int(sys.argv[2]) if len(sys.argv) > 2 else 256
runs = int(sys.argv[3]) if len(sys.argv) > 3 else 2
data = synth.fastq(shape, 2, target_bytes=mb * 1_000_000)
ctx = api.Context(0, max_batch_bytes=data.size + (1 << 20), max_subblocks=max(192, data.size // (6 << 20)))
prm = api.region_params(data.size, 1, 0)
ctx.upload(data)
for i in range(runs):
    d, res = ctx.compress_resident(data.size, prm)
    print(f"run {i}: {res.kernel_ms:.3f} ms {res.kernel_launches} launches")
ctx.close()
