"""Kernel-only timing of one shape under several settings of the experiment switches (each in its own process):
   python tests/gpu_sweep.py <shape> <MB> VAR=a,b VAR2=c,d ..."""
import itertools
import os
import subprocess
import sys

shape, mb = sys.argv[1], sys.argv[2]
axes = [(a.split("=")[0], a.split("=")[1].split(",")) for a in sys.argv[3:]]
code = ("import sys; sys.path.insert(0, '.'); from phyngsc_b200 import api, synth; import numpy as np\n"
        f"data = synth.fastq('{shape}', 2, target_bytes={mb} * 1000000)\n"
        "ctx = api.Context(0, max_batch_bytes=data.size + (1 << 20), max_subblocks=max(192, data.size // (6 << 20)))\n"
        "prm = api.region_params(data.size, 1, 0); ctx.upload(data)\n"
        "ms = [ctx.compress_resident(data.size, prm)[1].kernel_ms for _ in range(6)]\n"
        "ctx.profile(True); [ctx.compress_resident(data.size, prm) for _ in range(3)]; pr = ctx.profile_read()\n"
        "print(f'{min(ms[2:]):.3f} ms  {data.size / min(ms[2:]) / 1e6:.1f} GB/s  ' + ' '.join(f'{k}={v:.3f}' for k, v in pr.items() if v > 0.05))\n")
for combo in itertools.product(*[v for _, v in axes]):
    env = dict(os.environ)
    tag = []
    for (k, _), v in zip(axes, combo):
        env[k] = v
        tag.append(f"{k}={v}")
    p = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True)
    print(" ".join(tag), "->", (p.stdout.strip().splitlines() or [p.stderr.strip()[-300:]])[-1], flush=True)
