"""File-to-file timing probe of the drop-in driver (not the bench): python tests/gpu_driver_probe.py [shape] [MB] [ranks] [threads]
Writes the synthetic file to tmpfs, runs host/phyNGSC_b200 on it and prints the driver's own per-rank table."""
import os
import subprocess
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from phyngsc_b200 import build, container, synth  # noqa: E402

shape = sys.argv[1] if len(sys.argv) > 1 else "100bp"
mb = int(sys.argv[2]) if len(sys.argv) > 2 else 2000
ranks = int(sys.argv[3]) if len(sys.argv) > 3 else 1
threads = sys.argv[4] if len(sys.argv) > 4 else "1"
exe = build.build_driver()
seg = synth.fastq(shape, 3, target_bytes=min(mb, 1000) * 1_000_000)
src, dst = f"/dev/shm/phy_probe_{os.getpid()}.fastq", f"/dev/shm/phy_probe_{os.getpid()}.ngsc"
try:
    with open(src, "wb") as f:
        for _ in range(max(1, mb // 1000)):
            f.write(memoryview(seg))
    size = os.path.getsize(src)
    for it in range(2):
        env = dict(os.environ, PHY_SHIM_NP=str(ranks))
        t = time.perf_counter()
        p = subprocess.run([exe, src, dst, threads], env=env, capture_output=True, text=True, timeout=1200)
        wall = time.perf_counter() - t
        print(f"--- run {it}: rc {p.returncode}, wall {wall:.3f}s = {size / wall / 1e9:.2f} GB/s of {size} bytes")
        print(p.stdout[-3000:]); print(p.stderr[-1500:])
    ng = container.read_ngsc(dst)
    print("container:", ng["footer"]["n_blocks"], "blocks", ng["footer"]["n_subblocks"], "subblocks, fastq_size ok:", ng["footer"]["fastq_size"] == size)
finally:
    for q in (src, dst):
        if os.path.exists(q):
            os.remove(q)
