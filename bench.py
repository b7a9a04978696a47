#!/usr/bin/env python
"""bench.py -- FASTQ compress throughput (input GB/s) of the B200-native phyNGSC subblock path.

    python bench.py --gpus N --steps K --warmup W            # our arm (CUDA path through the C ABI)
    python bench.py --impl reference --gpus N --steps K ...  # the reference's own CPU implementation

A "step" is one pass of the hot path over one rank's shard of synthetic FASTQ (BASELINE.json configs[1]:
1 GB, 36 bp Illumina-style reads, ERR005195-like titles).  One process per GPU; ranks own independent
shards (weak scaling, no data-path collective); the only cross-rank datum is the exclusive scan of the
compressed sizes that fixes the file offsets (done with torch.distributed here, MPI_Exscan in the host driver).

value   kernel-only: shard resident in HBM, CUDA-event time of the whole kernel sequence, max over ranks
e2e     the same shard from pinned HOST memory through phy_compress_region: H2D + kernels + D2H of the
        payloads into pinned host memory, wall clock around the synchronous call, max over ranks
"""
import argparse
import json
import os
import re
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

WORKLOADS = {  # name -> (synth shape, description)
    "36bp": "synthetic 36 bp Illumina-style reads, ERR005195-like titles (BASELINE.json configs[1])",
    "100bp": "synthetic 100 bp reads, N runs, 41-symbol quality (configs[2] shape)",
    "150bp_paired": "synthetic 150 bp paired-style reads (configs[3] shape)",
    "var50_205": "synthetic 50-205 bp reads, 17-field titles (configs[4] shape, in-domain cap)",
}


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--shape", default="36bp", choices=sorted(WORKLOADS))
    ap.add_argument("--mb", type=int, default=1000, help="shard size per GPU in MB (10^6 bytes)")
    ap.add_argument("--cpu-sample-mb", type=int, default=256)
    ap.add_argument("--e2e-batch-mb", type=int, default=64, help="batch size (MiB) of the pipelined end-to-end run")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    return ap.parse_args()


class ClockSampler:
    """nvidia-smi clocks / throttle reasons while the timed region runs (the recipe's `-lms` loop, started
    before the region and stopped after it)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.rows, self.proc = [], None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(index), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except Exception:  # noqa: BLE001
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            parts = [x.strip() for x in line.split(",")]
            if len(parts) >= 7:
                self.rows.append((time.perf_counter(), parts))

    def wait_first(self, timeout=5.0):
        t = time.perf_counter()
        while self.proc is not None and not self.rows and time.perf_counter() - t < timeout:
            time.sleep(0.02)

    def summary(self, t0=None, t1=None):
        if self.proc is not None:
            self.proc.terminate()
        rows = [r for (t, r) in self.rows if (t0 is None or t >= t0 - 0.11) and (t1 is None or t <= t1 + 0.11)] or [r for _, r in self.rows]
        if not rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}

        def num(x):
            try:
                return float(x)
            except ValueError:
                return None
        sm = sorted(v for v in (num(r[0]) for r in rows) if v is not None)
        reasons = set()
        for r in rows:
            for name, v in zip(["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"], r[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": num(rows[0][1]), "reasons": sorted(reasons), "samples": len(rows),
                "power_w_max": max((v for v in (num(r[2]) for r in rows) if v is not None), default=None)}


def make_shard(shape, mb, seed, pinned):
    from phyngsc_b200 import synth
    n = mb * 1_000_000
    data = synth.fastq(shape, seed, target_bytes=n, out=pinned.array if pinned is not None else None)
    return data


def shard_stats(data):
    """records and title / sequence byte totals of a shard (for the per-kernel algorithmic bytes)."""
    nl = np.flatnonzero(data == 10)
    nrec = nl.size // 4
    nl = nl[: nrec * 4].reshape(nrec, 4)
    starts = np.concatenate(([0], nl[:-1, 3] + 1))
    title = int((nl[:, 0] - starts + 1).sum())
    seq = int((nl[:, 1] - nl[:, 0] - 1).sum())
    return nrec, title, seq


def host_cores():
    try:
        return len(os.sched_getaffinity(0))
    except Exception:  # noqa: BLE001
        return os.cpu_count() or 1


def run_reference_sample(data, sample_bytes, tmpdir="/dev/shm"):
    """Times the reference's own CPU implementation (oracle/_ref/phyNGSC_ref = the unmodified sources built
    against the fork-based MPI stand-in) on the first `sample_bytes` of the shard, cut at a record boundary,
    with one single-threaded rank per host core (rank scaling is the reference's effective axis).
    Falls back to the oracle port (one core) when the reference binary is not there.
    -> dict(value GB/s, cores, kind, sample, seconds)"""
    from oracle import phy_oracle as O
    n = min(sample_bytes, data.size)
    nl = np.flatnonzero(data[:n] == 10)
    cut = int(nl[(nl.size // 4) * 4 - 1]) + 1  # whole records only
    sample = data[:cut]
    cores = host_cores()
    if O.have_reference():
        npr = max(2, min(cores, 64))
        src = os.path.join(tmpdir, f"phy_bench_{os.getpid()}.fastq")
        dst = src + ".ngsc"
        sample.tofile(src)
        try:
            t = time.perf_counter()
            out = O.run_reference(src, dst, np_ranks=npr, threads=1, timeout=1200)
            wall = time.perf_counter() - t
        finally:
            for p in (src, dst):
                if os.path.exists(p):
                    os.remove(p)
        times = [float(m.group(1)) for m in re.finditer(r"^\s*\d+\s+([0-9.]+)\s+\d+\s+\d+\s*$", out, re.M)]
        secs = max(times) if times else wall
        return dict(value=cut / secs / 1e9, unit="GB/s", cores=npr, kind="reference", seconds=secs,
                    sample=f"first {cut} bytes of the shard, unmodified reference, np={npr} x threads=1 over the fork-based MPI stand-in, "
                           f"tmpfs I/O included (max COMP_TIME {secs:.3f}s, wall {wall:.3f}s)")
    O.build()
    small = sample[: min(cut, 64_000_000)]
    t = time.perf_counter()
    O.compress_rank(small, 1, 0)
    secs = time.perf_counter() - t
    return dict(value=small.size / secs / 1e9, unit="GB/s", cores=1, kind="port", seconds=secs,
                sample=f"first {small.size} bytes of the shard, oracle/phy_oracle.c (scalar port), 1 core")


def main():
    # stdout carries exactly one JSON line: everything libraries print (NCCL's version banner, warnings) goes to stderr
    real_stdout = os.dup(1)
    os.dup2(2, 1)

    def emit(obj):
        os.write(real_stdout, (json.dumps(obj) + "\n").encode())

    a = parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    config = {"workload": f"{a.mb} MB per GPU, {WORKLOADS[a.shape]}", "shape": a.shape, "shard_mb": a.mb, "window_bytes": 1 << 23,
              "partitioning": "one shard (np=1 region) per GPU", "l2": "inputs (>= 1 GB per step) larger than the 126 MB L2; no flush needed"}

    if a.impl == "reference":
        if rank != 0:
            return 0
        data = make_shard(a.shape, a.mb, 2, None)
        sample = min(a.cpu_sample_mb * 1_000_000, data.size)
        vals = []
        for i in range(a.warmup + a.steps):
            r = run_reference_sample(data, sample)
            if i >= a.warmup:
                vals.append(r)
        secs = sum(v["seconds"] for v in vals) / len(vals)
        value = float(np.mean([v["value"] for v in vals]))
        line = {"impl": "reference", "metric": "fastq_compress_input_throughput", "value": value, "unit": "GB/s", "n_gpus": a.gpus,
                "steps": a.steps, "warmup": a.warmup, "ms_per_step": secs * 1e3, "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": "u8", "data": "synthetic", "config": config,
                "cpu_baseline": {"value": value, "unit": "GB/s", "cores": vals[-1]["cores"], "kind": vals[-1]["kind"], "sample": vals[-1]["sample"]},
                "e2e": {"value": value, "unit": "GB/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0}
        emit(line)
        return 0

    import torch
    import torch.distributed as dist
    from phyngsc_b200 import api
    from phyngsc_b200 import dist as pdist
    if not torch.cuda.is_available():
        print("bench.py: no CUDA device -- the product path has no CPU fallback", file=sys.stderr)
        return 2
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()

    def max_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def sum_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return float(t.item())

    nbytes = a.mb * 1_000_000
    pin_in = api.pinned_array(nbytes + 8192)
    data = make_shard(a.shape, a.mb, 2 + rank, pin_in)
    pin_out = api.pinned_array(data.size // 2 + (1 << 20))
    ctx = api.Context(local, max_batch_bytes=data.size + (1 << 20), max_subblocks=max(64, data.size // (6 << 20) + 16))
    prm = api.region_params(data.size, 1, 0)

    # ---- kernel-only ---------------------------------------------------------------------------------------
    ctx.upload(data)
    for _ in range(a.warmup):
        descs, res = ctx.compress_resident(data.size, prm)
    sampler = ClockSampler(local)
    sampler.wait_first()
    barrier()
    k_ms, launches = 0.0, 0
    t0 = time.perf_counter()
    t_region0 = t0
    for _ in range(a.steps):
        descs, res = ctx.compress_resident(data.size, prm)
        k_ms += res.kernel_ms
        launches += res.kernel_launches
    barrier()
    wall_k = time.perf_counter() - t0
    ms_per_step = max_over_ranks(k_ms / a.steps)
    bytes_in, bytes_out = res.bytes_in, res.bytes_out
    total_in = sum_over_ranks(float(bytes_in))
    value = total_in / (ms_per_step * 1e-3) / 1e9

    # ---- end to end from pinned host memory ---------------------------------------------------------------------
    # a second context with 256 MiB batches: upload of batch b+1, kernels of batch b and download of batch b-1 overlap
    ctx_e = api.Context(local, max_batch_bytes=a.e2e_batch_mb << 20, max_subblocks=(a.e2e_batch_mb << 20) // (4 << 20) + 16)
    ctx_e.compress_region(data, prm, out=pin_out.array)
    barrier()
    t0 = time.perf_counter()
    for _ in range(a.steps):
        d2, o2, r2 = ctx_e.compress_region(data, prm, out=pin_out.array)
        _offset, _total = pdist.exscan_bytes(r2.bytes_out, device="cuda")  # file offsets (MPI_Exscan in the host driver)
    barrier()
    e2e_s = max_over_ranks((time.perf_counter() - t0) / a.steps)
    clocks = sampler.summary(t_region0, time.perf_counter())
    e2e = {"value": total_in / e2e_s / 1e9, "unit": "GB/s", "h2d_bytes_per_step": int(data.size), "d2h_bytes_per_step": int(r2.out_used),
           "ms_per_step": e2e_s * 1e3, "batches": int(r2.n_batches), "batch_mb": a.e2e_batch_mb, "h2d_ms_sum": r2.h2d_ms, "kernel_ms_sum": r2.kernel_ms,
           "d2h_ms_sum": r2.d2h_ms, "overlap": "upload / kernels / download of consecutive batches run on three streams"}
    assert r2.bytes_out == bytes_out and r2.bytes_in == bytes_in, "pipelined and resident runs disagree"
    ctx_e.close()

    # ---- per-kernel times -> roofline of the dominant kernel ---------------------------------------------------------
    ctx.profile(True)
    for _ in range(max(2, a.steps)):
        ctx.compress_resident(data.size, prm)
    stages = ctx.profile_read()
    ctx.profile(False)
    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(peaks_path):
        peak, peak_src = float(json.load(open(peaks_path))["hbm_gbs"]), "MEASURED_PEAKS.json hbm_gbs (measured copy, read+write)"
    else:
        peak, peak_src = 6650.0, "fallback 6.65 TB/s (B200_PROFILING.md)"
    line = None
    if rank == 0:
        nrec, title_b, seq_b = shard_stats(data[: int(bytes_in)])
        # algorithmic bytes per launch of each stage: what it must read + must write
        alg = {"nl_count": bytes_in + bytes_in // 8, "nl_emit": bytes_in // 8 + 12 * nrec, "stat1": bytes_in + 2 * nrec, "qhist": seq_b + 10 * nrec,
               "stat2": title_b + 8 * nrec, "lengths": bytes_in + 8 * nrec, "layout": 16 * nrec, "emit": bytes_in + bytes_out}
        dom = max((k for k in stages if k in alg), key=lambda k: stages[k])
        ach = alg[dom] / (stages[dom] * 1e-3) / 1e9
        traffic = None
        tpath = os.path.join(ROOT, "profiles", "ncu_traffic.json")
        if os.path.exists(tpath):  # one ncu --set full capture, scaled to this shard by algorithmic bytes
            tj = json.load(open(tpath)).get("k_" + dom)
            if tj:
                traffic = int(tj["traffic"] * alg[dom] / tj["algorithmic"])
        roofline = {"bound": "hbm", "kernel": "k_" + dom, "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak, "traffic": traffic,
                    "traffic_source": "profiles/ncu_traffic.json (ncu --set full capture, scaled by algorithmic bytes)" if traffic else None,
                    "peak_source": peak_src, "algorithmic_bytes": int(alg[dom]), "kernel_ms": stages[dom],
                    "pipeline": {"algorithmic_bytes": int(bytes_in + bytes_out), "achieved": (bytes_in + bytes_out) / (ms_per_step * 1e-3) / 1e9,
                                 "frac": (bytes_in + bytes_out) / (ms_per_step * 1e-3) / 1e9 / peak},
                    "stage_ms": {k: round(v, 4) for k, v in stages.items()}}
        cpu = None
        if world == 1 and not a.no_cpu_baseline:
            cpu = run_reference_sample(np.asarray(data), a.cpu_sample_mb * 1_000_000)
            cpu.pop("seconds", None)
        line = {"metric": "fastq_compress_input_throughput", "value": value, "unit": "GB/s", "n_gpus": world, "steps": a.steps, "warmup": a.warmup,
                "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
                "config": config, "e2e": e2e, "gpu_launches": int(launches), "clocks": clocks, "roofline": roofline, "cpu_baseline": cpu,
                "compression": {"bytes_in": int(bytes_in), "bytes_out": int(bytes_out), "ratio": bytes_in / max(1, bytes_out)},
                "wall_ms_per_step_kernel_loop": wall_k / a.steps * 1e3}
    ctx.close()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    if line is not None:
        emit(line)
    return 0


if __name__ == "__main__":
    sys.exit(main())
