"""Copy-only ceiling of the end-to-end leg: every rank moves what bench.py's e2e leg moves per step -- its region host -> device and
the payloads device -> host, from / to pinned memory in 64 MiB pieces on two streams -- with no kernels in between, all ranks at once.
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P profiles/copy_ceiling.py [GB per rank] [out fraction]
Rank 0 prints one JSON line: per-direction and concurrent GB/s per rank (min / mean) and the aggregate."""
import json
import os
import sys
import time

import torch
import torch.distributed as dist

rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
gb = float(sys.argv[1]) if len(sys.argv) > 1 else 2.0
frac = float(sys.argv[2]) if len(sys.argv) > 2 else 0.36
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
n_in, n_out, piece = int(gb * 1e9), int(gb * 1e9 * frac), 64 << 20
h_in = torch.empty(n_in, dtype=torch.uint8).pin_memory(); h_in.fill_(65)
h_out = torch.empty(n_out, dtype=torch.uint8).pin_memory()
d_in = torch.empty(n_in, dtype=torch.uint8, device="cuda")
d_out = torch.zeros(n_out, dtype=torch.uint8, device="cuda")
s_in, s_out = torch.cuda.Stream(), torch.cuda.Stream()


def barrier():
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()


def run(do_in, do_out, reps=3):
    best = 1e9
    for _ in range(reps + 1):
        barrier()
        t = time.perf_counter()
        if do_in:
            with torch.cuda.stream(s_in):
                for o in range(0, n_in, piece):
                    d_in[o:o + piece].copy_(h_in[o:o + piece], non_blocking=True)
        if do_out:
            with torch.cuda.stream(s_out):
                for o in range(0, n_out, piece):
                    h_out[o:o + piece].copy_(d_out[o:o + piece], non_blocking=True)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t
        barrier()
        best = min(best, dt)
    return best


def gather(x):
    if world == 1:
        return [x]
    t = torch.tensor([x], dtype=torch.float64, device="cuda")
    out = [torch.zeros_like(t) for _ in range(world)]
    dist.all_gather(out, t)
    return [float(v.item()) for v in out]


res = {}
for name, di, do, nbytes in (("h2d", True, False, n_in), ("d2h", False, True, n_out), ("both", True, True, n_in)):
    secs = gather(run(di, do))
    res[name] = {"per_rank_gbs_min": nbytes / max(secs) / 1e9, "per_rank_gbs_mean": sum(nbytes / s / 1e9 for s in secs) / world,
                 "aggregate_gbs": world * nbytes / max(secs) / 1e9, "seconds_max": max(secs)}
if rank == 0:
    print(json.dumps({"n_gpus": world, "gb_in_per_rank": gb, "out_fraction": frac,
                      "note": "'both': input bytes per second while the payload copy runs the other way at the same time (the e2e leg's ceiling)", **res}))
if world > 1:
    dist.destroy_process_group()
