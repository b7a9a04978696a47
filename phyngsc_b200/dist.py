"""Multi-rank plumbing shared by bench.py and the tests: the path shards by working region with no data-path
collective; the only cross-rank datum is the exclusive scan of the ranks' compressed sizes that fixes the
file offsets (MPI_Exscan in the C++ driver, torch.distributed here)."""
import torch
import torch.distributed as dist


def exscan_bytes(nbytes, device="cpu"):
    """-> (offset of this rank's output in the file, total bytes over all ranks)."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return 0, int(nbytes)
    world, rank = dist.get_world_size(), dist.get_rank()
    sizes = [torch.zeros(1, dtype=torch.int64, device=device) for _ in range(world)]
    dist.all_gather(sizes, torch.tensor([int(nbytes)], dtype=torch.int64, device=device))
    vals = [int(s.item()) for s in sizes]
    return sum(vals[:rank]), sum(vals)


def max_over_ranks(x, device="cpu"):
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return float(x)
    t = torch.tensor([float(x)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def copy_ceiling(gb=1.0, out_fraction=0.36, piece=64 << 20, reps=3, device=None):
    """Copy-only ceiling of the end-to-end leg on this box: every rank moves `gb` GB host -> device and out_fraction * gb
    GB device -> host, from / to pinned memory in 64 MiB pieces on two streams, all ranks at the same time, no kernels.
    -> dict on every rank: GB/s per direction alone and of the input direction while both run ('both'), min / mean over
    ranks and aggregate (the aggregate uses the slowest rank, like the bench's max-over-ranks timing)."""
    import time
    world = dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1
    n_in, n_out = int(gb * 1e9), max(1, int(gb * 1e9 * out_fraction))
    h_in = torch.empty(n_in, dtype=torch.uint8).pin_memory()
    h_in.fill_(65)
    h_out = torch.empty(n_out, dtype=torch.uint8).pin_memory()
    d_in = torch.empty(n_in, dtype=torch.uint8, device="cuda")
    d_out = torch.zeros(n_out, dtype=torch.uint8, device="cuda")
    s_in, s_out = torch.cuda.Stream(), torch.cuda.Stream()

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()

    def run(do_in, do_out):
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
            best = min(best, time.perf_counter() - t)
            barrier()
        return best

    def gather(x):
        if world == 1:
            return [x]
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        out = [torch.zeros_like(t) for _ in range(world)]
        dist.all_gather(out, t)
        return [float(v.item()) for v in out]

    res = {"n_gpus": world, "gb_in_per_rank": gb, "out_fraction": out_fraction}
    for name, di, do, nbytes in (("h2d", True, False, n_in), ("d2h", False, True, n_out), ("both", True, True, n_in)):
        secs = gather(run(di, do))
        res[name] = {"per_rank_gbs_min": nbytes / max(secs) / 1e9, "per_rank_gbs_mean": sum(nbytes / s / 1e9 for s in secs) / world,
                     "aggregate_gbs": world * nbytes / max(secs) / 1e9}
    del h_in, h_out, d_in, d_out
    return res
