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
