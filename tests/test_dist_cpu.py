"""world_size-2 gloo test of the multi-rank host logic (no GPU): region sharding as the reference does it
(phyNGSC.cpp:113-124), per-rank outputs placed by an exclusive scan of their sizes, and the resulting file
being exactly the rank-major .ngsc the C++ driver writes.  The per-rank bytes come from the oracle here (it is
the checker; the GPU path is compared against the same oracle in the -m gpu tests)."""
import os
import tempfile

import numpy as np
import torch.distributed as dist
import torch.multiprocessing as mp

from phyngsc_b200 import api, container, synth
from phyngsc_b200 import dist as pdist


def _worker(rank, world, path, port, shape):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle import phy_oracle as O
    data = synth.fastq(shape, 77, target_bytes=900_000)
    start, end = api.region_slice(data.size, world, rank, slack=4096)
    assert start == rank * (data.size // world) and end <= data.size
    mine = O.compress_rank(data, world, rank, window_bytes=128 * 1024, block_bytes=64 * 1024)
    blob = b"".join(mine["blocks"])
    off, total = pdist.exscan_bytes(len(blob))
    fd = os.open(path, os.O_RDWR | os.O_CREAT, 0o644)
    os.pwrite(fd, blob, off)
    # footer inputs travel to rank 0 exactly like the driver's MPI_Gather
    info = [None] * world
    dist.all_gather_object(info, (len(mine["blocks"]), len(mine["subblocks"]), mine["wr_overlap"], mine["last_block_size"]))
    if rank == 0:
        order = [r for r in range(world) for _ in range(info[r][0])]
        foot = api.make_footer(world, data.size, len(order), sum(i[1] for i in info), [i[2] for i in info], order, [i[3] for i in info])
        os.pwrite(fd, foot, total)
    os.close(fd)
    assert pdist.max_over_ranks(rank) == world - 1
    dist.barrier()
    dist.destroy_process_group()


def test_two_ranks_exscan_offsets_give_a_valid_rank_major_ngsc(oracle):
    api.lib()
    shape = "100bp"
    with tempfile.TemporaryDirectory() as d:
        path = os.path.join(d, "out.ngsc")
        mp.spawn(_worker, args=(2, path, 29517 + os.getpid() % 2000, shape), nprocs=2, join=True)
        data = synth.fastq(shape, 77, target_bytes=900_000)
        ng = container.read_ngsc(path, block_bytes=64 * 1024)
        assert ng["footer"]["np"] == 2 and ng["footer"]["fastq_size"] == data.size
        for r in range(2):
            want = oracle.compress_rank(data, 2, r, window_bytes=128 * 1024, block_bytes=64 * 1024)
            assert ng["per_rank_subblocks"][r] == want["subblocks"]
        assert ng["footer"]["block_order"] == sorted(ng["footer"]["block_order"])
