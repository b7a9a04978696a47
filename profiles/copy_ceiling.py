"""Copy-only ceiling of the end-to-end leg (phyngsc_b200.dist.copy_ceiling): every rank moves what bench.py's e2e leg moves -- its
region host -> device and the payloads device -> host, pinned memory, 64 MiB pieces, two streams -- with no kernels, all ranks at once.
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P profiles/copy_ceiling.py [GB per rank] [out fraction]
Rank 0 prints one JSON line."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

from phyngsc_b200 import dist as pdist  # noqa: E402

rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
res = pdist.copy_ceiling(float(sys.argv[1]) if len(sys.argv) > 1 else 2.0, float(sys.argv[2]) if len(sys.argv) > 2 else 0.36)
if rank == 0:
    print(json.dumps(res))
if world > 1:
    dist.destroy_process_group()
