#!/usr/bin/env python
"""Multi-GPU check (run under torchrun on >= 2 B200s; NOT collected by pytest -- bench.py runs the same check at N > 1,
outside its timed regions, and prints it as `multi_gpu_check`, so it lands in the driver's SCALE record):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 tests/multi_gpu_check.py

1. N ranks x 1 view == 1 rank x N views: the gradient arena summed over the ranks -- by NCCL and by this library's NVLink
   path (view-colour exchange + one all-reduce kernel, comm.ViewShardedGradients) -- equals the single-process gradient
   of the whole view batch (SURVEY.md section 8e "Check"); the summed arena is bit-identical on every rank.
2. 3 trainer steps: all replicas hold bit-identical parameters, equal (float tolerance) to a single-process trainer fed
   all N views."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

import bench  # noqa: E402


def main():
    world, rank, local = int(os.environ["WORLD_SIZE"]), int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    res = bench.run_multi_gpu_check(torch, dist, dev, rank, world)
    if rank == 0:
        print(json.dumps(res))
    assert res["ok_all_ranks"], res
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
