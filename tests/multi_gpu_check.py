#!/usr/bin/env python
"""Multi-GPU check (run under torchrun on >= 2 B200s; NOT collected by pytest):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 tests/multi_gpu_check.py

1. N ranks x 1 view == 1 rank x N views: the all-reduced gradient arena equals the single-process gradient of the
   whole view batch (SURVEY.md §8e "Check").
2. the chunk-pipelined step (SH ranges all-reduced / Adam-stepped while the projection backward continues) gives
   the same parameters as the plain step, and all replicas hold identical parameters after 3 steps.
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

from qed_splatter_b200.pipeline import FusedSplatStep  # noqa: E402
from qed_splatter_b200.scenes import scene_s0  # noqa: E402
from qed_splatter_b200.trainer import SplatTrainer, TrainConfig  # noqa: E402


def main():
    world, rank, local = int(os.environ["WORLD_SIZE"]), int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    s = scene_s0(N=20000, C=world, size=160).to(dev)
    bg = torch.tensor([0.2, 0.3, 0.4], device=dev)
    sl = slice(rank, rank + 1)
    mine = dict(viewmats=s.viewmats[sl].contiguous(), Ks=s.Ks[sl].contiguous(), gt_rgb=s.gt_rgb[sl].contiguous(), gt_depth=s.gt_depth[sl].contiguous())

    # 1. gradient equivalence
    fs = FusedSplatStep(dev)
    whole = fs.step(s.means, s.quats, s.scales, s.opacities, s.sh, s.viewmats, s.Ks, s.width, s.height, 3, s.gt_rgb, s.gt_depth, bg)
    ref = {k: v.clone() for k, v in whole.grads.items()}
    part = fs.step(s.means, s.quats, s.scales, s.opacities, s.sh, mine["viewmats"], mine["Ks"], s.width, s.height, 3, mine["gt_rgb"],
                   mine["gt_depth"], bg, grad_scale=1.0 / world)
    for k, v in part.grads.items():
        g = v.clone()
        dist.all_reduce(g)
        err = float((g - ref[k]).abs().max()) / (float(ref[k].abs().mean()) + 1e-20)
        assert err < 1e-2 and float(((g - ref[k]).abs() > 1e-4 * ref[k].abs() + 1e-5 * ref[k].abs().mean()).float().mean()) < 1e-3, (k, err)

    # 2. pipelined vs plain trainer step, replicas identical
    outs = {}
    for chunks in (1, 4):
        cfg = TrainConfig(comm_chunks=chunks)
        tr = SplatTrainer(s.means.clone(), s.quats.clone(), torch.log(s.scales), torch.logit(s.opacities), s.sh.clone(), cfg=cfg, rank=rank,
                          world_size=world, backend="cuda")
        tr.step_count = 3000  # SH degree 3
        for _ in range(3):
            tr.step(mine["viewmats"], mine["Ks"], s.width, s.height, mine["gt_rgb"], mine["gt_depth"], bg, total_views=world)
        outs[chunks] = tr.arena.param.clone()
        gathered = [torch.empty_like(tr.arena.param) for _ in range(world)]
        dist.all_gather(gathered, tr.arena.param)
        for r in range(1, world):
            assert torch.equal(gathered[0], gathered[r]), f"replica {r} diverged (chunks={chunks})"
    d = (outs[1] - outs[4]).abs()
    upd = (outs[1] - torch.cat([t.reshape(-1) for t in ()] or [outs[1] * 0])).abs()  # noqa: F841
    assert float(d.max()) < 2e-3 and float((d > 1e-5).float().mean()) < 1e-3, (float(d.max()), float((d > 1e-5).float().mean()))
    if rank == 0:
        print(f"multi_gpu_check ok on {world} GPUs: sharded == batched gradients; pipelined == plain step (max |dparam| {float(d.max()):.2e}); replicas identical")
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
