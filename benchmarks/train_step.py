#!/usr/bin/env python
"""BASELINE.json configs[2]/[3]: the qed-splatter training step (render + depth/RGB loss + backward + gradient
all-reduce + fused Adam + strategy statistics, densify on schedule), view-sharded one view per GPU.

    python benchmarks/train_step.py --scene s2 --gaussians 3000000                      # 1 GPU
    torchrun --nproc-per-node 8 --master-addr 127.0.0.1 benchmarks/train_step.py --scene s3 --gaussians 6000000 --refine-every 100

Prints one JSON line (rank 0): train iters/s (max over ranks, device timed), Mpix/s, Gaussian count trace.
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

from qed_splatter_b200.scenes import scene_s2, scene_s3  # noqa: E402
from qed_splatter_b200.trainer import SplatTrainer, TrainConfig  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--scene", default="s2", choices=["s2", "s3"])
    ap.add_argument("--gaussians", type=int, default=3_000_000)
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--refine-every", type=int, default=100)
    ap.add_argument("--start-step", type=int, default=600, help="schedule position (SH degree / densify windows)")
    ap.add_argument("--comm", default="auto", choices=["auto", "exchange", "nccl"], help="TrainConfig.comm (N > 1): this library's NVLink path or NCCL")
    ap.add_argument("--comm-chunks", type=int, default=None, help="override TrainConfig.comm_chunks")
    ap.add_argument("--no-chunk-bwd", action="store_true")
    a = ap.parse_args()
    world, rank, local = int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    make = scene_s2 if a.scene == "s2" else scene_s3
    s = make(N=a.gaussians, C=1, view_offset=rank, total_views=world)
    cfg = TrainConfig(refine_every=a.refine_every, render_mode="RGB+D", comm=a.comm)
    if a.comm_chunks is not None:
        cfg.comm_chunks = a.comm_chunks
    cfg.chunk_project_bwd = not a.no_chunk_bwd
    tr = SplatTrainer(s.means.to(dev), s.quats.to(dev), torch.log(s.scales).to(dev), torch.logit(s.opacities.clamp(1e-4, 1 - 1e-4)).to(dev),
                      s.sh.to(dev), cfg=cfg, rank=rank, world_size=world, backend="cuda")
    tr.step_count = a.start_step
    vm, Ks, rgb, depth = s.viewmats.to(dev), s.Ks.to(dev), s.gt_rgb.to(dev), s.gt_depth.to(dev)
    bg = torch.tensor([0.1, 0.2, 0.3], device=dev)
    W, H = s.width, s.height
    del s

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(a.warmup):
        tr.step(vm, Ks, W, H, rgb, depth, bg, total_views=world)
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    counts, refines = [tr.arena.N], []
    e0.record()
    for _ in range(a.steps):
        loss, info = tr.step(vm, Ks, W, H, rgb, depth, bg, total_views=world)
        if info is not None:
            refines.append(info)
            counts.append(tr.arena.N)
    e1.record()
    barrier()
    t = torch.tensor([e0.elapsed_time(e1)], device=dev)
    n = torch.tensor([tr.arena.N], device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        nmin, nmax = n.clone(), n.clone()
        dist.all_reduce(nmin, op=dist.ReduceOp.MIN)
        dist.all_reduce(nmax, op=dist.ReduceOp.MAX)
        assert int(nmin) == int(nmax), "replicas diverged"
    ms = float(t) / a.steps
    identical = None
    if world > 1:  # replicas hold bit-identical parameters (also after the refine steps inside the window)
        p_ = tr.arena.param
        chk = torch.stack([p_.double().sum(), p_.view(torch.int32).sum().double(), p_[::101].double().abs().sum()])
        allc = [torch.empty_like(chk) for _ in range(world)]
        dist.all_gather(allc, chk)
        identical = all(torch.equal(allc[0], c_) for c_ in allc)
    if rank == 0:
        print(json.dumps({"scene": a.scene, "gaussians_start": a.gaussians, "gaussians_end": tr.arena.N, "n_gpus": world, "steps": a.steps,
                          "ms_per_step": ms, "train_iters_per_s": 1e3 / ms, "views_per_s": world * 1e3 / ms,
                          "mpix_per_s": world * W * H / ms / 1e3, "comm": tr.comm, "comm_chunks": cfg.comm_chunks if tr.comm == "nccl" else None,
                          "replicas_bit_identical": identical, "n_refines_in_window": len(refines), "overflow_repeats": tr._fused.overflow_repeats, "width": W, "height": H, "refines": refines[:4], "loss": [float(x) for x in loss.tolist()],
                          "n_isects_last": tr._fused._fwd.get("n_isects_real", tr._fused._fwd["M"])}))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
