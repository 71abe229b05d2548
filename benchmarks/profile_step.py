#!/usr/bin/env python
"""One fused step of the bench workload (BASELINE.json configs[1]) between cudaProfilerStart / Stop, for ncu:

    python benchmarks/profile_step.py &&                                   # must exit 0 without ncu first
    ncu --set full --clock-control none --import-source on --profile-from-start off \\
        -o gpurun_out/r02_step python benchmarks/profile_step.py
    # launch list only:  ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv ...

then, here:  python benchmarks/ncu_dram_bytes.py gpurun_out/r02_step.ncu-rep   (writes profiles/r02_dram_bytes.json and
profiles/r02_step_ncu_summary.txt).  `--steps n` profiles n steps; `--train` profiles trainer steps instead."""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

from qed_splatter_b200.pipeline import FusedSplatStep  # noqa: E402
from qed_splatter_b200.scenes import scene_s1  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gaussians", type=int, default=1_000_000)
    ap.add_argument("--width", type=int, default=1920)
    ap.add_argument("--height", type=int, default=1080)
    ap.add_argument("--steps", type=int, default=1)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--mode", default="RGB+ED")
    ap.add_argument("--train", action="store_true", help="profile trainer.SplatTrainer steps (SSIM + L1 + depth loss, Adam, strategy statistics)")
    a = ap.parse_args()
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(0)
    s = scene_s1(N=a.gaussians, width=a.width, height=a.height).to(dev)
    bg = torch.tensor([0.1, 0.2, 0.3], device=dev)
    if a.train:
        from qed_splatter_b200.trainer import SplatTrainer, TrainConfig

        tr = SplatTrainer(s.means.clone(), s.quats.clone(), torch.log(s.scales), torch.logit(s.opacities.clamp(1e-4, 1 - 1e-4)), s.sh.clone(),
                          cfg=TrainConfig(render_mode=a.mode, refine_every=10 ** 9), backend="cuda")
        tr.step_count = 3000
        fs = tr._fused

        class _Out:
            pass

        def step():
            o = _Out()
            o.loss, _ = tr.step(s.viewmats, s.Ks, s.width, s.height, s.gt_rgb, s.gt_depth, bg)
            o.n_isects = fs._fwd.get("n_isects_real", fs._fwd["M"])
            return o
    else:
        fs = FusedSplatStep(dev)
        grads = {k: torch.zeros_like(getattr(s, k)) for k in ("means", "quats", "scales", "opacities", "sh")}

        def step():
            return fs.step(s.means, s.quats, s.scales, s.opacities, s.sh, s.viewmats, s.Ks, s.width, s.height, 3, s.gt_rgb, s.gt_depth, bg,
                           render_mode=a.mode, grad_out=grads)

    for _ in range(a.warmup):
        out = step()
    torch.cuda.synchronize()
    torch.cuda.profiler.start()
    for _ in range(a.steps):
        out = step()
    torch.cuda.synchronize()
    torch.cuda.profiler.stop()
    print(f"profile_step ok: loss {float(out.loss[0]):.6f} n_isects {out.n_isects} launches/step {fs.launches_per_step}")


if __name__ == "__main__":
    main()
