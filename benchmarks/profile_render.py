#!/usr/bin/env python
"""One forward render (projection -> intersections -> compositing) of a render-sweep point between cudaProfilerStart /
Stop, for an ncu launch list:

    python benchmarks/profile_render.py --gaussians 8000000 --res 4k &&
    ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --profile-from-start off \\
        --csv --log-file gpurun_out/launches_render.csv python benchmarks/profile_render.py --gaussians 8000000 --res 4k"""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

from qed_splatter_b200.pipeline import FusedSplatStep  # noqa: E402
from qed_splatter_b200.scenes import scene_s1  # noqa: E402

RES = {"720p": (1280, 720), "1080p": (1920, 1080), "1440p": (2560, 1440), "4k": (3840, 2160)}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gaussians", type=int, default=8_000_000)
    ap.add_argument("--res", default="4k")
    ap.add_argument("--mode", default="RGB+ED")
    a = ap.parse_args()
    dev = torch.device("cuda:0")
    W, H = RES[a.res]
    base = scene_s1(N=a.gaussians, targets=False)
    g = {k: getattr(base, k).to(dev) for k in ("means", "quats", "scales", "opacities", "sh", "viewmats")}
    Ks = torch.tensor([[[1200.0 * W / 1920.0, 0, W / 2.0], [0, 1200.0 * W / 1920.0, H / 2.0], [0, 0, 1]]], device=dev)
    fs = FusedSplatStep(dev)
    run = lambda: fs.forward(g["means"], g["quats"], g["scales"], g["opacities"], g["sh"], g["viewmats"], Ks, W, H, 3, render_mode=a.mode)
    for _ in range(3):
        run()
    torch.cuda.synchronize()
    torch.cuda.profiler.start()
    run()
    torch.cuda.synchronize()
    torch.cuda.profiler.stop()
    print(f"profile_render ok: {a.gaussians} Gaussians {a.res} n_isects (gsplat) {int(fs._buf['tiles'][:a.gaussians].sum())} exact {fs.n_isects_exact()}")


if __name__ == "__main__":
    main()
