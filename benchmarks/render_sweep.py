#!/usr/bin/env python
"""BASELINE.json configs[4]: render-only sweep 0.5M-8M Gaussians x 720p-4K, RGB+ED and depth-only (ED) modes.

Forward only (projection -> intersections -> compositing) through the fused pipeline, CUDA-event timed,
reported per stage with the work counters that make throughput meaningful (n_visible, n_isects, Gaussians
composited per pixel).  One JSON line per (N, resolution, mode) on stdout; `--md` prints a markdown table.
"""
import argparse
import json
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
    ap.add_argument("--gaussians", default="500000,1000000,2000000,4000000,8000000")
    ap.add_argument("--res", default="720p,1080p,1440p,4k")
    ap.add_argument("--modes", default="RGB+ED,ED")
    ap.add_argument("--iters", type=int, default=10)
    ap.add_argument("--md", action="store_true")
    a = ap.parse_args()
    dev = torch.device("cuda:0")
    rows = []
    for n in (int(x) for x in a.gaussians.split(",")):
        base = scene_s1(N=n, targets=False)
        g = {k: getattr(base, k).to(dev) for k in ("means", "quats", "scales", "opacities", "sh", "viewmats")}
        for res in a.res.split(","):
            W, H = RES[res]
            Ks = torch.tensor([[[1200.0 * W / 1920.0, 0, W / 2.0], [0, 1200.0 * W / 1920.0, H / 2.0], [0, 0, 1]]], device=dev)
            for mode in a.modes.split(","):
                fs = FusedSplatStep(dev)
                run = lambda: fs.forward(g["means"], g["quats"], g["scales"], g["opacities"], g["sh"], g["viewmats"], Ks, W, H, 3, render_mode=mode)
                for _ in range(3):
                    run()
                torch.cuda.synchronize()
                acc = {}
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                for _ in range(a.iters):
                    run()
                e1.record()
                torch.cuda.synchronize()
                ms = e0.elapsed_time(e1) / a.iters
                fs.marks = []
                run()
                torch.cuda.synchronize()
                for (_, x), (name, y) in zip(fs.marks[:-1], fs.marks[1:]):
                    acc[name] = x.elapsed_time(y)
                fs.marks = None
                f = fs._fwd
                row = {"gaussians": n, "res": res, "width": W, "height": H, "mode": mode, "ms": ms, "mpix_s": W * H / ms / 1e3,
                       "n_visible": int((f["radii"] > 0).sum()), "n_isects": f["M"], "alpha_mean": float(f["alphas"].mean()),
                       "stage_ms": {k: round(v, 4) for k, v in acc.items()}}
                rows.append(row)
                print(json.dumps(row), flush=True)
        del g
        torch.cuda.empty_cache()
    if a.md:
        print("\n| Gaussians | res | mode | ms | Mpix/s | visible | isects | project | isect | composite |")
        print("|---|---|---|---|---|---|---|---|---|---|")
        for r in rows:
            s = r["stage_ms"]
            isect = s.get("isect_prepare", 0) + s.get("sync", 0) + s.get("isect_fill", 0)
            print(f"| {r['gaussians'] / 1e6:g}M | {r['res']} | {r['mode']} | {r['ms']:.3f} | {r['mpix_s']:.0f} | {r['n_visible']} | {r['n_isects']} | "
                  f"{s.get('project_fwd', 0):.3f} | {isect:.3f} | {s.get('raster_fwd', 0):.3f} |")


if __name__ == "__main__":
    main()
