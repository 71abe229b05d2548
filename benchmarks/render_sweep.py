#!/usr/bin/env python
"""BASELINE.json configs[4]: render-only sweep 0.5M-8M Gaussians x 720p-4K, RGB+ED and depth-only (ED) modes.

Forward only (projection -> intersections -> compositing) through the fused pipeline, CUDA-event timed,
reported per stage with the work counters that make throughput meaningful (n_visible, n_isects, Gaussians
composited per pixel) and, per point, the fraction of the roofline that bounds each stage: projection vs the measured HBM
copy peak (MEASURED_PEAKS.json) on its algorithmic bytes, compositing vs the FP32 peak at the max SM
clock on 30 FLOP per composited (pixel, Gaussian) pair.  One JSON line per (N, resolution, mode) on stdout; `--md` prints a
markdown table.
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
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        peaks = {}
    hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
    fp32_peak = 148 * 128 * 2 * float(peaks.get("sm_max_mhz", 1965.0)) * 1e6 / 1e12
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
                n_vis = int((f["radii"] > 0).sum())
                pairs = fs.count_pairs(which=("fwd",)).get("fwd_pairs_contributing", 0)
                D = f["D"]
                flop_pair = 22.0 + 2.0 * D  # 30 FLOP + 1 ex2 per pair at D = 4 (8 of them the D fused multiply-adds)
                proj_bytes = n * 44.0 + n_vis * ((192.0 if f["n_color"] else 0.0) + 100.0)
                t_proj, t_ras = acc.get("project_fwd"), acc.get("raster_fwd")
                row = {"gaussians": n, "res": res, "width": W, "height": H, "mode": mode, "ms": ms, "mpix_s": W * H / ms / 1e3,
                       "n_visible": n_vis, "n_isects": int(fs._buf["tiles"][:n].sum()), "n_isects_exact": fs.n_isects_exact(),
                       "alpha_mean": float(f["alphas"].mean()), "composited_per_pixel": pairs / float(W * H),
                       "project_hbm_frac": (proj_bytes / (t_proj * 1e-3) / 1e9 / hbm_peak) if t_proj else None,
                       "raster_fp32_frac": (pairs * flop_pair / (t_ras * 1e-3) / 1e12 / fp32_peak) if t_ras else None,
                       "stage_ms": {k: round(v, 4) for k, v in acc.items()}}
                rows.append(row)
                print(json.dumps(row), flush=True)
        del g
        torch.cuda.empty_cache()
    if a.md:
        print(f"\nroofline denominators: HBM {hbm_peak:.0f} GB/s (measured copy), FP32 {fp32_peak:.1f} TFLOP/s (148 SM x 128 lanes x 2 x max SM clock)")
        print("\n| Gaussians | res | mode | ms | Mpix/s | visible | isects (exact) | composited/pixel | project ms (of HBM) | isect ms | composite ms (of FP32) |")
        print("|---|---|---|---|---|---|---|---|---|---|---|")
        for r in rows:
            s = r["stage_ms"]
            isect = s.get("isect_prepare", 0) + s.get("sync", 0) + s.get("isect_fill", 0)
            print(f"| {r['gaussians'] / 1e6:g}M | {r['res']} | {r['mode']} | {r['ms']:.3f} | {r['mpix_s']:.0f} | {r['n_visible']} | {r['n_isects']} ({r['n_isects_exact']}) | "
                  f"{r['composited_per_pixel']:.1f} | {s.get('project_fwd', 0):.3f} ({r['project_hbm_frac'] or 0:.2f}) | {isect:.3f} | "
                  f"{s.get('raster_fwd', 0):.3f} ({r['raster_fp32_frac'] or 0:.2f}) |")


if __name__ == "__main__":
    main()
