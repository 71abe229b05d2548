#!/usr/bin/env python
"""Generate tests/golden/*.npz from the CPU oracle (seeded).  The reference repo ships no golden vectors for
this path and its own arithmetic (gsplat) cannot be imported here, so these fixtures pin the ORACLE (and,
through the -m gpu tests, the CUDA path) against regressions; they are not reference outputs.

    python scripts/make_golden.py
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import oracle  # noqa: E402
from qed_splatter_b200.scenes import scene_s0  # noqa: E402


def golden_case(N=1500, C=2, size=64, mode="RGB+ED", seed=42):
    torch.set_num_threads(1)  # fixed reduction order
    s = scene_s0(N=N, C=C, size=size, seed=seed)
    names = ("means", "quats", "scales", "opacities", "sh")
    leaves = {k: getattr(s, k).clone().requires_grad_(True) for k in names}
    render, alpha, info = oracle.rasterization(leaves["means"], leaves["quats"], leaves["scales"], leaves["opacities"], leaves["sh"],
                                               s.viewmats, s.Ks, s.width, s.height, sh_degree=3, render_mode=mode, absgrad=True)
    bg = torch.tensor([0.2, 0.5, 0.8])
    total = 0.0
    for c in range(C):
        rgb, depth = oracle.composite_and_fill(render[c:c + 1], alpha[c:c + 1], bg)
        total = total + oracle.rgb_l1_loss(rgb, s.gt_rgb[c:c + 1]) + oracle.depth_l1_loss(depth, s.gt_depth[c:c + 1], 0.2)
    loss = total / C
    loss.backward()
    out = dict(N=N, C=C, size=size, seed=seed, mode=mode, loss=float(loss), render=render.detach().numpy(), alpha=alpha.detach().numpy(),
               radii=info["radii"].numpy(), tiles_per_gauss=info["tiles_per_gauss"].numpy(), isect_ids=info["isect_ids"].numpy(),
               flatten_ids=info["flatten_ids"].numpy(), isect_offsets=info["isect_offsets"].numpy(), means2d=info["means2d"].detach().numpy(),
               conics=info["conics"].detach().numpy(), depths=info["depths"].detach().numpy(), bg=bg.numpy())
    for k in names:
        out["grad_" + k] = leaves[k].grad.numpy()
    return out


if __name__ == "__main__":
    dst = os.path.join(ROOT, "tests", "golden")
    os.makedirs(dst, exist_ok=True)
    for name, kw in (("s0_small_rgbed", dict(mode="RGB+ED")), ("s0_small_rgbd", dict(mode="RGB+D", N=1200, size=48, seed=7))):
        g = golden_case(**kw)
        path = os.path.join(dst, name + ".npz")
        np.savez_compressed(path, **g)
        print(path, os.path.getsize(path) // 1024, "KiB", "loss", g["loss"], "isects", g["isect_ids"].shape[0])
