#!/usr/bin/env python
"""Generate tests/golden/ref_model_*.npz by EXECUTING THE REFERENCE'S OWN CODE in this container:

    /root/reference/qed_splatter/model.py  (unmodified)  QEDSplatterModel.get_outputs + get_loss_dict + backward

with nerfstudio / torchmetrics resolved to the minimal stand-ins under tests/stubs and `gsplat.rendering.rasterization`
resolved to the CPU oracle (the un-vendored gsplat cannot be installed here).  The fixtures therefore hold outputs of
reference-held lines (viewmat construction model.py:22-38, activations :241,269-271, composite / clamp / depth fill
:295-306, depth loss incl. `batch["mask"]` :87-116, splatfacto's RGB loss through super()) on seeded inputs; the
`-m gpu` tests compare the CUDA path against them on the B200, where /root/reference does not exist.

    python scripts/make_golden_reference_model.py
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import reference_model as rm  # noqa: E402
from qed_splatter_b200.scenes import scene_s0  # noqa: E402

CASES = {
    # name: (scene kwargs, camera, mask kind, step)
    "ref_model_plain": (dict(N=1500, C=2, size=48, seed=42), 1, None, 3000),
    "ref_model_mask": (dict(N=1500, C=2, size=48, seed=42), 0, "float", 3000),
    "ref_model_boolmask_sh1": (dict(N=1200, C=1, size=40, seed=7), 0, "bool", 1000),
}


def case_mask(kind, s, seed=5):
    if kind is None:
        return None
    m = torch.rand(s.height, s.width, 1, generator=torch.Generator().manual_seed(seed)) > 0.3
    return m.float() if kind == "float" else m


def run_case(scene_kw, cam, mask_kind, step):
    torch.set_num_threads(1)
    s = scene_s0(**scene_kw)
    mask = case_mask(mask_kind, s)
    with rm.reference_modules("oracle") as mod:
        model = rm.build_model(mod, s, "cpu", step=step)
        model.train()
        camera = rm.make_camera(s, cam, "cpu")
        out = model.get_outputs(camera)
        batch = {"image": s.gt_rgb[cam], "depth_image": s.gt_depth[cam]}
        if mask is not None:
            batch["mask"] = mask
        loss = model.get_loss_dict(out, batch)
        sum(loss.values()).backward()
        res = dict(cam=cam, step=step, mask_kind=str(mask_kind), background=model._get_background_color().numpy(),
                   c2w=camera.camera_to_worlds.numpy(), viewmat=mod.get_viewmat(camera.camera_to_worlds).numpy(),
                   rgb=out["rgb"].detach().numpy(), depth=out["depth"].detach().numpy(), accumulation=out["accumulation"].detach().numpy(),
                   radii=model.radii.numpy(), main_loss=float(loss["main_loss"]), depth_loss=float(loss["depth_loss"]))
        for k, v in scene_kw.items():
            res["scene_" + k] = v
        for k in model.gauss_params:
            res["grad_" + k] = model.gauss_params[k].grad.numpy()
    return res


if __name__ == "__main__":
    assert rm.reference_root() is not None, "needs the reference package (/root/reference)"
    dst = os.path.join(ROOT, "tests", "golden")
    for name, args in CASES.items():
        g = run_case(*args)
        path = os.path.join(dst, name + ".npz")
        np.savez_compressed(path, **g)
        print(path, os.path.getsize(path) // 1024, "KiB", "main", g["main_loss"], "depth", g["depth_loss"])
