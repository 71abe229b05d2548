"""Test helper: execute the UNMODIFIED reference model (`qed_splatter/model.py`: QEDSplatterModel.get_outputs /
get_loss_dict, model.py:73-118, 199-321) on top of a chosen `gsplat.rendering.rasterization`.

The reference package is taken from (first that exists)
  $QED_REFERENCE_PATH, /root/reference (the container the round is built in), <repo>/baseline/_ref
where `baseline/_ref` is the offline `pip install --no-deps --target` of the reference that `__graft_entry__.build()`
makes when /root/reference is present (git-ignored, travels to the GPU box).  Its third-party imports resolve to the
stand-ins under tests/stubs (see tests/stubs/README.md); `gsplat` resolves either to the product's import shim
(`qed_splatter_b200/shim`, CUDA) or to tests/stubs_oracle (the CPU oracle).
"""
from __future__ import annotations

import contextlib
import importlib
import os
import sys
from typing import Optional

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_PURGE = ("gsplat", "qed_splatter", "nerfstudio", "torchmetrics")


def reference_root() -> Optional[str]:
    for p in (os.environ.get("QED_REFERENCE_PATH"), "/root/reference", os.path.join(ROOT, "baseline", "_ref")):
        if p and os.path.isfile(os.path.join(p, "qed_splatter", "model.py")):
            return p
    return None


def _purge():
    for m in [k for k in sys.modules if any(k == p or k.startswith(p + ".") for p in _PURGE)]:
        del sys.modules[m]


@contextlib.contextmanager
def reference_modules(backend: str):
    """backend "cuda": gsplat -> qed_splatter_b200 (the product shim); "oracle": gsplat -> the CPU oracle.
    Yields the freshly imported `qed_splatter.model` module."""
    ref = reference_root()
    assert ref is not None, "reference package not found"
    shim = os.path.join(ROOT, "qed_splatter_b200", "shim") if backend == "cuda" else os.path.join(ROOT, "tests", "stubs_oracle")
    paths = [os.path.join(ROOT, "tests", "stubs"), shim, ref]
    _purge()
    sys.path[:0] = paths
    patched = None
    try:
        if backend == "oracle" and not torch.cuda.is_available():
            # model.py:247 calls `.cuda()` on the intrinsics; without a GPU the CPU-oracle run keeps them where they are
            patched = torch.Tensor.cuda
            torch.Tensor.cuda = lambda self, *a, **k: self
        yield importlib.import_module("qed_splatter.model")
    finally:
        if patched is not None:
            torch.Tensor.cuda = patched
        for p in paths:
            sys.path.remove(p)
        _purge()


def c2w_from_viewmat(viewmats: torch.Tensor) -> torch.Tensor:
    """nerfstudio camera-to-world [C,3,4] (OpenGL axes) whose `get_viewmat` (model.py:22-38) is `viewmats`."""
    inv = torch.linalg.inv(viewmats.double())
    R = inv[:, :3, :3] * torch.tensor([[[1.0, -1.0, -1.0]]], dtype=torch.float64)
    return torch.cat([R, inv[:, :3, 3:4]], dim=-1).to(torch.float32)


def build_model(mod, scene, device, step: int = 3000, num_downscales: int = 0, background: str = "white", rasterize_mode: str = "classic"):
    """QEDSplatterModel over the scene's Gaussians, parameters in splatfacto's stored form."""
    cfg = mod.QEDSplatterModelConfig(num_downscales=num_downscales, background_color=background, rasterize_mode=rasterize_mode)
    params = dict(means=scene.means, scales=torch.log(scene.scales), quats=scene.quats * 1.7,  # un-normalised on purpose (model.py:269)
                  features_dc=scene.sh[:, 0, :].contiguous(), features_rest=scene.sh[:, 1:, :].contiguous(),
                  opacities=torch.logit(scene.opacities)[:, None])
    model = cfg.setup(gauss_params={k: v.to(device) for k, v in params.items()})
    model.step = step
    return model.to(device)


def make_camera(scene, cam: int, device):
    from nerfstudio.cameras.cameras import Cameras

    K = scene.Ks[cam]
    return Cameras(c2w_from_viewmat(scene.viewmats[cam:cam + 1]).to(device), float(K[0, 0]), float(K[1, 1]), float(K[0, 2]), float(K[1, 2]),
                   scene.width, scene.height)


def oracle_reference_step(scene, cam: int, c2w: torch.Tensor, background: torch.Tensor, step: int = 3000, mask=None, down: int = 1,
                          depth_lambda: float = 0.2, ssim_lambda: float = 0.2, rasterize_mode: str = "classic"):
    """The same training step written with the oracle only (CPU): viewmat, activations, rasterization (RGB+D), composite /
    depth fill, splatfacto RGB loss + qed-splatter depth loss.  Returns (outputs, loss_dict, leaf parameters)."""
    import torch.nn.functional as F

    import oracle

    leaves = dict(means=scene.means.clone(), scales=torch.log(scene.scales), quats=(scene.quats * 1.7).clone(),
                  features_dc=scene.sh[:, 0, :].clone(), features_rest=scene.sh[:, 1:, :].clone(), opacities=torch.logit(scene.opacities)[:, None])
    leaves = {k: v.detach().clone().requires_grad_(True) for k, v in leaves.items()}
    viewmat = oracle.get_viewmat(c2w)
    K = scene.Ks[cam:cam + 1].clone()
    K[:, :2, :] = K[:, :2, :] / down
    W, H = int(scene.width / down + 0.5), int(scene.height / down + 0.5)
    colors = torch.cat((leaves["features_dc"][:, None, :], leaves["features_rest"]), dim=1)
    render, alpha, info = oracle.rasterization(
        leaves["means"], leaves["quats"] / leaves["quats"].norm(dim=-1, keepdim=True), torch.exp(leaves["scales"]),
        torch.sigmoid(leaves["opacities"]).squeeze(-1), colors, viewmat, K, W, H, render_mode="RGB+D",
        sh_degree=min(step // 1000, 3), absgrad=True, rasterize_mode=rasterize_mode)
    rgb, depth = oracle.composite_and_fill(render, alpha, background)
    rgb, depth = rgb.squeeze(0), depth.squeeze(0)

    def gt(img):
        img = img.float() / 255.0 if img.dtype == torch.uint8 else img
        if down > 1:
            w = (1.0 / (down * down)) * torch.ones((1, 1, down, down))
            img = F.conv2d(img.float().permute(2, 0, 1)[:, None, ...], w, stride=down).squeeze(1).permute(1, 2, 0)
        return img

    gt_rgb, gt_depth = gt(scene.gt_rgb[cam]), gt(scene.gt_depth[cam])
    m = gt(mask) if mask is not None else None
    pred_l, gt_l = (rgb * m, gt_rgb * m) if m is not None else (rgb, gt_rgb)
    loss = {"main_loss": oracle.rgb_loss(pred_l[None], gt_l[None], ssim_lambda),
            "depth_loss": oracle.depth_l1_loss(depth, gt_depth, depth_lambda, mask=m)}
    return {"rgb": rgb, "depth": depth, "accumulation": alpha.squeeze(0), "viewmat": viewmat, "info": info}, loss, leaves
