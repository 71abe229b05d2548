"""`rasterization` — the drop-in for `gsplat.rendering.rasterization` on the qed-splatter hot path.

The reference calls it at `/root/reference/qed_splatter/model.py:267-288` and uses the result at
`:289-306` (`info["means2d"].retain_grad()`, `info["radii"][0]`, `render[..., :3]`, `render[..., 3:4]`,
`alpha`).  Same keyword names, same return triple, same `info` keys as gsplat 1.4.0.  Arguments that
the reference never passes away from their defaults and that would need another code path
(`packed=True`, `sparse_grad=True`, `covars`, non-pinhole cameras, `distributed=True`) raise
NotImplementedError instead of silently doing something else.
"""
from __future__ import annotations

from typing import Dict, Optional, Tuple

import torch
from torch import Tensor

from . import ops

RENDER_MODES = ("RGB", "D", "ED", "RGB+D", "RGB+ED")


def rasterization(
    means: Tensor,  # [N,3]
    quats: Tensor,  # [N,4] wxyz
    scales: Tensor,  # [N,3]
    opacities: Tensor,  # [N]
    colors: Tensor,  # [N,K,3] SH coefficients (sh_degree given) or [N,3] / [C,N,3] colours
    viewmats: Tensor,  # [C,4,4]
    Ks: Tensor,  # [C,3,3]
    width: int,
    height: int,
    near_plane: float = 0.01,
    far_plane: float = 1e10,
    radius_clip: float = 0.0,
    eps2d: float = 0.3,
    sh_degree: Optional[int] = None,
    packed: bool = False,
    tile_size: int = 16,
    backgrounds: Optional[Tensor] = None,
    render_mode: str = "RGB",
    sparse_grad: bool = False,
    absgrad: bool = False,
    rasterize_mode: str = "classic",
    channel_chunk: int = 32,
    distributed: bool = False,
    camera_model: str = "pinhole",
    covars: Optional[Tensor] = None,
) -> Tuple[Tensor, Tensor, Dict]:
    """-> (render_colors[C,H,W,D], render_alphas[C,H,W,1], info)."""
    if render_mode not in RENDER_MODES:
        raise ValueError(f"render_mode must be one of {RENDER_MODES}, got {render_mode!r}")
    if rasterize_mode not in ("classic", "antialiased"):
        raise ValueError(f"rasterize_mode must be 'classic' or 'antialiased', got {rasterize_mode!r}")
    if packed or sparse_grad or distributed or covars is not None or camera_model != "pinhole":
        raise NotImplementedError(
            "packed / sparse_grad / distributed / covars / non-pinhole cameras are not reachable from "
            "qed_splatter/model.py:267-288 and are out of scope for this path")
    if tile_size != 16:
        raise NotImplementedError("tile_size is 16 on this path (qed_splatter/model.py:243)")
    N = means.shape[0]
    C = viewmats.shape[0]
    assert means.shape == (N, 3), means.shape
    assert quats.shape == (N, 4), quats.shape
    assert scales.shape == (N, 3), scales.shape
    assert opacities.shape == (N,), opacities.shape
    assert viewmats.shape == (C, 4, 4), viewmats.shape
    assert Ks.shape == (C, 3, 3), Ks.shape
    if sh_degree is None:
        assert (colors.dim() == 2 and colors.shape == (N, 3)) or (colors.dim() == 3 and colors.shape == (C, N, 3)), colors.shape
    else:
        assert colors.dim() == 3 and colors.shape[0] == N and colors.shape[2] == 3, colors.shape
        assert (sh_degree + 1) ** 2 <= colors.shape[1], colors.shape

    want_rgb = render_mode in ("RGB", "RGB+D", "RGB+ED")
    want_depth = render_mode in ("D", "ED", "RGB+D", "RGB+ED")
    normalize = render_mode in ("ED", "RGB+ED")

    radii, means2d, depths, conics, comps, cols, opac, tiles_per_gauss, geom = ops.project_gaussians(
        means, quats, scales, opacities, colors if want_rgb else None, viewmats, Ks, width, height, eps2d=eps2d,
        near_plane=near_plane, far_plane=far_plane, radius_clip=radius_clip,
        calc_compensations=(rasterize_mode == "antialiased"), sh_degree=sh_degree, n_color=3 if want_rgb else 0,
        append_depth=want_depth, tile_size=tile_size)

    tile_width, tile_height = ops.tile_grid(width, height, tile_size)
    _, isect_ids, flatten_ids, isect_offsets = ops.isect_tiles(means2d, radii, depths, tile_size, tile_width, tile_height,
                                                               tiles_per_gauss=tiles_per_gauss, return_offsets=True)

    if backgrounds is not None and want_rgb and want_depth:
        backgrounds = torch.cat([backgrounds, torch.zeros(C, 1, device=backgrounds.device, dtype=backgrounds.dtype)], dim=-1)

    render_colors, render_alphas = ops.rasterize_to_pixels(
        means2d, conics, cols, opac, width, height, tile_size, isect_offsets, flatten_ids, backgrounds=backgrounds,
        absgrad=absgrad, geom=geom, normalize_last=normalize)

    info = {
        "camera_ids": None,
        "gaussian_ids": None,
        "radii": radii,
        "means2d": means2d,
        "depths": depths,
        "conics": conics,
        "opacities": opac,
        "tile_width": tile_width,
        "tile_height": tile_height,
        "tiles_per_gauss": tiles_per_gauss,
        "isect_ids": isect_ids,
        "flatten_ids": flatten_ids,
        "isect_offsets": isect_offsets,
        "width": width,
        "height": height,
        "tile_size": tile_size,
        "n_cameras": C,
    }
    return render_colors, render_alphas, info
