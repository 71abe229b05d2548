"""-m gpu: edge cases the reference's domain has (SURVEY.md §4 iii): Gaussians behind the camera / on the
frustum edge, a splat covering the whole image, empty tiles, sizes that are not multiples of 16, equal depths,
tiny and empty scenes, thin anisotropic splats (stress for the exact culling), non-contiguous inputs."""
import math

import pytest
import torch

import oracle
from qed_splatter_b200 import ops, rasterization
from qed_splatter_b200.scenes import look_at, scene_s0
from helpers import assert_close_frac, scene_args

pytestmark = pytest.mark.gpu


def _both(a, W, H, cuda, **kw):
    ro, ao, io = oracle.rasterization(**a, width=W, height=H, **kw)
    rg, ag, ig = rasterization(**{k: v.to(cuda) for k, v in a.items()}, width=W, height=H, **kw)
    for k in ("radii", "tiles_per_gauss", "isect_ids", "flatten_ids", "isect_offsets"):
        assert torch.equal(ig[k].cpu(), io[k]), k
    return (ro, ao, io), (rg, ag, ig)


def _scene(N, seed, W, H, spread=1.0, smin=0.01, smax=0.1, eye=(0.0, 0.0, -3.0)):
    g = torch.Generator().manual_seed(seed)
    means = (torch.rand(N, 3, generator=g) * 2 - 1) * spread
    scales = torch.exp(math.log(smin) + (math.log(smax) - math.log(smin)) * torch.rand(N, 3, generator=g))
    quats = torch.randn(N, 4, generator=g)
    opac = 0.02 + 0.97 * torch.rand(N, generator=g)
    colors = torch.rand(N, 3, generator=g)
    vm = look_at(torch.tensor(eye), torch.zeros(3))[None]
    K = torch.tensor([[[0.9 * W, 0, W / 2.0], [0, 0.9 * W, H / 2.0], [0, 0, 1.0]]])
    return dict(means=means, quats=quats, scales=scales, opacities=opac, colors=colors, viewmats=vm, Ks=K)


@pytest.mark.parametrize("W,H", [(100, 60), (17, 33), (16, 16), (250, 9)])
def test_non_multiple_of_16_and_tiny_images(cuda, W, H):
    a = _scene(800, W * 7 + H, W, H)
    (ro, ao, _), (rg, ag, _) = _both(a, W, H, cuda, render_mode="RGB+ED", sh_degree=None)
    assert rg.shape == (1, H, W, 4)
    assert_close_frac(rg, ro, 1e-4, 1e-4, 3e-3, "render")
    assert_close_frac(ag, ao, 1e-4, 1e-4, 3e-3, "alpha")


def test_camera_inside_the_cloud_behind_and_frustum_edge(cuda):
    # camera in the middle of the scene: half of the Gaussians are behind it, many straddle the frustum edge,
    # near ones cover the whole image (radius >> image size)
    W, H = 96, 64
    a = _scene(3000, 11, W, H, spread=2.0, smin=0.02, smax=0.3, eye=(0.1, 0.05, 0.0))
    a["viewmats"] = look_at(torch.tensor([0.1, 0.05, 0.0]), torch.tensor([0.0, 0.0, 2.0]))[None]
    (ro, ao, io), (rg, ag, ig) = _both(a, W, H, cuda, render_mode="RGB+D", sh_degree=None)
    vis = io["radii"][0] > 0
    assert 0 < int(vis.sum()) < 3000 and int(io["radii"].max()) > max(W, H)  # some culled, some gigantic
    assert_close_frac(rg, ro, 1e-4, 1e-4, 5e-3, "render")
    assert_close_frac(ag, ao, 1e-4, 1e-4, 5e-3, "alpha")


def test_equal_depths_and_empty_tiles(cuda):
    # all Gaussians on one plane (identical camera depth -> sort ties decided by index), clustered in a corner
    W, H = 128, 96
    a = _scene(500, 5, W, H)
    a["means"][:, 2] = 0.0
    a["means"][:, :2] = a["means"][:, :2] * 0.2 - 0.6
    (ro, ao, io), (rg, ag, ig) = _both(a, W, H, cuda, render_mode="RGB+ED", sh_degree=None)
    d = io["depths"][io["radii"] > 0]
    assert float(d.max() - d.min()) < 1e-5
    off = io["isect_offsets"].flatten()
    assert int((off[1:] == off[:-1]).sum()) > 10  # many empty tiles
    assert float(ag.min()) == 0.0
    assert_close_frac(rg, ro, 1e-4, 1e-4, 3e-3, "render")


@pytest.mark.parametrize("N", [0, 1, 2])
def test_tiny_and_empty_scenes(cuda, N):
    W, H = 40, 40
    a = _scene(max(N, 1), 3, W, H, spread=0.3)
    a = {k: (v[:N] if k in ("means", "quats", "scales", "opacities", "colors") else v) for k, v in a.items()}
    rg, ag, ig = rasterization(**{k: v.to(cuda) for k, v in a.items()}, width=W, height=H, render_mode="RGB+ED", sh_degree=None)
    assert rg.shape == (1, H, W, 4) and torch.isfinite(rg).all()
    if N == 0:
        assert float(ag.abs().max()) == 0.0 and ig["isect_ids"].numel() == 0
    else:
        ro, ao, io = oracle.rasterization(**a, width=W, height=H, render_mode="RGB+ED", sh_degree=None)
        assert_close_frac(rg, ro, 1e-4, 1e-4, 3e-3, "render")


def test_culling_is_exact_on_thin_tilted_splats(cuda):
    """Needle-like, randomly tilted, mostly faint Gaussians: the worst case for the conservative ellipse test
    (cancellation in the quadratic form).  Culled and un-culled kernels must agree bit for bit."""
    W, H = 320, 200
    g = torch.Generator().manual_seed(21)
    N = 30000
    a = _scene(N, 21, W, H, spread=1.5)
    a["scales"] = torch.stack([torch.full((N,), 0.002), torch.exp(torch.rand(N, generator=g) * 4 - 5), torch.full((N,), 0.003)], -1)
    a["opacities"] = torch.cat([torch.rand(N // 2, generator=g) * 0.02 + 0.003, torch.rand(N - N // 2, generator=g)])
    ga = {k: v.to(cuda) for k, v in a.items()}
    outs = {}
    for tag, cull in (("cull", True), ("nocull", False)):
        ops.set_raster_cull(cull)
        leaves = {k: ga[k].clone().requires_grad_(True) for k in ("means", "quats", "scales", "opacities", "colors")}
        r, al, info = rasterization(**leaves, viewmats=ga["viewmats"], Ks=ga["Ks"], width=W, height=H, render_mode="RGB+ED", sh_degree=None,
                                    absgrad=True)
        info["means2d"].retain_grad()
        (r.sum() + al.sum()).backward()
        outs[tag] = (r.detach(), al.detach(), {k: v.grad.double() for k, v in leaves.items()}, info["means2d"].grad.double(),
                     info["means2d"].absgrad.double())
    ops.set_raster_cull(True)
    assert torch.equal(outs["cull"][0], outs["nocull"][0]) and torch.equal(outs["cull"][1], outs["nocull"][1])
    # gradients: the per-warp sums are identical in both modes; only the order of the float atomics differs.  The
    # screen-space gradients are well conditioned and must agree tightly; the world-space gradients of needle splats
    # are huge cancelling sums (run-to-run noise of the SAME mode reaches 1e-2 of the norm), so they get a loose,
    # element-wise check.
    for i, name in ((3, "v_means2d"), (4, "absgrad")):
        ref = outs["nocull"][i]
        assert_close_frac(outs["cull"][i], ref, 1e-4, 1e-5 * float(ref.abs().mean() + 1e-30), 1e-3, name)
    for k in outs["cull"][2]:
        ref = outs["nocull"][2][k]
        assert_close_frac(outs["cull"][2][k], ref, 5e-2, 1e-3 * float(ref.abs().mean() + 1e-30), 2e-2, f"v_{k}", max_outlier=5e-2)
    outs[True] = outs["cull"]
    # and against the oracle
    ro, ao, _ = oracle.rasterization(**a, width=W, height=H, render_mode="RGB+ED", sh_degree=None)
    assert_close_frac(outs[True][0], ro, 1e-4, 1e-4, 5e-3, "render vs oracle")


def test_pixels_per_lane_variants_agree(cuda):
    """Scalar compositor kernels with 1, 2 or 4 pixels per lane: bitwise-identical forward.  The default packed
    (two-wide fp32) kernels round log2(alpha) once more (add, then fma), so they agree to float tolerance."""
    s = scene_s0(N=8000, C=1, size=160).to(cuda)
    a = scene_args(s)

    def run():
        leaves = {k: a[k].clone().requires_grad_(True) for k in ("means", "quats", "scales", "opacities", "colors")}
        r, al, _ = rasterization(**leaves, viewmats=a["viewmats"], Ks=a["Ks"], width=s.width, height=s.height, render_mode="RGB+ED", sh_degree=3)
        (r * r).sum().backward()
        return r.detach(), al.detach(), leaves["means"].grad

    try:
        ops.set_raster_packed(False)
        ref = None
        for px_f, px_b in ((1, 1), (2, 2), (4, 4)):
            ops.set_raster_px(px_f, px_b)
            cur = run()
            if ref is None:
                ref = cur
            else:
                assert torch.equal(cur[0], ref[0]) and torch.equal(cur[1], ref[1])  # forward is bitwise identical
                assert_close_frac(cur[2], ref[2], 1e-3, 1e-5 * float(ref[2].abs().mean()), 1e-3, "means grad")
        ops.set_raster_packed(True)
        cur = run()
        assert_close_frac(cur[0], ref[0], 1e-5, 1e-5, 1e-3, "packed render")
        assert_close_frac(cur[1], ref[1], 1e-5, 1e-5, 1e-3, "packed alpha")
        assert_close_frac(cur[2], ref[2], 1e-3, 1e-5 * float(ref[2].abs().mean()), 2e-3, "packed means grad")
    finally:
        ops.set_raster_packed(True)
        ops.set_raster_px(4, 4)


def test_non_contiguous_and_misaligned_inputs(cuda):
    s = scene_s0(N=1000, C=1, size=64)
    a = scene_args(s, cuda)
    ref = rasterization(**a, width=64, height=64, render_mode="RGB+D", sh_degree=3)[0]
    big = {k: torch.cat([torch.zeros_like(v[:1]), v], 0)[1:] if v.shape[0] == s.N else v for k, v in a.items()}  # offset views
    big["colors"] = torch.cat([a["colors"], a["colors"]], dim=1)[:, :16]  # non-contiguous
    out = rasterization(**big, width=64, height=64, render_mode="RGB+D", sh_degree=3)[0]
    assert torch.equal(out, ref)
