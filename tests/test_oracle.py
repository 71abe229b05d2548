"""CPU: self-checks of the oracle (parity is UNPINNED by the reference — it ships no tests or vectors for
this path — so the oracle is checked against first principles instead) + the committed golden fixtures."""
import math
import os

import numpy as np
import pytest
import torch

import oracle
from oracle import torch_impl as ti
from qed_splatter_b200.scenes import scene_s0

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")


def _tiny(N=10, C=2, size=32, seed=3, dtype=torch.float64):
    s = scene_s0(N=N, C=C, size=size, seed=seed)
    s.scales = s.scales * 3.0  # bigger splats so every pixel sees a few Gaussians
    conv = lambda t: t.to(dtype)
    return s, dict(means=conv(s.means), quats=conv(s.quats), scales=conv(s.scales), opacities=conv(s.opacities), colors=conv(s.sh),
                   viewmats=conv(s.viewmats), Ks=conv(s.Ks))


def test_sh_bases_orthonormal():
    g = torch.Generator().manual_seed(0)
    d = torch.randn(400_000, 3, generator=g, dtype=torch.float64)
    d = d / d.norm(dim=-1, keepdim=True)
    B = torch.stack(ti.sh_bases(d[:, 0], d[:, 1], d[:, 2], 3), dim=-1)  # [S,16]
    gram = 4 * math.pi * (B.T @ B) / B.shape[0]
    assert torch.allclose(gram, torch.eye(16, dtype=torch.float64), atol=2e-2)


def test_projection_matches_matrix_form():
    """The pinned scalar op order equals the textbook matrix form (J W Sigma W^T J^T) in float64."""
    s, a = _tiny(N=50)
    radii, means2d, depths, conics, comps = oracle.fully_fused_projection(a["means"], a["quats"], a["scales"], a["viewmats"], a["Ks"],
                                                                        s.width, s.height, calc_compensations=True)
    q = a["quats"] / a["quats"].norm(dim=-1, keepdim=True)
    w, x, y, z = q.unbind(-1)
    R = torch.stack([1 - 2 * (y * y + z * z), 2 * (x * y - w * z), 2 * (x * z + w * y), 2 * (x * y + w * z), 1 - 2 * (x * x + z * z),
                     2 * (y * z - w * x), 2 * (x * z - w * y), 2 * (y * z + w * x), 1 - 2 * (x * x + y * y)], -1).reshape(-1, 3, 3)
    M = R * a["scales"][:, None, :]
    Sigma = M @ M.transpose(1, 2)
    for c in range(s.C):
        W, t = a["viewmats"][c, :3, :3], a["viewmats"][c, :3, 3]
        p = a["means"] @ W.T + t
        Sc = W @ Sigma @ W.T
        fx, fy, cx, cy = a["Ks"][c, 0, 0], a["Ks"][c, 1, 1], a["Ks"][c, 0, 2], a["Ks"][c, 1, 2]
        vis = radii[c] > 0
        J = torch.zeros(len(p), 2, 3, dtype=torch.float64)
        J[:, 0, 0] = fx / p[:, 2]
        J[:, 0, 2] = -fx * p[:, 0] / p[:, 2] ** 2
        J[:, 1, 1] = fy / p[:, 2]
        J[:, 1, 2] = -fy * p[:, 1] / p[:, 2] ** 2
        cov = J @ Sc @ J.transpose(1, 2) + 0.3 * torch.eye(2, dtype=torch.float64)
        inv = torch.linalg.inv(cov)
        # visible Gaussians of this scene are well inside the frustum, so the clamp in J is inactive
        assert torch.allclose(conics[c][vis], torch.stack([inv[:, 0, 0], inv[:, 0, 1], inv[:, 1, 1]], -1)[vis], rtol=1e-9, atol=1e-12)
        assert torch.allclose(means2d[c][vis], torch.stack([fx * p[:, 0] / p[:, 2] + cx, fy * p[:, 1] / p[:, 2] + cy], -1)[vis], rtol=1e-12)
        assert torch.allclose(depths[c][vis], p[:, 2][vis])
    assert radii.dtype == torch.int32 and int((radii > 0).sum()) > 0


def test_camera_positions_is_inverse():
    s, a = _tiny()
    pos = ti.camera_positions(a["viewmats"])
    assert torch.allclose(pos, torch.linalg.inv(a["viewmats"])[:, :3, 3], atol=1e-12)


def test_rasterization_autograd_vs_finite_differences():
    s, a = _tiny(N=8, C=1, size=24, seed=5)
    names = ["means", "quats", "scales", "opacities", "colors"]
    g = torch.Generator().manual_seed(1)

    def f(**over):
        args = {**a, **over}
        r, al, _ = oracle.rasterization(**args, width=s.width, height=s.height, sh_degree=3, render_mode="RGB+ED")
        return r, al

    leaves = {k: a[k].clone().requires_grad_(True) for k in names}
    r, al = f(**leaves)
    wr = torch.randn(r.shape, generator=g, dtype=torch.float64)
    wa = torch.randn(al.shape, generator=g, dtype=torch.float64)
    ((r * wr).sum() + (al * wa).sum()).backward()
    eps = 1e-6
    for k in names:
        d = torch.randn(a[k].shape, generator=g, dtype=torch.float64)
        d = d / d.norm()
        rp, ap = f(**{k: a[k] + eps * d})
        rm, am = f(**{k: a[k] - eps * d})
        fd = (((rp - rm) * wr).sum() + ((ap - am) * wa).sum()) / (2 * eps)
        an = (leaves[k].grad * d).sum()
        assert abs(float(fd - an)) <= 1e-5 * max(1.0, abs(float(an))), (k, float(fd), float(an))


def test_explicit_backward_matches_autograd():
    s, a = _tiny(N=40, C=2, size=40, seed=9)
    radii, means2d, depths, conics, _ = oracle.fully_fused_projection(a["means"], a["quats"], a["scales"], a["viewmats"], a["Ks"], s.width, s.height)
    tw = th = math.ceil(s.width / 16)
    _, ids, flat = oracle.isect_tiles(means2d, radii, depths, 16, tw, th)
    off = oracle.isect_offset_encode(ids, s.C, tw, th)
    g = torch.Generator().manual_seed(2)
    cols = torch.rand(s.C, s.N, 4, generator=g, dtype=torch.float64)
    op = a["opacities"][None].expand(s.C, -1).contiguous()
    bg = torch.rand(s.C, 4, generator=g, dtype=torch.float64)
    leaves = [t.clone().requires_grad_(True) for t in (means2d, conics, cols, op)]
    r, al, last = oracle.rasterize_to_pixels(*leaves, s.width, s.height, 16, off, flat, backgrounds=bg)
    vr = torch.randn(r.shape, generator=g, dtype=torch.float64)
    va = torch.randn(al.shape, generator=g, dtype=torch.float64)
    ((r * vr).sum() + (al * va).sum()).backward()
    v_m, v_abs, v_cn, v_co, v_op = oracle.rasterize_to_pixels_bwd(means2d, conics, cols, op, s.width, s.height, 16, off, flat, vr, va[..., 0], backgrounds=bg)
    for got, leaf in zip((v_m, v_cn, v_co, v_op), leaves):
        assert torch.allclose(got, leaf.grad, rtol=1e-9, atol=1e-12)
    assert bool((v_abs >= v_m.abs() - 1e-12).all())


def test_compositing_invariants_and_modes():
    s, a = _tiny(N=60, C=2, size=48, seed=4, dtype=torch.float32)
    outs = {m: oracle.rasterization(**a, width=s.width, height=s.height, sh_degree=3, render_mode=m) for m in ("RGB", "D", "ED", "RGB+D", "RGB+ED")}
    alpha = outs["RGB"][1]
    assert float(alpha.min()) >= 0.0 and float(alpha.max()) < 1.0
    assert torch.allclose(outs["RGB+D"][0][..., :3], outs["RGB"][0], atol=1e-6) and torch.allclose(outs["RGB+D"][0][..., 3:], outs["D"][0], atol=1e-6)
    ed, d = outs["ED"][0], outs["D"][0]
    assert torch.allclose(ed * alpha.clamp(min=1e-10), d, rtol=1e-5, atol=1e-6)
    info = outs["RGB+ED"][2]
    vis = info["radii"] > 0
    zmin, zmax = float(info["depths"][vis].min()), float(info["depths"][vis].max())
    covered = alpha[..., 0] > 1e-3
    assert float(ed[..., 0][covered].min()) >= zmin - 1e-4 and float(ed[..., 0][covered].max()) <= zmax + 1e-4
    # background: render = acc + T*bg
    bg = torch.tensor([[0.3, 0.6, 0.9], [0.1, 0.2, 0.3]])
    rb = oracle.rasterization(**a, width=s.width, height=s.height, sh_degree=3, render_mode="RGB", backgrounds=bg)[0]
    assert torch.allclose(rb, outs["RGB"][0] + (1 - alpha) * bg[:, None, None, :], atol=1e-6)


@pytest.mark.parametrize("W,H", [(64, 48), (72, 40), (16, 16), (100, 9)])
def test_isect_properties(W, H):
    g = torch.Generator().manual_seed(W * 100 + H)
    C, N = 3, 400
    means2d = (torch.rand(C, N, 2, generator=g) * 1.4 - 0.2) * torch.tensor([W, H])
    radii = torch.randint(0, 30, (C, N), generator=g, dtype=torch.int32)
    depths = torch.rand(C, N, generator=g) * 5 + 0.1
    depths[:, ::5] = 1.0  # ties
    tw, th = math.ceil(W / 16), math.ceil(H / 16)
    tiles, ids, flat = oracle.isect_tiles(means2d, radii, depths, 16, tw, th)
    off = oracle.isect_offset_encode(ids, C, tw, th)
    assert ids.numel() == int(tiles.sum()) == flat.numel()
    assert bool((ids[1:] >= ids[:-1]).all())
    nb = (tw * th).bit_length()
    cam = ids >> (32 + nb)
    tile = (ids >> 32) & ((1 << nb) - 1)
    assert bool((cam == flat.long() // N).all()) and int(tile.max()) < tw * th
    assert bool(((ids & 0xFFFFFFFF) == (depths.flatten().view(torch.int32).long() & 0xFFFFFFFF)[flat.long()]).all())
    # ties keep ascending flat index
    same = ids[1:] == ids[:-1]
    assert bool((flat[1:][same] > flat[:-1][same]).all())
    # every (gaussian, tile) pair appears exactly once and lies inside the gaussian's tile box
    ty, tx = tile // tw, tile % tw
    m = means2d.reshape(-1, 2)[flat.long()]
    r = radii.flatten()[flat.long()].float()
    assert bool((tx >= torch.floor((m[:, 0] - r) / 16).clamp(0, tw)).all() and (tx < torch.ceil((m[:, 0] + r) / 16).clamp(0, tw)).all())
    assert bool((ty >= torch.floor((m[:, 1] - r) / 16).clamp(0, th)).all() and (ty < torch.ceil((m[:, 1] + r) / 16).clamp(0, th)).all())
    assert torch.unique(flat.long() * (tw * th) + tile).numel() == ids.numel()
    # offsets: exclusive cumsum of per-tile counts
    counts = torch.bincount(cam * tw * th + tile, minlength=C * tw * th)
    assert torch.equal(off.flatten().long(), torch.cumsum(counts, 0) - counts)


def test_reference_call_site_functions():
    # get_viewmat (model.py:22-38): inverse of the flipped c2w
    g = torch.Generator().manual_seed(0)
    A = torch.linalg.qr(torch.randn(3, 3, generator=g))[0]
    c2w = torch.eye(4)[None, :3].clone()
    c2w[0, :3, :3] = A
    c2w[0, :3, 3] = torch.tensor([1.0, 2.0, 3.0])
    vm = oracle.get_viewmat(c2w)
    flipped = torch.eye(4)
    flipped[:3, :3] = A * torch.tensor([[1.0, -1.0, -1.0]])
    flipped[:3, 3] = c2w[0, :3, 3]
    assert torch.allclose(vm[0] @ flipped, torch.eye(4), atol=1e-5)
    # depth loss (model.py:87-116): only finite & gt>0 pixels; empty mask -> 0 (model.py:111-114)
    d = torch.tensor([[1.0], [2.0], [float("inf")], [4.0]])
    gt = torch.tensor([[1.5], [0.0], [3.0], [float("nan")]])
    assert float(oracle.depth_l1_loss(d, gt, 0.2)) == pytest.approx(0.2 * 0.5)
    assert float(oracle.depth_l1_loss(d, torch.zeros_like(gt), 0.2)) == 0.0
    # composite + depth fill (model.py:295-306)
    render = torch.tensor([[[[0.2, 0.9, 0.5, 3.0], [0.0, 0.0, 0.0, 0.0]]]])
    alpha = torch.tensor([[[[0.5], [0.0]]]])
    rgb, depth = oracle.composite_and_fill(render, alpha, torch.tensor([1.0, 1.0, 0.0]))
    assert torch.allclose(rgb[0, 0, 0], torch.tensor([0.7, 1.0, 0.5])) and torch.allclose(rgb[0, 0, 1], torch.tensor([1.0, 1.0, 0.0]))
    assert float(depth[0, 0, 1, 0]) == 3.0  # filled with the detached max


@pytest.mark.parametrize("name", ["s0_small_rgbed", "s0_small_rgbd"])
def test_golden_fixture(name):
    """Regression pin of the oracle itself (generated by scripts/make_golden.py)."""
    torch.set_num_threads(1)
    z = np.load(os.path.join(GOLDEN, name + ".npz"))
    s = scene_s0(N=int(z["N"]), C=int(z["C"]), size=int(z["size"]), seed=int(z["seed"]))
    names = ("means", "quats", "scales", "opacities", "sh")
    leaves = {k: getattr(s, k).clone().requires_grad_(True) for k in names}
    render, alpha, info = oracle.rasterization(leaves["means"], leaves["quats"], leaves["scales"], leaves["opacities"], leaves["sh"],
                                               s.viewmats, s.Ks, s.width, s.height, sh_degree=3, render_mode=str(z["mode"]), absgrad=True)
    for k in ("radii", "tiles_per_gauss", "isect_ids", "flatten_ids", "isect_offsets"):
        assert np.array_equal(info[k].numpy(), z[k]), k
    assert np.allclose(render.detach().numpy(), z["render"], rtol=1e-5, atol=1e-6)
    assert np.allclose(alpha.detach().numpy(), z["alpha"], rtol=1e-5, atol=1e-6)
    bg = torch.from_numpy(z["bg"])
    total = 0.0
    for c in range(s.C):
        rgb, depth = oracle.composite_and_fill(render[c:c + 1], alpha[c:c + 1], bg)
        total = total + oracle.rgb_l1_loss(rgb, s.gt_rgb[c:c + 1]) + oracle.depth_l1_loss(depth, s.gt_depth[c:c + 1], 0.2)
    (total / s.C).backward()
    assert float(total / s.C) == pytest.approx(float(z["loss"]), rel=1e-5)
    for k in names:
        ref = z["grad_" + k]
        assert np.allclose(leaves[k].grad.numpy(), ref, rtol=1e-3, atol=1e-5 * float(np.abs(ref).mean() + 1e-12) * 100), k


def test_ssim_against_direct_window_sum():
    """Separable conv == direct 11x11 window sum; SSIM(x, x) == 1; range sanity."""
    g = torch.Generator().manual_seed(0)
    x = torch.rand(1, 20, 23, 3, generator=g, dtype=torch.float64)
    y = (x + 0.1 * torch.randn(x.shape, generator=g, dtype=torch.float64)).clamp(0, 1)
    assert float(oracle.ssim(x, x)) == pytest.approx(1.0, abs=1e-12)
    c = torch.arange(11, dtype=torch.float64) - 5
    w1 = torch.exp(-(c ** 2) / (2 * 1.5 ** 2))
    w1 = w1 / w1.sum()
    w2 = w1[:, None] * w1[None, :]
    C1, C2 = 0.01 ** 2, 0.03 ** 2
    vals = []
    for ch in range(3):
        for i in range(20 - 10):
            for j in range(23 - 10):
                px, py = x[0, i:i + 11, j:j + 11, ch], y[0, i:i + 11, j:j + 11, ch]
                m1, m2 = (w2 * px).sum(), (w2 * py).sum()
                s1, s2, s12 = (w2 * px * px).sum() - m1 * m1, (w2 * py * py).sum() - m2 * m2, (w2 * px * py).sum() - m1 * m2
                vals.append(((2 * m1 * m2 + C1) * (2 * s12 + C2)) / ((m1 * m1 + m2 * m2 + C1) * (s1 + s2 + C2)))
    direct = torch.stack(vals).mean()
    assert float(oracle.ssim(y, x)) == pytest.approx(float(direct), rel=1e-10)
    assert 0.0 < float(direct) < 1.0
