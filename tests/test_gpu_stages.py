"""-m gpu: per-stage parity of the CUDA path (through the C-ABI) against the CPU oracle.

Tolerances (BASELINE.json north_star): forward rel 1e-4, gradients rel 1e-3; integer stages bit-exact.
"""
import math

import pytest
import torch

import oracle
from qed_splatter_b200 import ops
from qed_splatter_b200.scenes import scene_s0
from helpers import assert_close_frac

pytestmark = pytest.mark.gpu


def _project_oracle(s, deg=3, comp=False):
    radii, means2d, depths, conics, comps = oracle.fully_fused_projection(
        s.means, s.quats, s.scales, s.viewmats, s.Ks, s.width, s.height, calc_compensations=comp)
    campos = oracle.torch_impl.camera_positions(s.viewmats)
    dirs = s.means[None] - campos[:, None]
    cols = oracle.spherical_harmonics(deg, dirs, s.sh[None].expand(s.C, -1, -1, -1), masks=radii > 0)
    cols = torch.clamp_min(cols + 0.5, 0.0)
    return radii, means2d, depths, conics, comps, cols


@pytest.mark.parametrize("deg", [0, 1, 2, 3])
def test_projection_bit_exact(cuda, deg):
    s = scene_s0(N=5000, C=3, size=128)
    s.means = s.means * 2.5  # spread the scene so a good share of the Gaussians is culled (behind / off-screen)
    r_o, m_o, d_o, c_o, _, col_o = _project_oracle(s, deg)
    g = s.to(cuda)
    radii, means2d, depths, conics, comps, cols, opac, tiles, geom, tiles_exact = ops.project_gaussians(
        g.means, g.quats, g.scales, g.opacities, g.sh, g.viewmats, g.Ks, g.width, g.height, sh_degree=deg)
    assert torch.equal(radii.cpu(), r_o), "radii must be bit-exact"
    vis = r_o > 0
    assert vis.sum() > 100
    assert torch.equal(means2d.cpu()[vis], m_o[vis]), "means2d bit-exact (pinned op order)"
    assert torch.equal(depths.cpu()[vis], d_o[vis])
    assert torch.equal(conics.cpu()[vis], c_o[vis])
    assert torch.equal(cols.cpu()[..., :3][vis], col_o[vis]), "SH colours bit-exact (pinned op order)"
    assert torch.equal(cols.cpu()[..., 3][vis], d_o[vis])
    assert torch.equal(opac.cpu()[vis], s.opacities[None].expand(s.C, -1)[vis])
    # culled entries are zero
    if bool((~vis).any()):
        assert float(means2d.cpu()[~vis].abs().max()) == 0.0 and float(cols.cpu()[~vis].abs().max()) == 0.0
    # tile counts + geom record
    tw, th = ops.tile_grid(s.width, s.height, 16)
    t_o, _, _ = oracle.isect_tiles(m_o, r_o, d_o, 16, tw, th, sort=False)
    assert torch.equal(tiles.cpu(), t_o)
    gg = geom.cpu()
    assert torch.equal(gg[..., 0:2][vis], m_o[vis]) and torch.equal(gg[..., 4:7][vis], c_o[vis])


def test_projection_antialiased_and_passthrough(cuda):
    s = scene_s0(N=3000, C=2, size=96)
    r_o, m_o, d_o, c_o, comp_o, _ = _project_oracle(s, 0, comp=True)
    g = s.to(cuda)
    rgb = torch.rand(s.N, 3, generator=torch.Generator().manual_seed(1))
    radii, means2d, depths, conics, comps, cols, opac, tiles, geom, tiles_exact = ops.project_gaussians(
        g.means, g.quats, g.scales, g.opacities, rgb.to(cuda), g.viewmats, g.Ks, g.width, g.height,
        calc_compensations=True, sh_degree=None)
    vis = r_o > 0
    assert torch.equal(radii.cpu(), r_o)
    assert torch.equal(comps.cpu()[vis], comp_o[vis])
    assert torch.equal(opac.cpu()[vis], (s.opacities[None] * comp_o)[vis])
    assert torch.equal(cols.cpu()[..., :3][vis], rgb[None].expand(s.C, -1, -1)[vis])


@pytest.mark.parametrize("deg,comp", [(3, False), (2, True), (0, False)])
def test_projection_backward(cuda, deg, comp):
    s = scene_s0(N=4000, C=3, size=128)
    gen = torch.Generator().manual_seed(3)
    leaves = [t.clone().requires_grad_(True) for t in (s.means, s.quats, s.scales, s.opacities, s.sh)]
    means, quats, scales, opacs, sh = leaves
    radii, m2, dep, con, cmp_, = oracle.fully_fused_projection(means, quats, scales, s.viewmats, s.Ks, s.width, s.height,
                                                              calc_compensations=comp)
    campos = oracle.torch_impl.camera_positions(s.viewmats)
    cols = oracle.spherical_harmonics(deg, means[None] - campos[:, None], sh[None].expand(s.C, -1, -1, -1), masks=radii > 0)
    cols = torch.clamp_min(cols + 0.5, 0.0)
    op = opacs[None].expand(s.C, -1)
    if comp:
        op = op * cmp_
    w = [torch.randn(t.shape, generator=gen) for t in (m2, dep, con, cols, op)]
    vis = (radii > 0)
    loss = sum((a * b * vis.reshape(vis.shape + (1,) * (a.dim() - 2))).sum() for a, b in zip((m2, dep, con, cols, op), w))
    loss.backward()

    g = s.to(cuda)
    gl = [t.detach().to(cuda).requires_grad_(True) for t in (s.means, s.quats, s.scales, s.opacities, s.sh)]
    out = ops.project_gaussians(gl[0], gl[1], gl[2], gl[3], gl[4], g.viewmats, g.Ks, g.width, g.height,
                                calc_compensations=comp, sh_degree=deg)
    radii_g, m2g, depg, cong, _, colg, opg, _, _, _ = out
    assert torch.equal(radii_g.cpu(), radii)
    visg = vis.to(cuda)
    wg = [t.to(cuda) for t in w]
    # the CUDA colour tensor has the depth channel appended: weight it with zero
    lossg = (m2g * wg[0] * visg[..., None]).sum() + (depg * wg[1] * visg).sum() + (cong * wg[2] * visg[..., None]).sum() \
        + (colg[..., :3] * wg[3] * visg[..., None]).sum() + (opg * wg[4] * visg).sum()
    lossg.backward()
    names = ["means", "quats", "scales", "opacities", "sh"]
    for name, a, b in zip(names, gl, leaves):
        scale = float(b.grad.abs().mean()) + 1e-12
        assert_close_frac(a.grad, b.grad, rtol=1e-3, atol=1e-3 * scale, max_frac=2e-3, what=f"v_{name}")


def test_isect_bit_exact(cuda):
    s = scene_s0(N=6000, C=3, size=200)  # 200 is not a multiple of 16
    r_o, m_o, d_o, _, _, _ = _project_oracle(s, 0)
    tw, th = ops.tile_grid(s.width, s.height, 16)
    t_o, ids_o, flat_o = oracle.isect_tiles(m_o, r_o, d_o, 16, tw, th)
    off_o = oracle.isect_offset_encode(ids_o, s.C, tw, th)
    for impl in ops.SORT_IMPLS:
        ops.set_sort_impl(impl)
        t, ids, flat = ops.isect_tiles(m_o.to(cuda), r_o.to(cuda), d_o.to(cuda), 16, tw, th)
        off = ops.isect_offset_encode(ids, s.C, tw, th)
        assert torch.equal(t.cpu(), t_o)
        assert ids.numel() == ids_o.numel() > 1000
        assert torch.equal(ids.cpu(), ids_o), f"isect_ids ({impl})"
        assert torch.equal(flat.cpu(), flat_o), f"flatten_ids ({impl})"
        assert torch.equal(off.cpu(), off_o), f"isect_offsets ({impl})"
    ops.set_sort_impl("two_level")


def test_isect_ties_and_empty(cuda):
    # equal depths -> sort ties must keep emission order (ascending flat index); also zero intersections
    C, N, W, H = 2, 500, 64, 48
    g = torch.Generator().manual_seed(5)
    means2d = torch.rand(C, N, 2, generator=g) * torch.tensor([W, H])
    radii = torch.randint(0, 12, (C, N), generator=g, dtype=torch.int32)
    depths = torch.full((C, N), 2.5)
    depths[:, ::7] = 1.25
    tw, th = ops.tile_grid(W, H, 16)
    t_o, ids_o, flat_o = oracle.isect_tiles(means2d, radii, depths, 16, tw, th)
    off_o = oracle.isect_offset_encode(ids_o, C, tw, th)
    zero = torch.zeros(C, N, dtype=torch.int32)
    for impl in ops.SORT_IMPLS:
        t, ids, flat = ops.isect_tiles(means2d.to(cuda), radii.to(cuda), depths.to(cuda), 16, tw, th, impl=impl)
        off = ops.isect_offset_encode(ids, C, tw, th)
        assert torch.equal(ids.cpu(), ids_o) and torch.equal(flat.cpu(), flat_o) and torch.equal(off.cpu(), off_o), impl
        t, ids, flat = ops.isect_tiles(means2d.to(cuda), zero.to(cuda), depths.to(cuda), 16, tw, th, impl=impl)
        off = ops.isect_offset_encode(ids, C, tw, th)
        assert ids.numel() == 0 and flat.numel() == 0 and int(off.abs().sum()) == 0 and int(t.sum()) == 0, impl


@pytest.mark.parametrize("n", [1, 31, 4096, 4097, 100_003])
def test_sort_pairs_matches_stable_sort(cuda, n):
    g = torch.Generator().manual_seed(n)
    keys = torch.randint(0, 1 << 40, (n,), generator=g, dtype=torch.int64)
    keys[::3] = keys[0]  # many ties
    vals = torch.arange(n, dtype=torch.int32)
    ks, order = torch.sort(keys, stable=True)
    for impl in ("own", "cub"):
        ko, vo = ops.sort_pairs(keys.to(cuda), vals.to(cuda), 40, impl=impl)
        assert torch.equal(ko.cpu(), ks) and torch.equal(vo.cpu(), vals[order]), impl


def _raster_inputs(s, deg=3, mode_depth=True):
    r_o, m_o, d_o, c_o, _, col_o = _project_oracle(s, deg)
    cols = torch.cat([col_o, d_o[..., None]], -1) if mode_depth else col_o
    tw, th = ops.tile_grid(s.width, s.height, 16)
    _, ids, flat = oracle.isect_tiles(m_o, r_o, d_o, 16, tw, th)
    off = oracle.isect_offset_encode(ids, s.C, tw, th)
    op = s.opacities[None].expand(s.C, -1).contiguous()
    return m_o, c_o, cols.contiguous(), op, off, flat


@pytest.mark.parametrize("size,D,bg", [(96, 4, False), (100, 3, True), (64, 1, False)])
def test_raster_forward_backward(cuda, size, D, bg):
    s = scene_s0(N=3000, C=2, size=size)
    m, c, cols, op, off, flat = _raster_inputs(s)
    cols = {4: cols, 3: cols[..., :3].contiguous(), 1: cols[..., 3:].contiguous()}[D]
    gen = torch.Generator().manual_seed(11)
    bgs = torch.rand(s.C, D, generator=gen) if bg else None
    leaves = [t.clone().requires_grad_(True) for t in (m, c, cols, op)]
    render_o, alpha_o, last_o = oracle.rasterize_to_pixels(*leaves, s.width, s.height, 16, off, flat, backgrounds=bgs)
    vr = torch.randn(render_o.shape, generator=gen)
    va = torch.randn(alpha_o.shape, generator=gen)
    ((render_o * vr).sum() + (alpha_o * va).sum()).backward()
    exp = oracle.rasterize_to_pixels_bwd(m, c, cols, op, s.width, s.height, 16, off, flat, vr, va[..., 0], backgrounds=bgs)

    gl = [t.detach().to(cuda).requires_grad_(True) for t in (m, c, cols, op)]
    render, alpha, last = ops.rasterize_to_pixels(*gl, s.width, s.height, 16, off.to(cuda), flat.to(cuda),
                                                  backgrounds=bgs.to(cuda) if bg else None, absgrad=True, return_last_ids=True)
    assert_close_frac(render, render_o, 1e-4, 1e-4, 2e-3, "render")
    assert_close_frac(alpha, alpha_o, 1e-4, 1e-4, 2e-3, "alpha")
    assert float((last.cpu() != last_o).float().mean()) < 2e-3, "last_ids"
    ((render * vr.to(cuda)).sum() + (alpha * va.to(cuda)).sum()).backward()
    for name, a, b in zip(["means2d", "conics", "colors", "opacities"], gl, leaves):
        scale = float(b.grad.abs().mean()) + 1e-12
        assert_close_frac(a.grad, b.grad, 1e-3, 1e-3 * scale, 5e-3, f"v_{name}")
    scale = float(exp[1].abs().mean()) + 1e-12
    assert_close_frac(gl[0].absgrad, exp[1], 1e-3, 1e-3 * scale, 5e-3, "absgrad")


def test_raster_cull_is_exact(cuda):
    """The warp-level culling must not change a single bit of the forward, and only atomics order in the backward."""
    s = scene_s0(N=20000, C=1, size=256)
    m, c, cols, op, off, flat = [t.to(cuda) for t in _raster_inputs(s)]
    outs = {}
    for cull in (True, False):
        ops.set_raster_cull(cull)
        gl = [t.clone().requires_grad_(True) for t in (m, c, cols, op)]
        render, alpha, last = ops.rasterize_to_pixels(*gl, s.width, s.height, 16, off, flat, absgrad=True, return_last_ids=True)
        (render.sum() + alpha.sum()).backward()
        outs[cull] = (render.detach(), alpha.detach(), last, [t.grad for t in gl])
    ops.set_raster_cull(True)
    assert torch.equal(outs[True][0], outs[False][0]) and torch.equal(outs[True][1], outs[False][1])
    assert torch.equal(outs[True][2], outs[False][2])
    for a, b in zip(outs[True][3], outs[False][3]):
        assert_close_frac(a, b, 1e-4, 1e-6 * float(b.abs().mean() + 1e-12), 1e-3, "grad cull vs no-cull")


@pytest.mark.parametrize("D,bg", [(4, True), (3, False), (1, False)])
def test_raster_backward_packed_matches_scalar(cuda, D, bg):
    """The default backward (two-wide fp32 + warp streams + shared-memory transpose) and the scalar kernel it
    replaced are the same function: identical up to summation order, on a scene with image-edge tiles."""
    s = scene_s0(N=20000, C=2, size=200)  # 200 = 12.5 tiles: clipped blocks on the right / bottom edge
    m, c, cols, op, off, flat = [t.to(cuda) for t in _raster_inputs(s)]
    cols = {4: cols, 3: cols[..., :3].contiguous(), 1: cols[..., 3:].contiguous()}[D]
    gen = torch.Generator().manual_seed(5)
    bgs = torch.rand(s.C, D, generator=gen).to(cuda) if bg else None
    vr = torch.randn(s.C, s.height, s.width, D, generator=gen).to(cuda)
    va = torch.randn(s.C, s.height, s.width, 1, generator=gen).to(cuda)
    grads = {}
    try:
        for packed in (True, False):
            ops.set_raster_packed(packed)
            gl = [t.clone().requires_grad_(True) for t in (m, c, cols, op)]
            render, alpha = ops.rasterize_to_pixels(*gl, s.width, s.height, 16, off, flat, backgrounds=bgs, absgrad=True)
            ((render * vr).sum() + (alpha * va).sum()).backward()
            grads[packed] = [t.grad for t in gl] + [gl[0].absgrad]
    finally:
        ops.set_raster_packed(True)
    for name, a, b in zip(["means2d", "conics", "colors", "opacities", "absgrad"], grads[True], grads[False]):
        assert_close_frac(a, b, 1e-4, 1e-5 * float(b.abs().mean() + 1e-12), 1e-3, f"packed vs scalar v_{name}")


def test_isect_full_size_paths_agree(cuda):
    """BASELINE configs[1] size (1M Gaussians, 1080p): the two-level build, the own 64-bit radix sort and the
    CUB baseline must produce the same bytes; keys must be sorted and ranges consistent."""
    from qed_splatter_b200.scenes import scene_s1

    s = scene_s1(N=1_000_000, targets=False).to(cuda)
    radii, means2d, depths, conics, comps, cols, opac, tiles, geom, tiles_exact = ops.project_gaussians(
        s.means, s.quats, s.scales, s.opacities, s.sh, s.viewmats, s.Ks, s.width, s.height, sh_degree=3)
    tw, th = ops.tile_grid(s.width, s.height, 16)
    res = {impl: ops.isect_tiles(means2d, radii, depths, 16, tw, th, tiles_per_gauss=tiles, impl=impl) for impl in ops.SORT_IMPLS}
    ids, flat = res["cub"][1], res["cub"][2]
    assert ids.numel() == int(tiles.sum()) > 1_000_000
    for impl in ("two_level", "own"):
        assert torch.equal(res[impl][1], ids), impl
        assert torch.equal(res[impl][2], flat), impl
    assert bool((ids[1:] >= ids[:-1]).all()), "sorted"
    off = ops.isect_offset_encode(ids, 1, tw, th).flatten().long()
    assert bool((off[1:] >= off[:-1]).all()) and int(off[0]) == 0
    # every entry's tile id matches the range it sits in
    tile_of_entry = (ids >> 32) & ((1 << (tw * th).bit_length()) - 1)
    counts = torch.bincount(tile_of_entry, minlength=tw * th)
    ends = torch.cat([off[1:], torch.tensor([ids.numel()], device=cuda)])
    assert torch.equal(ends - off, counts)


@pytest.mark.parametrize("N", [5003, 129])
def test_single_view_specialisations_equal_generic_kernels(cuda, N):
    """The reference renders one camera per step (model.py:211): the projection kernels have single-view instantiations
    (view loops collapsed at compile time, forward at 4 resident blocks, backward writing the SH gradient over the staged
    row).  They must be the generic kernels' results: forward bit for bit against a 3-view launch's slice, backward
    against the generic kernel (hook) to float reassociation."""
    from qed_splatter_b200 import _lib

    lib = _lib.load()
    s = scene_s0(N=N, C=3, size=112)  # N not a multiple of the 128 / 256-thread blocks
    s.means = s.means * 2.0
    g = s.to(cuda)
    full = ops.project_gaussians(g.means, g.quats, g.scales, g.opacities, g.sh, g.viewmats, g.Ks, g.width, g.height, sh_degree=3)
    for c in range(s.C):
        one = ops.project_gaussians(g.means, g.quats, g.scales, g.opacities, g.sh, g.viewmats[c:c + 1].contiguous(), g.Ks[c:c + 1].contiguous(),
                                    g.width, g.height, sh_degree=3)
        for a, b, name in zip(one, full, ("radii", "means2d", "depths", "conics", "comps", "colors", "opac", "tiles", "geom", "tiles_exact")):
            if b is not None and b.numel():
                assert torch.equal(a[0], b[c]), f"view {c}: {name}"
    # backward: specialised vs generic instantiation on the same single view
    grads = {}
    for mode in (5, 0, 6):
        old_b = lib.qed_debug_set_project_bwd_one(mode)
        old_f = lib.qed_debug_set_project_fwd_one(4 if mode else 0)
        try:
            leaves = [t.clone().requires_grad_(True) for t in (g.means, g.quats, g.scales, g.opacities, g.sh)]
            out = ops.project_gaussians(*leaves, g.viewmats[1:2].contiguous(), g.Ks[1:2].contiguous(), g.width, g.height, sh_degree=3)
            gen = torch.Generator(device="cpu").manual_seed(3)
            loss = sum((o * torch.randn(o.shape, generator=gen).to(cuda)).sum() for o in (out[1], out[2], out[3], out[5], out[6]))
            loss.backward()
            grads[mode] = [t.grad.clone() for t in leaves]
        finally:
            lib.qed_debug_set_project_bwd_one(old_b)
            lib.qed_debug_set_project_fwd_one(old_f)
    for mode in (5, 6):
        for a, b, name in zip(grads[mode], grads[0], ("means", "quats", "scales", "opacities", "sh")):
            torch.testing.assert_close(a, b, rtol=2e-5, atol=1e-6 * float(b.abs().max()), msg=lambda m: f"mode {mode} v_{name}: {m}")
    assert torch.equal(grads[5][4], grads[0][4]), "SH coefficient gradient of one view is a product, not a sum: identical"
