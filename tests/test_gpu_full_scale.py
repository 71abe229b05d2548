"""GPU parity AT THE SIZES THE BENCH RUNS (BASELINE.json configs[1] and a configs[3]-style forest), against the oracle:

* projection + SH over all Gaussians: bit-exact (radii, means2d, depths, conics, colours);
* gsplat's intersection lists over all ~6.5 M entries: isect_ids / flatten_ids / isect_offsets `torch.equal` to the
  oracle's stable 64-bit sort (not only to CUB);
* compositing forward AND backward on a sample of >= 64 tiles -- the densest tile, image-edge / corner tiles (1080 rows
  = 67.5 tiles: the last tile row is half empty) and random ones -- for both list kinds the product composites on:
  gsplat's bounding-box lists and the exact lists.  The oracle evaluates exactly those tiles (`_tile_forward` + the
  explicit A.6 backward through `rasterize_to_pixels(_bwd)` on a list restricted to the sample); the CUDA backward gets
  output gradients that are non-zero only on the sampled tiles, so its per-Gaussian gradients are the contributions of
  those tiles only.
"""
import pytest
import torch

import oracle
from helpers import assert_close_frac
from qed_splatter_b200 import ops
from qed_splatter_b200.scenes import scene_s1, scene_s3

pytestmark = pytest.mark.gpu
TILE = 16
_cache = {}


def _scene(name):
    if name == "s1":
        return scene_s1(N=1_000_000, width=1920, height=1080)
    return scene_s3(N=2_000_000, C=1, width=1440, height=1080)


def _setup(name, cuda):
    """Scene + oracle projection (CPU) + CUDA projection, shared by the tests of one scene."""
    if name in _cache:
        return _cache[name]
    _cache.clear()  # one full-size scene at a time
    s = _scene(name)
    radii_o, means2d_o, depths_o, conics_o, _ = oracle.fully_fused_projection(s.means, s.quats, s.scales, s.viewmats, s.Ks, s.width, s.height)
    dirs = s.means[None] - oracle.torch_impl.camera_positions(s.viewmats)[:, None, :]
    cols_o = torch.clamp_min(oracle.spherical_harmonics(3, dirs, s.sh[None], masks=radii_o > 0) + 0.5, 0.0)
    cols_o = torch.cat([cols_o, depths_o[..., None]], dim=-1)
    g = {k: getattr(s, k).to(cuda) for k in ("means", "quats", "scales", "opacities", "sh", "viewmats", "Ks")}
    radii, means2d, depths, conics, _, cols, opac, tiles, geom, tiles_exact = ops.project_gaussians(
        g["means"], g["quats"], g["scales"], g["opacities"], g["sh"], g["viewmats"], g["Ks"], s.width, s.height, sh_degree=3, n_color=3,
        append_depth=True)
    out = dict(s=s, o=dict(radii=radii_o, means2d=means2d_o, depths=depths_o, conics=conics_o, cols=cols_o, opac=s.opacities[None].contiguous()),
               g=dict(radii=radii, means2d=means2d, depths=depths, conics=conics, cols=cols, opac=opac, tiles=tiles, geom=geom, tiles_exact=tiles_exact))
    _cache[name] = out
    return out


@pytest.mark.timeout(900)
@pytest.mark.parametrize("name", ["s1", "s3"])
def test_projection_and_intersections_bit_exact_at_full_size(cuda, name):
    st = _setup(name, cuda)
    s, o, g = st["s"], st["o"], st["g"]
    vis = o["radii"] > 0
    assert torch.equal(g["radii"].cpu(), o["radii"])
    for k in ("means2d", "depths", "conics", "cols"):
        a, b = g[k].cpu(), o[k]
        m = vis if a.dim() == 2 else vis[..., None].expand_as(a)
        assert torch.equal(a[m], b[m]), f"{k} not bit-exact at full size"
    tw, th = ops.tile_grid(s.width, s.height, TILE)
    tiles_o, ids_o, flat_o = oracle.isect_tiles(o["means2d"], o["radii"], o["depths"], TILE, tw, th)
    off_o = oracle.isect_offset_encode(ids_o, 1, tw, th)
    tiles_g, ids_g, flat_g, off_g = ops.isect_tiles(g["means2d"], g["radii"], g["depths"], TILE, tw, th, tiles_per_gauss=g["tiles"],
                                                    return_offsets=True)
    assert ids_o.numel() > 3_000_000  # this IS the bench-scale list
    assert torch.equal(tiles_g.cpu(), tiles_o) and torch.equal(ids_g.cpu(), ids_o) and torch.equal(flat_g.cpu(), flat_o)
    assert torch.equal(off_g.cpu(), off_o)
    st["lists_o"] = (flat_o, off_o)


def _sample_tiles(starts, ends, tw, th, n_random=61, seed=11):
    lens = ends - starts
    nonempty = torch.nonzero(lens > 0).flatten()
    picks = {int(torch.argmax(lens))}  # the densest tile
    last_row = [t for t in range((th - 1) * tw, th * tw) if lens[t] > 0]  # half-empty tile row (1080 = 67.5 tiles)
    right_col = [t for t in range(tw - 1, tw * th, tw) if lens[t] > 0]
    for cand in (last_row[:1], last_row[-1:], right_col[:1], right_col[len(right_col) // 2:len(right_col) // 2 + 1]):
        picks.update(cand)
    gen = torch.Generator().manual_seed(seed)
    perm = nonempty[torch.randperm(nonempty.numel(), generator=gen)]
    for t in perm.tolist():
        if len(picks) >= n_random + 5:
            break
        picks.add(t)
    return sorted(picks)


@pytest.mark.timeout(1200)
@pytest.mark.parametrize("lists", ["exact", "bbox"])
@pytest.mark.parametrize("name", ["s1", "s3"])
def test_sampled_tiles_forward_backward_match_oracle_at_full_size(cuda, name, lists):
    st = _setup(name, cuda)
    s, o, g = st["s"], st["o"], st["g"]
    W, H = s.width, s.height
    tw, th = ops.tile_grid(W, H, TILE)
    n_tiles = tw * th
    if lists == "exact":
        flat, offsets, n_exact, _ = ops.isect_tiles_exact(g["means2d"], g["radii"], g["depths"], g["geom"], W, H, TILE, tw, th, g["tiles_exact"])
        assert flat.numel() == int(n_exact.item()) == int(g["tiles_exact"].sum())
        bounds = offsets.cpu().long()
        assert int(bounds[-1]) == int(n_exact.item())
    else:
        _, _, flat, off = ops.isect_tiles(g["means2d"], g["radii"], g["depths"], TILE, tw, th, tiles_per_gauss=g["tiles"], return_offsets=True)
        offsets = off
        bounds = torch.cat([off.flatten().cpu().long(), torch.tensor([flat.numel()])])
    starts, ends = bounds[:-1], bounds[1:]
    picks = _sample_tiles(starts, ends, tw, th)
    assert len(picks) >= 64
    flat_c = flat.cpu().long()

    # ---- CUDA: forward over the whole image, backward with gradients on the sampled tiles only ----
    leaf = {k: g[k].detach().clone().requires_grad_(True) for k in ("means2d", "conics", "cols", "opac")}
    render, alphas, last_ids = ops.rasterize_to_pixels(leaf["means2d"], leaf["conics"], leaf["cols"], leaf["opac"], W, H, TILE, offsets, flat,
                                                       absgrad=True, geom=g["geom"], normalize_last=False, return_last_ids=True)
    render_ed, _ = ops.rasterize_to_pixels(g["means2d"], g["conics"], g["cols"], g["opac"], W, H, TILE, offsets, flat, geom=g["geom"],
                                           normalize_last=True)
    gen = torch.Generator().manual_seed(3)
    sel = torch.zeros(1, H, W, 1)
    for t in picks:
        ty, tx = divmod(t, tw)
        sel[0, ty * TILE:(ty + 1) * TILE, tx * TILE:(tx + 1) * TILE] = 1.0
    v_render = torch.randn(1, H, W, 4, generator=gen) * sel
    v_alphas = torch.randn(1, H, W, 1, generator=gen) * sel
    torch.autograd.backward([render, alphas], [v_render.to(cuda), v_alphas.to(cuda)])

    # ---- oracle on exactly those tiles: the lists restricted to the sample ----
    keep = torch.zeros(n_tiles, dtype=torch.bool)
    keep[picks] = True
    lens_sub = torch.where(keep, ends - starts, torch.zeros_like(starts))
    off_sub = (torch.cumsum(lens_sub, 0) - lens_sub)
    flat_sub = torch.cat([flat_c[starts[t]:ends[t]] for t in picks]).to(torch.int32)
    off_sub_t = off_sub.reshape(1, th, tw).to(torch.int32)
    if lists == "exact":
        # every entry the exact list dropped from gsplat's list of a sampled tile is below 1/255 at every pixel centre
        # of that tile (float64), and the kept ones are an ordered subset
        if "lists_o" not in st:
            _, ids_b, flat_b = oracle.isect_tiles(o["means2d"], o["radii"], o["depths"], TILE, tw, th)
            st["lists_o"] = (flat_b, oracle.isect_offset_encode(ids_b, 1, tw, th))
        flat_b, off_b = st["lists_o"]
        bb = torch.cat([off_b.flatten().long(), torch.tensor([flat_b.numel()])])
        m2, cn, op = o["means2d"][0].double(), o["conics"][0].double(), o["opac"][0].double()
        for t in picks[:24]:
            full = flat_b[bb[t]:bb[t + 1]].long()
            kept = flat_c[starts[t]:ends[t]]
            pos = {int(v): i for i, v in enumerate(full.tolist())}
            idx = [pos[int(v)] for v in kept.tolist()]
            assert idx == sorted(idx) and len(set(idx)) == len(idx), "exact list is not an ordered subset of gsplat's list"
            dropped = full[~torch.isin(full, kept)]
            if dropped.numel():
                ty, tx = divmod(t, tw)
                _, _, px, py = oracle.torch_impl._tile_pixels(ty, tx, TILE, W, H, torch.float64)
                alpha = oracle.torch_impl._tile_forward(px, py, m2[dropped], cn[dropped], op[dropped])[4]  # alpha_raw
                assert float(alpha.max()) < 1.0 / 255.0, "an entry that can reach alpha >= 1/255 was dropped"
    render_o, alphas_o, last_o = oracle.rasterize_to_pixels(o["means2d"], o["conics"], o["cols"], o["opac"], W, H, TILE, off_sub_t, flat_sub)
    vm, vabs, vcn, vco, vop = oracle.rasterize_to_pixels_bwd(o["means2d"], o["conics"], o["cols"], o["opac"], W, H, TILE, off_sub_t, flat_sub,
                                                             v_render, v_alphas)
    m = sel.bool()
    r_g, a_g = render.detach().cpu(), alphas.detach().cpu()
    assert_close_frac(r_g[m.expand_as(r_g)], render_o[m.expand_as(render_o)], 1e-4, 1e-4, 2e-3, f"{name}/{lists} render on sampled tiles")
    assert_close_frac(a_g[m], alphas_o[m], 1e-4, 1e-4, 2e-3, f"{name}/{lists} alpha")
    ed_o = render_o[..., 3:4] / alphas_o.clamp(min=1e-10)
    assert_close_frac(render_ed.cpu()[..., 3:4][m], ed_o[m], 1e-4, 1e-4, 2e-3, f"{name}/{lists} expected depth")
    # last contributing index: oracle indices are relative to the restricted list
    shift = torch.zeros(1, H, W, dtype=torch.long)
    for t in picks:
        ty, tx = divmod(t, tw)
        shift[0, ty * TILE:(ty + 1) * TILE, tx * TILE:(tx + 1) * TILE] = int(starts[t] - off_sub[t])
    hit = (alphas_o[..., 0] > 0) & m[..., 0]
    same = (last_ids.cpu().long()[hit] == (last_o.long() + shift)[hit]).float().mean()
    assert float(same) > 1 - 2e-3, f"last_ids agree on {float(same):.5f} of the sampled pixels"
    for what, a, b in (("v_means2d", leaf["means2d"].grad, vm), ("absgrad", leaf["means2d"].absgrad, vabs), ("v_conics", leaf["conics"].grad, vcn),
                       ("v_colors", leaf["cols"].grad, vco), ("v_opacities", leaf["opac"].grad, vop)):
        scale = float(b.abs().mean() + 1e-20) * (b.numel() / max(int((b != 0).sum()), 1))  # mean over the touched rows
        assert_close_frac(a.cpu(), b, 1e-3, 1e-3 * scale, 5e-3 * (int((b != 0).sum()) / b.numel()) + 1e-6, f"{name}/{lists} {what}")
