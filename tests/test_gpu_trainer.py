"""-m gpu: fused step / trainer kernels against the oracle and the torch reference arithmetic."""
import os

import numpy as np
import pytest
import torch

import oracle
from qed_splatter_b200 import _lib, rasterization
from qed_splatter_b200.pipeline import FusedSplatStep
from qed_splatter_b200.scenes import scene_s0
from qed_splatter_b200.trainer import GROUPS, GaussianArena, SplatTrainer, TrainConfig, adam_step_torch
from helpers import assert_close_frac, scene_args

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(__file__), "golden")
NAMES = ("means", "quats", "scales", "opacities", "sh")


def _ref_loss(render, alpha, gt_rgb, gt_depth, bg):
    total = 0.0
    for c in range(render.shape[0]):
        rgb, depth = oracle.composite_and_fill(render[c:c + 1], alpha[c:c + 1], bg)
        total = total + oracle.rgb_l1_loss(rgb, gt_rgb[c:c + 1]) + oracle.depth_l1_loss(depth, gt_depth[c:c + 1], 0.2)
    return total / render.shape[0]


@pytest.mark.parametrize("name", ["s0_small_rgbed", "s0_small_rgbd"])
def test_cuda_matches_golden_fixture(cuda, name):
    z = np.load(os.path.join(GOLDEN, name + ".npz"))
    s = scene_s0(N=int(z["N"]), C=int(z["C"]), size=int(z["size"]), seed=int(z["seed"]))
    lg = {k: getattr(s, k).to(cuda).requires_grad_(True) for k in NAMES}
    render, alpha, info = rasterization(lg["means"], lg["quats"], lg["scales"], lg["opacities"], lg["sh"], s.viewmats.to(cuda), s.Ks.to(cuda),
                                        s.width, s.height, sh_degree=3, render_mode=str(z["mode"]), absgrad=True)
    for k in ("radii", "tiles_per_gauss", "isect_ids", "flatten_ids", "isect_offsets"):
        assert np.array_equal(info[k].cpu().numpy(), z[k]), k
    assert_close_frac(render, torch.from_numpy(z["render"]), 1e-4, 1e-4, 2e-3, "render")
    assert_close_frac(alpha, torch.from_numpy(z["alpha"]), 1e-4, 1e-4, 2e-3, "alpha")
    loss = _ref_loss(render, alpha, s.gt_rgb.to(cuda), s.gt_depth.to(cuda), torch.from_numpy(z["bg"]).to(cuda))
    loss.backward()
    assert float(loss) == pytest.approx(float(z["loss"]), rel=1e-4)
    for k in NAMES:
        ref = torch.from_numpy(z["grad_" + k])
        assert_close_frac(lg[k].grad, ref, 1e-3, 1e-3 * float(ref.abs().mean() + 1e-12), 5e-3, f"v_{k}")


@pytest.mark.parametrize("mode,C", [("RGB+ED", 1), ("RGB+D", 2)])
def test_fused_step_matches_oracle(cuda, mode, C):
    s = scene_s0(N=2500, C=C, size=80)
    lo = {k: getattr(s, k).clone().requires_grad_(True) for k in NAMES}
    bg = torch.tensor([0.3, 0.1, 0.6])
    ro, ao, io = oracle.rasterization(lo["means"], lo["quats"], lo["scales"], lo["opacities"], lo["sh"], s.viewmats, s.Ks, s.width, s.height,
                                      sh_degree=2, render_mode=mode, absgrad=True)
    loss_o = _ref_loss(ro, ao, s.gt_rgb, s.gt_depth, bg)
    io["means2d"].retain_grad()
    loss_o.backward()
    g = s.to(cuda)
    fs = FusedSplatStep(cuda)
    out = fs.step(g.means, g.quats, g.scales, g.opacities, g.sh, g.viewmats, g.Ks, g.width, g.height, 2, g.gt_rgb, g.gt_depth, bg.to(cuda),
                  render_mode=mode)
    assert float(out.loss[0]) == pytest.approx(float(loss_o), rel=1e-4)
    for k in NAMES:
        scale = float(lo[k].grad.abs().mean()) + 1e-12
        assert_close_frac(out.grads[k], lo[k].grad, 1e-3, 1e-3 * scale, 5e-3, f"v_{k}")
    # packed record: slots 0,1 = v_means2d of the loss
    pk = out.packed_grads.view(C, s.N, 12).cpu()
    scale = float(io["means2d"].grad.abs().mean()) + 1e-12
    assert_close_frac(pk[..., :2], io["means2d"].grad, 1e-3, 1e-3 * scale, 5e-3, "packed v_means2d")


@pytest.mark.parametrize("size", [80, 53])
def test_fused_step_with_ssim_matches_oracle(cuda, size):
    """Full splatfacto RGB loss (0.8 L1 + 0.2 (1 - SSIM)) + depth-L1, gradient through the fused kernels."""
    s = scene_s0(N=2500, C=2, size=size)
    s.scales = s.scales * 2.0
    lo = {k: getattr(s, k).clone().requires_grad_(True) for k in NAMES}
    bg = torch.tensor([0.3, 0.1, 0.6])
    ro, ao, io = oracle.rasterization(lo["means"], lo["quats"], lo["scales"], lo["opacities"], lo["sh"], s.viewmats, s.Ks, s.width, s.height,
                                      sh_degree=3, render_mode="RGB+D")
    total = 0.0
    for c in range(s.C):
        rgb, depth = oracle.composite_and_fill(ro[c:c + 1], ao[c:c + 1], bg)
        total = total + oracle.rgb_loss(rgb, s.gt_rgb[c:c + 1], 0.2) + oracle.depth_l1_loss(depth, s.gt_depth[c:c + 1], 0.2)
    loss_o = total / s.C
    loss_o.backward()
    g = s.to(cuda)
    fs = FusedSplatStep(cuda)
    out = fs.step(g.means, g.quats, g.scales, g.opacities, g.sh, g.viewmats, g.Ks, g.width, g.height, 3, g.gt_rgb, g.gt_depth, bg.to(cuda),
                  render_mode="RGB+D", rgb_weight=0.8, ssim_lambda=0.2)
    assert float(out.loss[0]) == pytest.approx(float(loss_o), rel=2e-4)
    for k in NAMES:
        scale = float(lo[k].grad.abs().mean()) + 1e-12
        assert_close_frac(out.grads[k], lo[k].grad, 1e-3, 1e-3 * scale, 5e-3, f"v_{k}")


@pytest.mark.parametrize("mode,C,ssim", [("RGB+ED", 1, 0.0), ("RGB+D", 2, 0.2)])
def test_depth_supervised_loss_is_the_reference_loss(cuda, mode, C, ssim):
    """`rasterization()` + `depth_supervised_loss()` (one fused launch group, autograd-visible) == the reference's
    torch formulation (model.py:295-306, 73-118) on the oracle's render: same number, same parameter gradients,
    and the gradient scales with whatever the caller multiplies the loss by."""
    from qed_splatter_b200 import depth_supervised_loss

    s = scene_s0(N=2500, C=C, size=72)
    s.scales = s.scales * 1.5
    bg = torch.tensor([0.3, 0.1, 0.6])
    lo = {k: getattr(s, k).clone().requires_grad_(True) for k in NAMES}
    ro, ao, _ = oracle.rasterization(lo["means"], lo["quats"], lo["scales"], lo["opacities"], lo["sh"], s.viewmats, s.Ks, s.width, s.height,
                                     sh_degree=3, render_mode=mode)
    total = 0.0
    for c in range(C):
        rgb, depth = oracle.composite_and_fill(ro[c:c + 1], ao[c:c + 1], bg)
        total = total + oracle.rgb_loss(rgb, s.gt_rgb[c:c + 1], ssim) + oracle.depth_l1_loss(depth, s.gt_depth[c:c + 1], 0.2)
    loss_o = 0.5 * total / C
    loss_o.backward()

    g = s.to(cuda)
    lg = {k: getattr(g, k).clone().requires_grad_(True) for k in NAMES}
    render, alpha, info = rasterization(lg["means"], lg["quats"], lg["scales"], lg["opacities"], lg["sh"], g.viewmats, g.Ks, g.width, g.height,
                                        sh_degree=3, render_mode=mode, absgrad=True)
    tot, l_rgb, l_depth = depth_supervised_loss(render, alpha, g.gt_rgb, g.gt_depth, bg.to(cuda), rgb_weight=1.0 - ssim, depth_lambda=0.2,
                                                ssim_lambda=ssim)
    (0.5 * tot).backward()
    assert float(0.5 * tot) == pytest.approx(float(loss_o), rel=2e-4)
    assert float(l_rgb + l_depth) == pytest.approx(float(tot), rel=1e-5)
    for k in NAMES:
        scale = float(lo[k].grad.abs().mean()) + 1e-12
        assert_close_frac(lg[k].grad, lo[k].grad, 1e-3, 1e-3 * scale, 5e-3, f"v_{k}")


@pytest.mark.parametrize("ssim", [0.0, 0.2])
def test_uint8_ground_truth_is_converted_in_kernel(cuda, ssim):
    """The data side caches images as uint8 (config.py:37); splatfacto converts `image.float() / 255.0` every step.
    Passing the uint8 tensor gives the same loss and bit-identical gradients without materialising the float image."""
    from qed_splatter_b200 import depth_supervised_loss

    s = scene_s0(N=2000, C=2, size=64).to(cuda)
    bg = torch.tensor([0.3, 0.1, 0.6], device=cuda)
    gt_u8 = (s.gt_rgb * 255.0).round().clamp(0, 255).to(torch.uint8)
    gt_f = gt_u8.float() / 255.0
    res = []
    for gt in (gt_f, gt_u8):
        lg = {k: getattr(s, k).clone().requires_grad_(True) for k in NAMES}
        render, alpha, _ = rasterization(lg["means"], lg["quats"], lg["scales"], lg["opacities"], lg["sh"], s.viewmats, s.Ks, s.width, s.height,
                                         sh_degree=3, render_mode="RGB+ED")
        render.retain_grad()
        tot = depth_supervised_loss(render, alpha, gt, s.gt_depth, bg, rgb_weight=1.0 - ssim, depth_lambda=0.2, ssim_lambda=ssim)[0]
        tot.backward()
        res.append((tot.detach().clone(), render.grad.clone()))
    # the loss sums are double atomics across blocks (order varies run to run): equal to float rounding; the per-pixel
    # gradients are bit-identical
    assert torch.allclose(res[0][0], res[1][0], rtol=1e-6, atol=0) and torch.equal(res[0][1], res[1][1])
    fs = FusedSplatStep(cuda)
    a = fs.step(s.means, s.quats, s.scales, s.opacities, s.sh, s.viewmats, s.Ks, s.width, s.height, 3, gt_f, s.gt_depth, bg, ssim_lambda=ssim,
                rgb_weight=1.0 - ssim).loss.clone()
    b = fs.step(s.means, s.quats, s.scales, s.opacities, s.sh, s.viewmats, s.Ks, s.width, s.height, 3, gt_u8, s.gt_depth, bg, ssim_lambda=ssim,
                rgb_weight=1.0 - ssim).loss.clone()
    assert torch.allclose(a, b, rtol=1e-6, atol=0)


@pytest.mark.timeout(120)
def test_pair_counters_of_instrumented_kernels(cuda):
    """bench.py's roofline uses the work counters of the instrumented (STATS) compositor kernels: they must run
    (a divergent warp vote once hung them), leave the results alone and be consistent with the oracle's count."""
    s = scene_s0(N=4000, C=2, size=120)
    bg = torch.tensor([0.3, 0.1, 0.6])
    g = s.to(cuda)
    fs = FusedSplatStep(cuda)
    out = fs.step(g.means, g.quats, g.scales, g.opacities, g.sh, g.viewmats, g.Ks, g.width, g.height, 3, g.gt_rgb, g.gt_depth, bg.to(cuda))
    before = out.packed_grads.clone()
    c = fs.count_pairs()
    assert torch.equal(out.packed_grads, before)
    for w in ("fwd", "bwd"):
        assert c[f"{w}_entries_loaded"] >= c[f"{w}_entries_staged"] > 0
        assert c[f"{w}_pairs_evaluated"] >= c[f"{w}_pairs_contributing"] > 0
    # forward and backward composite the same (pixel, Gaussian) pairs, up to threshold flips
    assert abs(c["fwd_pairs_contributing"] - c["bwd_pairs_contributing"]) <= 2e-3 * c["fwd_pairs_contributing"]
    # the oracle's count: pairs with alpha >= 1/255 in front of each pixel's last composited Gaussian
    ro, ao, io = oracle.rasterization(s.means, s.quats, s.scales, s.opacities, s.sh, s.viewmats, s.Ks, s.width, s.height, sh_degree=3,
                                      render_mode="RGB+ED")
    n_o = oracle.count_composited_pairs(io["means2d"], io["conics"], io["opacities"], s.width, s.height, 16, io["isect_offsets"], io["flatten_ids"])
    assert abs(c["fwd_pairs_contributing"] - n_o) <= 2e-3 * n_o, (c["fwd_pairs_contributing"], n_o)


@pytest.mark.parametrize("size,scale", [(120, 1.0), (100, 0.3)])
def test_exact_tile_lists(cuda, size, scale):
    """The fused step's exact tile lists: an order-preserving subset of gsplat's bounding-box lists, every dropped
    (Gaussian, tile) entry stays below alpha = 1/255 at every pixel centre of its tile, and so the image is
    bit-identical and the gradients differ by summation order only."""
    s = scene_s0(N=5000, C=2, size=size)
    s.scales = s.scales * scale
    g = s.to(cuda)
    bg = torch.tensor([0.3, 0.1, 0.6], device=cuda)
    res = {}
    for exact in (False, True):
        fs = FusedSplatStep(cuda, exact_tile_lists=exact)
        out = fs.step(g.means, g.quats, g.scales, g.opacities, g.sh, g.viewmats, g.Ks, g.width, g.height, 3, g.gt_rgb, g.gt_depth, bg)
        f = fs._fwd
        n = fs.n_isects_exact()
        n_tiles = s.C * f["tw"] * f["th"]
        off = f["offsets"][:n_tiles].tolist() + [n]
        res[exact] = dict(render=out.render.clone(), alphas=out.alphas.clone(), loss=out.loss.clone(), grads={k: v.clone() for k, v in out.grads.items()},
                          flat=f["flat"][:n].tolist(), off=off, n=n, M=out.n_isects, geom=f["geom"].clone(), tw=f["tw"], th=f["th"])
    a, b = res[False], res[True]
    assert a["n"] == a["M"] and b["n"] == b["M"] and 0 < b["n"] < a["n"]
    assert torch.equal(a["render"], b["render"]) and torch.equal(a["alphas"], b["alphas"]) and torch.equal(a["loss"], b["loss"])
    for k in a["grads"]:
        sc = float(a["grads"][k].abs().mean()) + 1e-12
        assert_close_frac(b["grads"][k], a["grads"][k], 1e-4, 1e-4 * sc, 1e-3, f"v_{k}")
    geom = a["geom"].view(-1, 8).cpu()
    tw, th = a["tw"], a["th"]
    dropped = 0
    for t in range(s.C * tw * th):
        full, kept = a["flat"][a["off"][t]:a["off"][t + 1]], b["flat"][b["off"][t]:b["off"][t + 1]]
        it = iter(full)
        assert all(any(x == y for y in it) for x in kept), f"tile {t}: not an ordered subset"
        gone = sorted(set(full) - set(kept))  # a Gaussian appears at most once per tile
        assert len(gone) == len(full) - len(kept)
        if not gone:
            continue
        dropped += len(gone)
        tile = t % (tw * th)
        ty, tx = divmod(tile, tw)
        ys = torch.arange(ty * 16, min(ty * 16 + 16, s.height)).float() + 0.5
        xs = torch.arange(tx * 16, min(tx * 16 + 16, s.width)).float() + 0.5
        gg = geom[gone].double()  # mx, my, opacity, depth, a, b, c, -
        dx = gg[:, 0, None, None] - xs[None, None, :].double()
        dy = gg[:, 1, None, None] - ys[None, :, None].double()
        sigma = 0.5 * (gg[:, 4, None, None] * dx * dx + gg[:, 6, None, None] * dy * dy) + gg[:, 5, None, None] * dx * dy
        alpha = gg[:, 2, None, None] * torch.exp(-sigma)
        assert float(alpha.max()) < 1.0 / 255.0, f"tile {t}: a dropped entry reaches alpha {float(alpha.max())}"
    assert dropped == a["n"] - b["n"]


def test_folded_activations_match_torch_chain_rule(cuda):
    """`activations` = log-scales + logit-opacities inside the projection kernels (model.py:269-271) == torch exp / sigmoid
    outside + their chain rule: same image bit for bit, same gradients."""
    s = scene_s0(N=3000, C=2, size=96).to(cuda)
    bg = torch.tensor([0.2, 0.2, 0.2], device=cuda)
    log_s, logit_o = torch.log(s.scales), torch.logit(s.opacities.clamp(1e-4, 1 - 1e-4))
    sc, op = torch.exp(log_s), torch.sigmoid(logit_o)
    fs = FusedSplatStep(cuda)
    ref = fs.step(s.means, s.quats, sc, op, s.sh, s.viewmats, s.Ks, s.width, s.height, 3, s.gt_rgb, s.gt_depth, bg)
    ref_render, ref_g = ref.render.clone(), {k: v.clone() for k, v in ref.grads.items()}
    out = fs.step(s.means, s.quats, log_s, logit_o, s.sh, s.viewmats, s.Ks, s.width, s.height, 3, s.gt_rgb, s.gt_depth, bg, activations=3)
    assert torch.equal(out.render, ref_render)
    want = dict(ref_g, scales=ref_g["scales"] * sc, opacities=ref_g["opacities"] * op * (1 - op))
    for k in NAMES:
        scale = float(want[k].abs().mean()) + 1e-12
        assert_close_frac(out.grads[k], want[k], 1e-4, 1e-5 * scale, 1e-3, f"v_{k}")


@pytest.mark.parametrize("step", [600, 3100, 4100])
def test_fused_densify_equals_torch_densify(cuda, step):
    """`qed_arena_gather` (one gather over the three arenas, decisions on index tensors) builds exactly the arenas the
    torch cat / index formulation of gsplat's duplicate / split / remove builds: same order, same bits, same moments.
    step 600: grow only; 3100: + prune by scale and screen size; 4100: screen-size rules off."""
    from qed_splatter_b200.trainer import GaussianArena, StrategyState, TrainConfig, refine_gaussians

    N = 40_000
    g = torch.Generator().manual_seed(step)
    mk = lambda *shape: torch.randn(*shape, generator=g)
    means, quats, sh = mk(N, 3), mk(N, 4), mk(N, 16, 3) * 0.1
    log_s = torch.rand(N, 3, generator=g) * 5.0 - 6.0          # exp: 0.0025 .. 0.37 (some above the 0.5 cull only after max)
    log_s[:200] += 3.0                                         # a few huge ones (pruned by scale once that rule is on)
    logit_o = mk(N) * 3.0                                      # some below sigmoid^-1(0.005)
    res = {}
    for impl in ("torch", "fused"):
        arena = GaussianArena(means.to(cuda), quats.to(cuda), log_s.to(cuda), logit_o.to(cuda), sh.to(cuda))
        gg = torch.Generator().manual_seed(7)
        arena.exp_avg.copy_(torch.randn(arena.exp_avg.numel(), generator=gg).to(cuda))
        arena.exp_avg_sq.copy_(torch.rand(arena.exp_avg_sq.numel(), generator=gg).to(cuda))
        st = StrategyState.zeros(N, cuda)
        st.count = torch.randint(0, 5, (N,), generator=gg).float().to(cuda)
        st.grad2d = (torch.rand(N, generator=gg) * 0.004).to(cuda)   # mean gradient straddles the 0.0005 threshold
        st.radii = (torch.rand(N, generator=gg) * 0.2).to(cuda)      # some above split (0.05) / cull (0.15) screen sizes
        info = refine_gaussians(arena, st, TrainConfig(), step, torch.Generator().manual_seed(99), impl=impl)
        res[impl] = (info, arena)
    (ia, a), (ib, b) = res["torch"], res["fused"]
    assert ia == ib and ia["n_dupli"] > 0 and ia["n_split"] > 0 and ia["n_prune"] > 0, (ia, ib)
    assert a.N == b.N and a.offsets == b.offsets and torch.equal(a.group_ends, b.group_ends)
    for name in ("param", "exp_avg", "exp_avg_sq", "grad"):
        assert torch.equal(getattr(a, name), getattr(b, name)), name


def test_view_sharding_equals_single_rank(cuda):
    """2 'ranks' x 1 view with grad_scale = 1/2, summed (what the all-reduce does) == 1 rank x 2 views."""
    s = scene_s0(N=3000, C=2, size=96).to(cuda)
    bg = torch.tensor([0.2, 0.2, 0.2], device=cuda)
    fs = FusedSplatStep(cuda)
    whole = fs.step(s.means, s.quats, s.scales, s.opacities, s.sh, s.viewmats, s.Ks, s.width, s.height, 3, s.gt_rgb, s.gt_depth, bg)
    whole_g = {k: v.clone() for k, v in whole.grads.items()}
    acc = {k: torch.zeros_like(v) for k, v in whole_g.items()}
    for r in range(2):
        sl = slice(r, r + 1)
        part = fs.step(s.means, s.quats, s.scales, s.opacities, s.sh, s.viewmats[sl].contiguous(), s.Ks[sl].contiguous(), s.width, s.height, 3,
                       s.gt_rgb[sl].contiguous(), s.gt_depth[sl].contiguous(), bg, grad_scale=0.5)
        for k in acc:
            acc[k] += part.grads[k]
    for k in acc:
        scale = float(whole_g[k].abs().mean()) + 1e-12
        assert_close_frac(acc[k], whole_g[k], 1e-4, 1e-5 * scale, 1e-3, f"sharded v_{k}")


def test_adam_kernel_and_strategy_kernel(cuda):
    cfg = TrainConfig()
    g = torch.Generator().manual_seed(0)
    N = 5000
    p = [torch.randn(N, 3, generator=g), torch.randn(N, 4, generator=g), torch.randn(N, 3, generator=g), torch.randn(N, generator=g),
         torch.randn(N, 16, 3, generator=g)]
    cpu = SplatTrainer(*p, cfg=cfg, backend="torch")
    gpu = SplatTrainer(*[t.to(cuda) for t in p], cfg=cfg, backend="cuda")
    for it in range(3):
        grad = torch.randn(cpu.arena.grad.shape, generator=g) * 1e-3
        cpu.arena.grad.copy_(grad)
        gpu.arena.grad.copy_(grad.to(cuda))
        cpu.optimizer_step()
        gpu.optimizer_step()
        cpu.step_count += 1
        gpu.step_count += 1
    # float32 rounding of lr/bias1 * m / (sqrt(v)/bias2 + eps) and of `param -= update` differs in the last bits
    # (measured: <= 1.5e-6 absolute on updates of ~0.09): compare updates to 1e-4 relative of their typical size
    for name, p0 in zip(GROUPS, p):
        du_gpu = gpu.arena.view(gpu.arena.param, name).cpu() - p0
        du_cpu = cpu.arena.view(cpu.arena.param, name) - p0
        assert float((du_gpu - du_cpu).abs().max()) <= 1e-4 * float(du_cpu.abs().mean()) + 5e-7, name
    assert torch.allclose(gpu.arena.exp_avg_sq.cpu(), cpu.arena.exp_avg_sq, rtol=1e-5, atol=1e-12)
    # strategy statistics
    C, W, H = 3, 640, 360
    packed = torch.rand(C * N, 12, generator=g)
    radii = torch.randint(0, 6, (C, N), generator=g, dtype=torch.int32)
    cpu.accumulate_stats(packed, radii, W, H, packed=True, n_cameras=6)
    gpu.accumulate_stats(packed.to(cuda), radii.to(cuda), W, H, packed=True, n_cameras=6)
    assert torch.allclose(gpu.state.grad2d.cpu(), cpu.state.grad2d, rtol=1e-5)
    assert torch.equal(gpu.state.count.cpu(), cpu.state.count) and torch.allclose(gpu.state.radii.cpu(), cpu.state.radii)


def test_trainer_steps_reduce_loss_and_refine(cuda):
    s = scene_s0(N=4000, C=2, size=96)
    cfg = TrainConfig(warmup_length=2, refine_every=3, reset_alpha_every=1000, pause_refine_after_reset=0, densify_grad_thresh=1e-5)
    tr = SplatTrainer(s.means.to(cuda), s.quats.to(cuda), torch.log(s.scales).to(cuda), torch.logit(s.opacities).to(cuda), s.sh.to(cuda),
                      cfg=cfg, backend="cuda")
    g = s.to(cuda)
    bg = torch.tensor([0.5, 0.5, 0.5], device=cuda)
    losses, infos = [], []
    for it in range(8):
        loss, info = tr.step(g.viewmats, g.Ks, g.width, g.height, g.gt_rgb, g.gt_depth, bg)
        losses.append(float(loss[0]))
        infos.append(info)
    assert all(np.isfinite(losses))
    assert min(losses[4:]) < losses[1], losses  # step 0 resets opacities (gsplat quirk), then the loss goes down
    fired = [i for i, x in enumerate(infos) if x is not None]
    assert fired == [3, 6] and infos[3]["n"] == tr.arena.N or infos[6]["n"] == tr.arena.N
    assert infos[3]["n_dupli"] + infos[3]["n_split"] > 0


@pytest.mark.parametrize("exact", [True, False])
def test_deferred_size_read_equals_synchronous_path_and_survives_overflow(cuda, exact):
    """FusedSplatStep(defer_sync=True): the intersection counts are read by the host only after the rest of the pass has
    been queued (buffers sized from earlier calls).  Same images as the synchronous path; a call whose lists exceed
    the capacity is repeated with larger buffers and still gives the right answer."""
    small = scene_s0(N=3000, C=1, size=96).to(cuda)
    big = scene_s0(N=3000, C=1, size=96, seed=42)
    big.scales = big.scales * 4.0  # ~16x the intersections of `small`: does not fit the capacity learnt from it
    big = big.to(cuda)
    bg = torch.tensor([0.1, 0.2, 0.3], device=cuda)

    def run(fs, s):
        out = fs.step(s.means, s.quats, s.scales, s.opacities, s.sh, s.viewmats, s.Ks, s.width, s.height, 3, s.gt_rgb, s.gt_depth, bg)
        return out.render.clone(), out.alphas.clone(), out.loss.clone(), {k: v.clone() for k, v in out.grads.items()}, out.n_isects

    sync = FusedSplatStep(cuda, defer_sync=False, exact_tile_lists=exact)
    defer = FusedSplatStep(cuda, defer_sync=True, exact_tile_lists=exact)
    ref_small, ref_big = run(sync, small), run(sync, big)
    first = run(defer, small)       # first call: capacity unknown -> synchronous
    second = run(defer, small)      # deferred
    assert defer._cap_isects > 0 and defer.overflow_repeats == 0
    third = run(defer, big)         # deferred, overflows, repeated
    assert defer.overflow_repeats == 1 and ref_big[4] > ref_small[4] + ref_small[4] // 4 + 4096  # did not fit the capacity learnt from `small`
    fourth = run(defer, big)        # deferred, fits now
    assert defer.overflow_repeats == 1
    for got, ref in ((first, ref_small), (second, ref_small), (third, ref_big), (fourth, ref_big)):
        assert torch.equal(got[0], ref[0]) and torch.equal(got[1], ref[1]) and got[4] == ref[4]
        assert float(got[2][0]) == pytest.approx(float(ref[2][0]), rel=1e-6)
        for k in ref[3]:
            scale = float(ref[3][k].abs().mean()) + 1e-20
            assert_close_frac(got[3][k], ref[3][k], 1e-4, 1e-5 * scale, 1e-3, f"deferred v_{k}")
    # forward-only use resolves by itself
    r1, a1 = defer.forward(small.means, small.quats, small.scales, small.opacities, small.sh, small.viewmats, small.Ks, 96, 96, 3)
    assert torch.equal(r1, ref_small[0]) and torch.equal(a1, ref_small[1])


@pytest.mark.parametrize("ssim", [0.0, 0.2])
def test_uint16_depth_is_scaled_in_kernel(cuda, ssim):
    """The raw uint16 depth image + the dataparser's unit scale (qed_splatter/dataparser.py:15) consumed by the loss kernels
    == the float32 image nerfstudio's loader would have produced (float64 product, rounded): bit-identical loss and
    gradients, through both the autograd drop-in and the fused step."""
    from qed_splatter_b200 import depth_supervised_loss, rasterization

    s = scene_s0(N=3000, C=2, size=64)
    g = torch.Generator().manual_seed(9)
    raw = torch.randint(0, 6000, (2, 64, 64, 1), generator=g, dtype=torch.int32)
    raw[torch.rand(2, 64, 64, 1, generator=g) < 0.15] = 0  # holes
    scale = 0.001 * 1.37  # unit scale x scene scale
    as_float = (raw.double() * scale).float().to(cuda)
    raw16 = raw.to(torch.uint16).to(cuda)
    a = scene_args(s, cuda)
    bg = torch.tensor([0.2, 0.4, 0.6], device=cuda)
    render, alpha, _ = rasterization(**a, width=64, height=64, sh_degree=3, render_mode="RGB+D")
    render, alpha = render.detach().requires_grad_(True), alpha.detach().requires_grad_(True)
    outs = []
    for d in (as_float, raw16, raw16.view(torch.int16)):
        render.grad = alpha.grad = None
        total, l_rgb, l_d = depth_supervised_loss(render, alpha, s.gt_rgb.to(cuda), d, bg, ssim_lambda=ssim, depth_unit_scale=scale)
        total.backward()
        outs.append((total.detach().clone(), l_d.clone(), render.grad.clone(), alpha.grad.clone()))
    for o in outs[1:]:
        assert all(torch.equal(x, y) for x, y in zip(o, outs[0]))
    fs = FusedSplatStep(cuda)
    ref = fs.step(a["means"], a["quats"], a["scales"], a["opacities"], a["colors"], a["viewmats"], a["Ks"], 64, 64, 3, s.gt_rgb.to(cuda), as_float, bg,
                  render_mode="RGB+D", ssim_lambda=ssim)
    ref_loss = ref.loss.clone()
    got = fs.step(a["means"], a["quats"], a["scales"], a["opacities"], a["colors"], a["viewmats"], a["Ks"], 64, 64, 3, s.gt_rgb.to(cuda), raw16, bg,
                  render_mode="RGB+D", ssim_lambda=ssim, depth_unit_scale=scale)
    assert torch.equal(got.loss, ref_loss)
