"""CPU: host-side trainer logic (arena layout, Adam arithmetic, densification schedule / surgery) and the
world_size-2 gloo path of the view-sharded step (gradient + statistic all-reduce, synchronized refine)."""
import math
import os

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from qed_splatter_b200.trainer import (FLOATS_PER_GAUSSIAN, GROUPS, GaussianArena, SplatTrainer, StrategyState, TrainConfig,
                                       adam_step_torch, refine_gaussians, reset_opacities)


def _params(N, seed=0):
    g = torch.Generator().manual_seed(seed)
    means = torch.randn(N, 3, generator=g)
    quats = torch.randn(N, 4, generator=g)
    log_scales = torch.log(torch.rand(N, 3, generator=g) * 0.05 + 0.001)
    logit_op = torch.logit(torch.rand(N, generator=g) * 0.9 + 0.05)
    sh = torch.randn(N, 16, 3, generator=g)
    return means, quats, log_scales, logit_op, sh


def test_arena_layout_and_views():
    p = _params(7)
    a = GaussianArena(*p)
    assert FLOATS_PER_GAUSSIAN == 59 and 59 * 7 <= a.param.numel() <= 59 * 7 + 3 * 5
    for name, t in zip(GROUPS, p):
        assert torch.equal(a.view(a.param, name), t)
        assert a.view(a.grad, name).data_ptr() == a.grad.data_ptr() + 4 * a.offsets[name][0]
        assert a.offsets[name][0] % 4 == 0  # 16-byte aligned group starts
    assert a.group_ends.tolist() == [24, 52, 76, 84, 420] and a.param.numel() == 420


def test_adam_matches_torch_optim():
    cfg = TrainConfig()
    p = _params(50, 1)
    a = GaussianArena(*p)
    ref = [t.clone().requires_grad_(True) for t in p]
    dc, rest = ref[4][:, :1].detach().clone().requires_grad_(True), ref[4][:, 1:].detach().clone().requires_grad_(True)
    opt = torch.optim.Adam([
        {"params": [ref[0]], "lr": cfg.lr_means}, {"params": [ref[1]], "lr": cfg.lr_quats}, {"params": [ref[2]], "lr": cfg.lr_scales},
        {"params": [ref[3]], "lr": cfg.lr_opacities}, {"params": [dc], "lr": cfg.lr_features_dc}, {"params": [rest], "lr": cfg.lr_features_rest}],
        eps=cfg.adam_eps)
    g = torch.Generator().manual_seed(2)
    lrs = {"means": cfg.lr_means, "quats": cfg.lr_quats, "scales": cfg.lr_scales, "opacities": cfg.lr_opacities, "sh": cfg.lr_features_dc}
    for t in range(1, 4):
        grads = [torch.randn(x.shape, generator=g) * 1e-3 for x in p]
        for name, gr in zip(GROUPS, grads):
            a.view(a.grad, name).copy_(gr)
        for x, gr in zip(ref[:4], grads[:4]):
            x.grad = gr.clone()
        dc.grad, rest.grad = grads[4][:, :1].clone(), grads[4][:, 1:].clone()
        opt.step()
        adam_step_torch(a, lrs, cfg.lr_features_rest, cfg, t)
    for name, x in zip(GROUPS[:4], ref[:4]):
        assert torch.allclose(a.view(a.param, name), x.detach(), rtol=1e-5, atol=1e-7), name
    assert torch.allclose(a.view(a.param, "sh")[:, :1], dc.detach(), rtol=1e-5, atol=1e-7)
    assert torch.allclose(a.view(a.param, "sh")[:, 1:], rest.detach(), rtol=1e-5, atol=1e-7)


def test_lr_schedule_endpoints():
    cfg = TrainConfig()
    assert cfg.lr_means_at(0) == pytest.approx(1.6e-4) and cfg.lr_means_at(30000) == pytest.approx(1.6e-6)
    assert cfg.lr_means_at(15000) == pytest.approx(1.6e-5)


def test_refine_duplicate_split_prune():
    cfg = TrainConfig()
    N = 40
    means, quats, log_scales, logit_op, sh = _params(N, 3)
    log_scales[:20] = math.log(0.005)  # small -> duplicate when the gradient is high
    log_scales[20:] = math.log(0.05)   # large -> split
    logit_op[:] = 2.0
    logit_op[[3, 25]] = -8.0           # pruned (sigmoid < 0.005)
    a = GaussianArena(means, quats, log_scales, logit_op, sh)
    a.exp_avg.fill_(1.0)
    st = StrategyState.zeros(N, a.device)
    st.count[:] = 2.0
    st.grad2d[:] = 0.0
    hot = [1, 2, 21, 22, 23]
    st.grad2d[hot] = 2 * 0.01  # mean 0.01 > 5e-4
    gen = torch.Generator().manual_seed(0)
    info = refine_gaussians(a, st, cfg, step=600, generator=gen)
    assert (info["n_dupli"], info["n_split"], info["n_prune"]) == (2, 3, 2)
    assert a.N == N + 2 + 3 * 2 - 3 - 2 == info["n"]
    m = a.view(a.param, "means")
    # order: survivors (without split parents), duplicated copies, then the split children
    assert torch.equal(m[-6 - 2:-6], means[[1, 2]])
    sc = a.view(a.param, "scales")
    assert torch.allclose(sc[-6:], torch.log(torch.exp(log_scales[[21, 22, 23]]) / 1.6).repeat(2, 1))
    # Adam moments: kept for survivors, zero for new ones
    ea = a.view(a.exp_avg, "means")
    assert float(ea[:-8].min()) == 1.0 and float(ea[-8:].abs().max()) == 0.0
    assert st.grad2d.numel() == a.N and float(st.count.sum()) == 0.0
    # same seed -> same result (replica determinism)
    a2 = GaussianArena(means, quats, log_scales, logit_op, sh)
    st2 = StrategyState.zeros(N, a.device)
    st2.count[:] = 2.0
    st2.grad2d[hot] = 0.02
    refine_gaussians(a2, st2, cfg, 600, torch.Generator().manual_seed(0))
    assert torch.equal(a2.param, a.param)


def test_reset_opacities_and_schedule():
    cfg = TrainConfig()
    tr = SplatTrainer(*_params(30, 4), cfg=cfg, backend="torch")
    reset_opacities(tr.arena, cfg)
    cap = math.log(0.01 / 0.99)
    assert float(tr.arena.view(tr.arena.param, "opacities").max()) <= cap + 1e-6
    fired = []
    for step in (0, 100, 500, 600, 700, 2900, 3000, 3100, 3200, 14900, 15000):
        tr.state.count[:] = 1.0
        info = tr.maybe_refine(step)
        fired.append(info is not None)
    # refine only after warm-up, every 100, paused for one interval after each opacity reset, stops at 15000
    assert fired == [False, False, False, True, True, True, False, True, True, True, False]


def test_accumulate_stats_matches_gsplat_formula():
    tr = SplatTrainer(*_params(20, 5), backend="torch")
    g = torch.Generator().manual_seed(1)
    C, N, W, H = 2, 20, 64, 48
    absgrad = torch.rand(C, N, 2, generator=g)
    radii = torch.randint(0, 5, (C, N), generator=g, dtype=torch.int32)
    tr.accumulate_stats(absgrad, radii, W, H, packed=False)
    sel = radii > 0
    gg = absgrad * torch.tensor([W / 2.0 * C, H / 2.0 * C])
    assert torch.allclose(tr.state.grad2d, (gg.norm(dim=-1) * sel).sum(0))
    assert torch.equal(tr.state.count, sel.sum(0).float())
    assert torch.allclose(tr.state.radii, (radii.float() / 64 * sel).max(0).values)


def _worker(rank, world, port, tmp):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    cfg = TrainConfig()
    tr = SplatTrainer(*_params(64, 7), cfg=cfg, rank=rank, world_size=world, backend="torch")
    g = torch.Generator().manual_seed(100 + rank)  # every rank sees different views -> different grads / stats
    for it in range(3):
        tr.arena.grad.copy_(torch.randn(tr.arena.grad.shape, generator=g) * 1e-3 / world)
        absgrad = torch.rand(1, tr.arena.N, 2, generator=g) * 1e-4
        radii = torch.randint(0, 4, (1, tr.arena.N), generator=g, dtype=torch.int32)
        tr.accumulate_stats(absgrad, radii, 64, 64, packed=False, n_cameras=world)
        tr._all_reduce(tr.arena.grad)
        tr.optimizer_step()
        tr.step_count += 1
    tr.state.grad2d += 1.0 * (rank + 1) * (torch.arange(tr.arena.N) % 5 == 0)  # make some Gaussians hot
    info = tr.maybe_refine(600)
    torch.save({"param": tr.arena.param, "m": tr.arena.exp_avg, "N": tr.arena.N, "info": info, "grad2d_sum": float(tr.state.grad2d.sum())},
               os.path.join(tmp, f"r{rank}.pt"))
    dist.barrier()
    dist.destroy_process_group()


def test_gloo_world2_replicas_stay_identical(tmp_path):
    world, port = 2, 29500 + (os.getpid() % 2000)
    mp.spawn(_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    r0, r1 = (torch.load(tmp_path / f"r{r}.pt") for r in range(world))
    assert r0["N"] == r1["N"] and r0["info"] == r1["info"] and r0["info"] is not None and r0["info"]["n_dupli"] + r0["info"]["n_split"] > 0
    assert torch.equal(r0["param"], r1["param"]) and torch.equal(r0["m"], r1["m"])


def test_sharded_gradients_equal_single_rank_sum():
    """loss = mean over ALL views: rank-local gradients carry 1/total_views, so SUM all-reduce == single rank."""
    g = torch.Generator().manual_seed(0)
    per_view = [torch.randn(59 * 8, generator=g) for _ in range(4)]
    single = sum(per_view) / 4
    sharded = sum((per_view[r * 2] + per_view[r * 2 + 1]) / 2 * (2 / 4) for r in range(2))
    assert torch.allclose(single, sharded, atol=1e-7)


def test_arena_layout_matches_build_and_adopt():
    """`GaussianArena.layout` (used by the fused densify to size the new arenas) is the layout `_build` produces, and
    `adopt` leaves the arena in the state a rebuild from the same tensors would."""
    import torch

    from qed_splatter_b200.trainer import GROUPS, GaussianArena

    for N in (1, 7, 10, 33):
        g = torch.Generator().manual_seed(N)
        a = GaussianArena(torch.randn(N, 3, generator=g), torch.randn(N, 4, generator=g), torch.randn(N, 3, generator=g),
                          torch.randn(N, generator=g), torch.randn(N, 16, 3, generator=g))
        off, n = GaussianArena.layout(N)
        assert off == a.offsets and n == a.param.numel()
        assert all(off[k][0] % 4 == 0 for k in GROUPS)  # 16-byte aligned groups
        b = GaussianArena(torch.zeros(1, 3), torch.zeros(1, 4), torch.zeros(1, 3), torch.zeros(1), torch.zeros(1, 16, 3))
        b.adopt(N, a.param.clone(), a.exp_avg.clone(), a.exp_avg_sq.clone())
        assert b.N == a.N and b.offsets == a.offsets and torch.equal(b.group_ends, a.group_ends)
        assert torch.equal(b.param, a.param) and float(b.grad.abs().sum()) == 0.0 and b.grad.numel() == n
