#!/usr/bin/env python
"""bench.py — the driver's measurement contract for the depth-supervised splat hot path.

    python bench.py --gpus N --steps K --warmup W            (N>1: launched by torchrun, one rank per GPU)
    python bench.py --impl reference --gpus N --steps K --warmup W

Workload = BASELINE.json configs[1]: synthetic 1M-Gaussian scene, one 1920x1080 view per GPU,
RGB+ED forward + qed-splatter loss (RGB-L1 + masked depth-L1) + full backward to the six parameter
tensors.  A "step" is one pass of that hot path over one view per GPU; for N>1 the views are sharded over
the ranks against a replicated Gaussian set and the step ends with the NCCL all-reduce of the flat
gradient arena and of the densification statistics (weak scaling).

metric  = fwd+bwd RGB+depth Mpix/s (whole job: N*H*W*K / time, device-timed, max over ranks).
value   = fused C-ABI pipeline with everything resident in HBM.
e2e     = the same work through the public API a qed-splatter maintainer binds: `rasterization()`
          (gsplat surface, torch autograd) + `depth_supervised_loss()` (model.py:295-306, 73-118 as one autograd op;
          `--torch-loss` writes those lines with torch ops as the reference does) + `backward()`, with the step's
          inputs (camera, ground-truth RGB as the uint8 image cache of config.py:37 + float depth) copied from pinned host
          memory and the loss read back every
          step.  The Gaussian parameters are model state and stay resident, as in the reference.  For N>1 the
          five parameter gradients go out as one coalesced NCCL group call (as DDP / FSDP issue them).
train   = full trainer iterations/s (trainer.SplatTrainer: 0.8 L1 + 0.2 (1-SSIM) + 0.2 depth-L1, Adam, strategy
          statistics, pipelined gradient all-reduce) -- the second half of BASELINE.json's metric.
--impl reference = the CPU arm: the oracle (`oracle/`, a port — the reference's own arithmetic lives in the
          un-vendored gsplat and cannot be run here) on the host cores on a bounded 1/16 sample.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "fwd+bwd RGB+depth Mpix/s"
UNIT = "Mpix/s"
PARAM_FLOATS = 3 + 4 + 3 + 1 + 48  # means, quats, scales, opacity, SH(16x3) = 59 floats / Gaussian

# dram__bytes_read.sum + dram__bytes_write.sum per launch from the committed `ncu --set full` captures of this
# workload (profiles/r01_raster_ncu_summary.txt: raster_bwd_ws_kernel / raster_fwd_ws_kernel on the exact tile lists); None = not captured
NCU_DRAM_BYTES = {"raster_bwd": 140.47e6 + 7.91e6, "raster_fwd": 41.19e6 + 12.53e6}

# SURVEY.md §8(d) per-unit figures (D = 4 channels)
FLOP_PER_PAIR_FWD = 30.0
FLOP_PER_PAIR_BWD = 100.0
BYTES_PROJ_FWD_PER_GAUSS = 44.0
BYTES_PROJ_FWD_PER_VISIBLE = 192.0 + 100.0  # SH read + records written
BYTES_PROJ_BWD_PER_GAUSS = 44.0 + 4.0 + 236.0  # inputs + radii read, gradients written
BYTES_PROJ_BWD_PER_VISIBLE = 48.0 + 12.0 + 192.0  # packed grads + conic + SH read


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--gaussians", type=int, default=1_000_000)
    ap.add_argument("--width", type=int, default=1920)
    ap.add_argument("--height", type=int, default=1080)
    ap.add_argument("--mode", default="RGB+ED", choices=["RGB+ED", "RGB+D"])
    ap.add_argument("--sort", default="two_level", choices=["two_level", "own", "cub"])
    ap.add_argument("--comm-chunks", type=int, default=1, help="Gaussian ranges of the projection backward whose SH gradients are all-reduced while the next range computes (N>1)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--e2e-bucket", action="store_true", help="e2e, N>1: all-reduce one flat bucket of gradient views instead of a coalesced group call")
    ap.add_argument("--gt-float", action="store_true", help="e2e: ground-truth RGB as float32 instead of the uint8 image cache")
    ap.add_argument("--torch-loss", action="store_true", help="e2e: write the loss with torch ops as the reference does instead of depth_supervised_loss")
    ap.add_argument("--no-train", action="store_true")
    return ap.parse_args()


# ------------------------------------------------------------------------------------------------
# clocks
# ------------------------------------------------------------------------------------------------
class ClockSampler:
    FIELDS = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.gpu = gpu_index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.FIELDS}", "--format=csv,noheader,nounits", "-lms", "100",
                                          "-i", str(self.gpu)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons, power = [], None, set(), []
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                mx = float(f[2])
                power.append(float(f[3]))
            except ValueError:
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons),
                "samples": len(sm), "power_w_max": max(power) if power else None}


# ------------------------------------------------------------------------------------------------
# CPU arm (oracle) on a bounded sample of the workload
# ------------------------------------------------------------------------------------------------
def cpu_sample_step(sample, torch, oracle):
    names = ("means", "quats", "scales", "opacities", "sh")
    leaves = {k: getattr(sample, k).clone().requires_grad_(True) for k in names}
    r, a, _ = oracle.rasterization(leaves["means"], leaves["quats"], leaves["scales"], leaves["opacities"], leaves["sh"],
                                   sample.viewmats, sample.Ks, sample.width, sample.height, sh_degree=3, render_mode="RGB+ED")
    rgb, d = oracle.composite_and_fill(r, a, torch.tensor([0.1, 0.2, 0.3]))
    loss = oracle.rgb_l1_loss(rgb, sample.gt_rgb) + oracle.depth_l1_loss(d, sample.gt_depth, 0.2)
    loss.backward()
    return float(loss)


def make_cpu_sample(args, n_steps_total: int = 3):
    """A bounded sample of the workload for the CPU arm: 1/k of the Gaussians and 1/k of the pixels at the same field
    of view, with k chosen so that `n_steps_total` oracle steps finish in a few minutes (~5 s per step at k = 16 on
    16 cores)."""
    from qed_splatter_b200.scenes import scene_s1

    k = 16 if n_steps_total <= 8 else (64 if n_steps_total <= 40 else 256)
    r = int(round(k ** 0.5))
    n = max(args.gaussians // k, 1000)
    w, h = args.width // r, args.height // r
    return scene_s1(N=n, width=w, height=h, f=1200.0 * w / 1920.0), f"1/{k} sample: {n} Gaussians, one {w}x{h} view, RGB+ED fwd+loss+bwd, float32"


def run_cpu_baseline(args, steps: int, warmup: int):
    import torch

    import oracle

    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    sample, desc = make_cpu_sample(args, steps + warmup)
    for _ in range(warmup):
        cpu_sample_step(sample, torch, oracle)
    t0 = time.perf_counter()
    for _ in range(steps):
        cpu_sample_step(sample, torch, oracle)
    dt = (time.perf_counter() - t0) / max(steps, 1)
    mpix = sample.width * sample.height / dt / 1e6
    return {"value": mpix, "unit": UNIT, "cores": cores, "kind": "port", "sample": desc, "ms_per_step": dt * 1e3}


def main_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cb = run_cpu_baseline(args, steps=max(args.steps, 1), warmup=max(args.warmup, 0))
    line = {
        "impl": "reference", "metric": METRIC, "value": cb["value"], "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": cb["ms_per_step"], "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"S1: {args.gaussians} Gaussians, one {args.width}x{args.height} view per GPU, RGB+ED fwd+loss+bwd",
                   "note": "CPU oracle port (gsplat reference rasterizer restated in torch) on host cores, bounded sample"},
        "cpu_baseline": {k: cb[k] for k in ("value", "unit", "cores", "kind", "sample")},
        "e2e": {"value": cb["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


# ------------------------------------------------------------------------------------------------
# GPU arm
# ------------------------------------------------------------------------------------------------
def main_ours(args):
    import torch
    import torch.distributed as dist

    from qed_splatter_b200 import _lib, depth_supervised_loss, ops, rasterization
    from qed_splatter_b200.pipeline import FusedSplatStep
    from qed_splatter_b200.scenes import scene_s1

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (no CPU fallback); use --impl reference for the CPU arm")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    lib = _lib.load()
    K, W_ = args.steps, max(args.warmup, 3)
    width, height, N = args.width, args.height, args.gaussians

    # replicated Gaussians (same seed everywhere), one view per rank
    scene = scene_s1(N=N, width=width, height=height, C=max(world, 1))
    sl = slice(rank, rank + 1)
    means, quats, scales, opac, sh = (t.to(dev) for t in (scene.means, scene.quats, scene.scales, scene.opacities, scene.sh))
    viewmats, Ks = scene.viewmats[sl].contiguous().to(dev), scene.Ks[sl].contiguous().to(dev)
    gt_rgb_h = scene.gt_rgb[sl].contiguous().pin_memory()
    gt_depth_h = scene.gt_depth[sl].contiguous().pin_memory()
    vm_h, K_h = scene.viewmats[sl].contiguous().pin_memory(), scene.Ks[sl].contiguous().pin_memory()
    gt_rgb, gt_depth = gt_rgb_h.to(dev), gt_depth_h.to(dev)
    bg = torch.tensor([0.1, 0.2, 0.3], device=dev)
    del scene

    fs = FusedSplatStep(dev, sort_impl=args.sort)
    arena = torch.zeros(PARAM_FLOATS * N, device=dev)  # flat gradient arena: one all-reduce
    o = 0
    views = {}
    for name, shape in (("means", (N, 3)), ("quats", (N, 4)), ("scales", (N, 3)), ("opacities", (N,)), ("sh", (N, 16, 3))):
        n = 1
        for s_ in shape:
            n *= s_
        views[name] = arena[o:o + n].view(*shape)
        o += n
    stats = torch.zeros(3, N, device=dev)  # grad2d, count, radii_max (DefaultStrategy state)

    n_chunks = args.comm_chunks if world > 1 else 1
    sh_begin = 11 * N  # arena = [means 3N | quats 4N | scales 3N | opacities N | sh 48N]
    pending = []

    def on_chunk(k, n0, n1):
        # the SH block is 81 % of the gradient bytes: reduce each finished Gaussian range while the projection
        # backward of the next range is still running (NCCL runs on its own stream)
        pending.append(dist.all_reduce(arena[sh_begin + 48 * n0: sh_begin + 48 * n1], async_op=True))

    def step():
        out = fs.step(means, quats, scales, opac, sh, viewmats, Ks, width, height, 3, gt_rgb, gt_depth, bg, render_mode=args.mode,
                      grad_scale=1.0 / world, grad_out=views, n_chunks=n_chunks, on_chunk=on_chunk if world > 1 else None)
        _lib.check(lib.qed_strategy_update(1, N, _lib.ptr(out.packed_grads), 1, _lib.ptr(out.radii), width, height, world, _lib.ptr(stats[0]),
                                           _lib.ptr(stats[1]), _lib.ptr(stats[2]), _lib.current_stream()), "qed_strategy_update")
        if world > 1:
            # gradients every step (SH ranges already in flight, then the 11 small floats per Gaussian); the
            # densification accumulators are reduced only when a refine step consumes them (SUM / MAX are
            # associative), exactly as trainer.SplatTrainer does
            pending.append(dist.all_reduce(arena[:sh_begin], async_op=True))
            for w in pending:
                w.wait()
            pending.clear()
        return out

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(W_):
        out = step()
    barrier()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for _ in range(K):
        out = step()
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1)
    clocks = sampler.stop() if rank == 0 else None
    t = torch.tensor([ms], device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_total = float(t.item())
    ms_per_step = ms_total / K
    value = world * width * height / (ms_per_step * 1e-3) / 1e6

    # ---- stage timing (same inputs, CUDA events between the stages on the launching stream) ----
    stage_ms = {}
    for _ in range(K):
        fs.marks = []
        out = step()
        torch.cuda.synchronize()
        m = fs.marks
        for (_, a), (name, b) in zip(m[:-1], m[1:]):
            stage_ms[name] = stage_ms.get(name, 0.0) + a.elapsed_time(b) / K
    fs.marks = None
    n_visible = int((out.radii > 0).sum())
    M = out.n_isects
    loss = [float(x) for x in out.loss.tolist()]

    # ---- pair counters for the compositing roofline (instrumented launches, outside any timed region) ----
    counters = fs.count_pairs()
    launches_per_step = fs.launches_per_step + ((n_chunks - 1) if world > 1 else 0)  # this library's kernels only (chunked: extra project_bwd launches)

    # ---- second half of the metric: train iters/s (full qed-splatter step through trainer.SplatTrainer: render,
    # 0.8 L1 + 0.2 (1-SSIM) + 0.2 depth-L1, backward, gradient all-reduce pipelined with Adam, strategy statistics) ----
    train = None
    if not args.no_train:
        from qed_splatter_b200.trainer import SplatTrainer, TrainConfig

        tcfg = TrainConfig(render_mode="RGB+D" if args.mode == "RGB+D" else "RGB+ED", refine_every=10 ** 9)
        tr = SplatTrainer(means.clone(), quats.clone(), torch.log(scales), torch.logit(opac.clamp(1e-4, 1 - 1e-4)), sh.clone(), cfg=tcfg,
                          rank=rank, world_size=world, backend="cuda")
        tr.step_count = 3000  # SH degree 3, no opacity reset inside the timed window
        for _ in range(W_):
            tr.step(viewmats, Ks, width, height, gt_rgb, gt_depth, bg, total_views=world)
        barrier()
        e0.record()
        for _ in range(K):
            tr.step(viewmats, Ks, width, height, gt_rgb, gt_depth, bg, total_views=world)
        e1.record()
        barrier()
        t = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        tms = float(t.item()) / K
        train = {"train_iters_per_s": 1e3 / tms, "ms_per_step": tms, "views_per_s": world * 1e3 / tms,
                 "loss": "0.8 L1 + 0.2 (1-SSIM) + 0.2 depth-L1", "optimizer": "fused Adam over the 59-float/Gaussian arena",
                 "comm_chunks": tcfg.comm_chunks if world > 1 else 1}
        del tr
        torch.cuda.empty_cache()

    # ---- e2e through the public API with host buffers ----
    e2e = None
    if not args.no_e2e:
        params = [t.clone().requires_grad_(True) for t in (means, quats, scales, opac, sh)]
        # N > 1: the five parameter gradients are reduced as ONE coalesced NCCL group call (what DDP / FSDP do), no
        # flattening copies; `--e2e-bucket` uses a flat bucket of gradient views (gradient_as_bucket_view) instead
        grad_bucket, grad_views = None, None
        if world > 1 and args.e2e_bucket:
            grad_bucket = torch.zeros(sum(p_.numel() for p_ in params), device=dev)
            grad_views, o_ = [], 0
            for p_ in params:
                grad_views.append(grad_bucket[o_:o_ + p_.numel()].view_as(p_))
                o_ += p_.numel()
        # the step's inputs come from pinned host memory on a copy stream (what a data loader does); the loss
        # waits on the copy event, so the H2D of the ground truth overlaps projection/sort/compositing
        copy_stream = torch.cuda.Stream()
        vm_d, K_d = torch.empty_like(viewmats), torch.empty_like(Ks)
        # ground-truth RGB travels as the reference's data side holds it: the uint8 image cache (config.py:37
        # cache_images_type="uint8"); depth_supervised_loss converts it in-kernel (--torch-loss: torch .float() / 255)
        e2e_rgb_h = gt_rgb_h if args.gt_float else (gt_rgb_h * 255.0).round().clamp(0, 255).to(torch.uint8).contiguous().pin_memory()
        rgb_d, depth_d = torch.empty(e2e_rgb_h.shape, dtype=e2e_rgb_h.dtype, device=dev), torch.empty_like(gt_depth)
        cam_ready, gt_ready = torch.cuda.Event(), torch.cuda.Event()

        def e2e_step():
            copy_stream.wait_stream(torch.cuda.current_stream())  # previous step is done with the buffers
            with torch.cuda.stream(copy_stream):
                vm_d.copy_(vm_h, non_blocking=True)
                K_d.copy_(K_h, non_blocking=True)
                cam_ready.record(copy_stream)
                rgb_d.copy_(e2e_rgb_h, non_blocking=True)
                depth_d.copy_(gt_depth_h, non_blocking=True)
                gt_ready.record(copy_stream)
            torch.cuda.current_stream().wait_event(cam_ready)
            vm, Kc, rgb_gt, d_gt = vm_d, K_d, rgb_d, depth_d
            if grad_bucket is None:
                for p_ in params:
                    p_.grad = None
            else:
                grad_bucket.zero_()
                for p_, g_ in zip(params, grad_views):
                    p_.grad = g_  # autograd accumulates in place
            render, alpha, info = rasterization(params[0], params[1], params[2], params[3], params[4], vm, Kc, width, height,
                                                tile_size=16, packed=False, near_plane=0.01, far_plane=1e10, render_mode=args.mode,
                                                sh_degree=3, sparse_grad=False, absgrad=True, rasterize_mode="classic")
            info["means2d"].retain_grad()
            torch.cuda.current_stream().wait_event(gt_ready)
            if args.torch_loss:
                rgb_gt = rgb_gt.float() / 255.0 if rgb_gt.dtype == torch.uint8 else rgb_gt  # splatfacto get_gt_img
                # qed_splatter/model.py:295-306 and :87-116 as the reference writes them (a dozen torch element-wise ops)
                rgb = torch.clamp(render[..., :3] + (1 - alpha) * bg, 0.0, 1.0)
                depth = render[..., 3:4]
                depth = torch.where(alpha > 0, depth, depth.detach().max())
                valid = torch.isfinite(depth) & torch.isfinite(d_gt) & (d_gt > 0)
                l_rgb = 0.8 * (rgb_gt - rgb).abs().mean()
                l_d = 0.2 * ((depth - d_gt).abs() * valid).sum() / valid.sum().clamp(min=1)
                loss_t = (l_rgb + l_d) / world
            else:
                # the same lines through the package's drop-in for them (losses.depth_supervised_loss)
                loss_t = depth_supervised_loss(render, alpha, rgb_gt, d_gt, bg, rgb_weight=0.8, depth_lambda=0.2)[0] / world
            loss_t.backward()
            if world > 1:
                if grad_bucket is not None:
                    dist.all_reduce(grad_bucket)
                else:
                    with dist._coalescing_manager(device=dev):
                        for p_ in params:
                            dist.all_reduce(p_.grad)
            return float(loss_t.item())  # D2H read of the step's result

        for _ in range(W_):
            e2e_step()
        barrier()
        e0.record()
        for _ in range(K):
            e2e_loss = e2e_step()
        e1.record()
        barrier()
        t = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_ms = float(t.item()) / K
        h2d = vm_h.numel() * 4 + K_h.numel() * 4 + e2e_rgb_h.numel() * e2e_rgb_h.element_size() + gt_depth_h.numel() * 4
        e2e = {"value": world * width * height / (e2e_ms * 1e-3) / 1e6, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 4,
               "ms_per_step": e2e_ms, "loss": e2e_loss,
               "api": "qed_splatter_b200.rasterization (gsplat surface) + " +
                      ("torch autograd loss as qed_splatter/model.py writes it" if args.torch_loss else
                       "qed_splatter_b200.depth_supervised_loss (model.py:295-306, 73-118 as one autograd op)") + " + backward()"}

    if rank == 0:
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
        hbm_src = "measured (MEASURED_PEAKS.json)" if "hbm_gbs" in peaks else "fallback 6650 GB/s (B200_PROFILING.md)"
        sm_mhz = (clocks or {}).get("sm_mhz") or 1965.0
        fp32_peak = 148 * 128 * 2 * sm_mhz * 1e6 / 1e12  # TFLOP/s at the SM clock seen during the timed region

        def fp32_roof(name, pairs, evaluated, flop_per_pair):
            # algorithmic work = (pixel, Gaussian) pairs that pass the alpha test and are composited / differentiated:
            # independent of how well an implementation culls; pairs_evaluated (>= pairs) is what this one touched
            t_ms = stage_ms.get(name)
            if not t_ms or not pairs:
                return None
            ach = pairs * flop_per_pair / (t_ms * 1e-3) / 1e12
            return {"kernel": name, "bound": "fp32", "achieved": ach, "peak": fp32_peak, "unit": "TFLOP/s", "frac": ach / fp32_peak,
                    "traffic": NCU_DRAM_BYTES.get(name), "ms": t_ms, "pairs_composited": pairs, "pairs_evaluated": evaluated,
                    "flop_per_pair": flop_per_pair,
                    "peak_source": f"derived: 148 SM x 128 lanes x 2 FLOP x {sm_mhz:.0f} MHz (median SM clock sampled during the timed region)"}

        def hbm_roof(name, bytes_):
            t_ms = stage_ms.get(name)
            if not t_ms:
                return None
            ach = bytes_ / (t_ms * 1e-3) / 1e9
            return {"kernel": name, "bound": "hbm", "achieved": ach, "peak": hbm_peak, "unit": "GB/s", "frac": ach / hbm_peak,
                    "traffic": NCU_DRAM_BYTES.get(name), "ms": t_ms, "algorithmic_bytes": bytes_, "peak_source": hbm_src}

        tile_bits = (((width + 15) // 16) * ((height + 15) // 16)).bit_length()
        tile_passes = (tile_bits + 7) // 8
        # two-level intersection build: algorithmic bytes
        #   prepare: flags scan (4 r + 8 w + 8 r per (c,n)), compaction (8 w), 4 radix passes over the visible entries
        #            (4 r upsweep + 8 r + 8 w downsweep each), gather scan (4 + 4 r, 8 w)
        #   fill   : emit 8 w per isect, tile_passes radix passes (20 B each), ranges pass (4 r key + 4 r flat)
        bytes_prepare = N * 20.0 + n_visible * (8.0 + 4 * 20.0 + 16.0)
        bytes_fill = M * (8.0 + tile_passes * 20.0 + 8.0)
        stages = [
            fp32_roof("raster_bwd", counters.get("bwd_pairs_contributing"), counters.get("bwd_pairs_evaluated"), FLOP_PER_PAIR_BWD),
            fp32_roof("raster_fwd", counters.get("fwd_pairs_contributing"), counters.get("fwd_pairs_evaluated"), FLOP_PER_PAIR_FWD),
            hbm_roof("project_fwd", N * BYTES_PROJ_FWD_PER_GAUSS + n_visible * BYTES_PROJ_FWD_PER_VISIBLE),
            hbm_roof("project_bwd", N * BYTES_PROJ_BWD_PER_GAUSS + n_visible * BYTES_PROJ_BWD_PER_VISIBLE),
            hbm_roof("isect_prepare", bytes_prepare),
            hbm_roof("isect_fill", bytes_fill),
            hbm_roof("loss", width * height * (24.0 + 68.0)),
            # the reference formulation (gsplat: 64-bit keys, full radix sort), only present with --sort own|cub
            hbm_roof("sort", M * (((32 + tile_bits + 1 + 7) // 8) * 24.0 + 8.0)),
            hbm_roof("emit", M * 12.0),
        ]
        stages = [s for s in stages if s]
        dominant = max(stages, key=lambda s: s["ms"]) if stages else None
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W_, "ms_per_step": ms_per_step,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {
                "workload": f"S1 (BASELINE.json configs[1]): {N} Gaussians (SH degree 3), one {width}x{height} view per GPU, "
                            f"{args.mode} forward + RGB-L1/depth-L1 loss + backward to means/quats/scales/opacities/SH",
                "parallelism": f"view-sharded x{world}, replicated Gaussians, NCCL all-reduce of the {PARAM_FLOATS * 4} B/Gaussian gradient arena every step "
                               "(densification accumulators are all-reduced when a refine step consumes them, not per step)",
                "cache": "inputs larger than L2: 236 MB parameters + %.0f MB intermediates per step vs 126 MB L2; no explicit flush" % ((N * 150 + M * 24) / 1e6),
                "sort": args.sort, "n_visible": n_visible, "n_isects": M, "loss": loss,
                "mean_gaussians_composited_per_pixel": counters.get("fwd_pairs_contributing", 0) / float(width * height),
            },
            "clocks": clocks,
            "e2e": e2e,
            "train": train,
            "gpu_launches": launches_per_step * K,
            "gpu_launches_per_step": launches_per_step,
            "roofline": dominant,
            "stages": stages,
            "stage_ms": stage_ms,
            "pair_counters": counters,
        }
        if world == 1 and not args.no_cpu_baseline:
            cb = run_cpu_baseline(args, steps=2, warmup=1)
            line["cpu_baseline"] = {k: cb[k] for k in ("value", "unit", "cores", "kind", "sample")}
        else:
            line["cpu_baseline"] = None
        print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    a = parse_args()
    if a.impl == "reference":
        main_reference(a)
    else:
        main_ours(a)
