#!/usr/bin/env python
"""Gradient-arena all-reduce over NVLink: this library's one-kernel two-shot reduction (qed_comm_allreduce_f32, NVLS
multimem or peer loads) against NCCL and torch's symmetric-memory op, on the arena size of the bench workload.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 benchmarks/comm_bench.py [--gaussians 1000000]

Checks (before timing): result == NCCL all-reduce of the same data to float tolerance; replicas bit-identical; a
sub-range call leaves the rest untouched."""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

from qed_splatter_b200.comm import SymmetricArena  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gaussians", type=int, default=1_000_000)
    ap.add_argument("--iters", type=int, default=20)
    a = ap.parse_args()
    world, rank, local = int(os.environ["WORLD_SIZE"]), int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    n = 59 * a.gaussians
    n = (n + 3) // 4 * 4
    res = {"world": world, "floats": n, "mbytes": n * 4 / 1e6}
    g = torch.Generator(device=dev).manual_seed(100 + rank)
    data = torch.randn(n, device=dev, generator=g)
    want = data.clone()
    dist.all_reduce(want)

    def bit_identical(t):
        chk = torch.stack([t.double().sum(), t[::97].double().abs().sum(), t.view(torch.int32).sum().double()])
        allc = [torch.empty_like(chk) for _ in range(world)]
        dist.all_gather(allc, chk)
        return all(torch.equal(allc[0], c) for c in allc)

    def timed(fn, iters):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        for _ in range(3):
            fn()
        dist.barrier()
        torch.cuda.synchronize()
        e0.record()
        for _ in range(iters):
            fn()
        e1.record()
        torch.cuda.synchronize()
        t = torch.tensor([e0.elapsed_time(e1) / iters], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    for force_peer in (False, True):
        arena = SymmetricArena(n, dev, force_peer_path=force_peer)
        arena.reduce_in_switch = bool(arena.multicast_ptr)  # this benchmark times both paths at every world size
        key = "nvls-multimem" if arena.reduce_in_switch else "peer-load-store"
        if force_peer is False and not arena.multicast_ptr:
            res["nvls"] = "no multicast support on this box"
            continue
        arena.buf.copy_(data)
        arena.all_reduce_()
        torch.cuda.synchronize()
        err = float((arena.buf - want).abs().max()) / float(want.abs().mean())
        ident = bit_identical(arena.buf)
        # sub-range: only [lo, hi) changes
        arena.buf.copy_(data)
        lo, hi = 11 * a.gaussians // 4 * 4, (11 * a.gaussians + 48 * (a.gaussians // 3)) // 4 * 4
        arena.all_reduce_(lo, hi)
        torch.cuda.synchronize()
        sub_ok = bool(torch.equal(arena.buf[:lo], data[:lo]) and torch.equal(arena.buf[hi:], data[hi:]) and
                      float((arena.buf[lo:hi] - want[lo:hi]).abs().max()) / float(want.abs().mean()) < 1e-5)
        times = {}
        for blocks in (16, 32, 64, 128, 256):
            times[blocks] = timed(lambda b=blocks: arena.all_reduce_(blocks=b), a.iters)
        res[key] = {"max_err_over_mean_abs": err, "replicas_bit_identical": ident, "subrange_ok": sub_ok, "ms_by_blocks": times,
                    "best_ms": min(times.values()), "algbw_gbs": n * 4 / 1e6 / min(times.values())}
        del arena
    nccl_buf = data.clone()
    res["nccl_all_reduce_ms"] = timed(lambda: dist.all_reduce(nccl_buf), a.iters)
    try:
        import torch.distributed._symmetric_memory as symm_mem

        t = symm_mem.empty(n, dtype=torch.float32, device=dev)
        symm_mem.rendezvous(t, dist.group.WORLD)
        t.copy_(data)
        name = dist.group.WORLD.group_name
        res["torch_multimem_all_reduce_ms"] = timed(lambda: torch.ops.symm_mem.multimem_all_reduce_(t, "sum", name), a.iters)
        res["torch_two_shot_all_reduce_ms"] = timed(lambda: torch.ops.symm_mem.two_shot_all_reduce_(t, "sum", name), a.iters)
    except Exception as e:  # library comparison only
        res["torch_symm_mem_ops"] = f"unavailable: {type(e).__name__}: {e}"[:300]
    if rank == 0:
        print(json.dumps(res))
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
