#!/usr/bin/env python
"""A/B of one of the library's thread-local test hooks on the bench workload (BASELINE configs[1], fused step, one B200):

    python benchmarks/ab_hooks.py qed_debug_set_project_bwd_one 0 5 6
    python benchmarks/ab_hooks.py qed_debug_set_raster_bwd_minb 10 8
    python benchmarks/ab_hooks.py qed_debug_set_flat_scan 0 1

For every value (the list is run twice, interleaved) : K back-to-back steps timed by CUDA events, the per-stage times of
pipeline.FusedSplatStep and the largest relative gradient difference against the first configuration.  This is how the
round-2 choices recorded in profiles/r02_kernel_notes.md (occupancy of the compositors, single-view projection kernels,
single-launch scans, radix tile size / look-back depth) were measured."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

from qed_splatter_b200 import _lib  # noqa: E402
from qed_splatter_b200.pipeline import FusedSplatStep  # noqa: E402
from qed_splatter_b200.scenes import scene_s1  # noqa: E402


def main():
    hook, values = sys.argv[1], [int(v) for v in sys.argv[2:]]
    lib = _lib.load()
    setter = getattr(lib, hook)
    s = scene_s1(N=1_000_000).to("cuda")
    bg = torch.tensor([0.1, 0.2, 0.3], device="cuda")
    gt_d = s.gt_depth.contiguous()
    fs = FusedSplatStep("cuda")

    def run():
        return fs.step(s.means, s.quats, s.scales, s.opacities, s.sh, s.viewmats, s.Ks, s.width, s.height, 3, s.gt_rgb, gt_d, bg)

    ref = None
    for v in values + values:
        old = setter(v)
        for _ in range(5):
            out = run()
        torch.cuda.synchronize()
        K = 40
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(K):
            out = run()
        e1.record()
        torch.cuda.synchronize()
        acc = {}
        for _ in range(10):
            fs.marks = []
            out = run()
            torch.cuda.synchronize()
            m = fs.marks
            for (_, a), (n1, b) in zip(m[:-1], m[1:]):
                acc[n1] = acc.get(n1, 0.0) + a.elapsed_time(b)
            fs.marks = None
        g = {k: t.clone() for k, t in out.grads.items()}
        if ref is None:
            ref = g
        d = max(float((ref[k] - g[k]).abs().max() / (ref[k].abs().max() + 1e-30)) for k in g)
        print(f"{hook}({v}): step {e0.elapsed_time(e1) / K:.4f} ms  " + "  ".join(f"{k[:13]}={t / 10:.4f}" for k, t in acc.items() if k != "sync") +
              f"  max rel grad diff vs first {d:.2e}", flush=True)
        setter(old)


if __name__ == "__main__":
    main()
