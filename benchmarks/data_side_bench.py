#!/usr/bin/env python
"""Data-side kernels on one B200: depth back-projection and voxel merge of `qed-init-pc`
(create_init_pointcloud.py:148-261) at a 1920x1080 uint16 depth frame, CUDA-event timed, against the HBM roofline
(MEASURED_PEAKS.json) with the CPU oracle (numpy restatement of the Open3D calls) timed beside it.

    python benchmarks/data_side_bench.py > gpurun_out/r02_data_side.jsonl

Algorithmic bytes: back-projection reads 2 B/pixel twice (flag scan + final phase) and writes 12 B per surviving point;
voxel merge = key pass (12 B read, 12 B written per point) + radix sort of (int64, int32) pairs over 63 bits (8 passes x
24 B) + head scan + mean (12 B gathered per point)."""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402
import torch  # noqa: E402

from oracle import pointcloud as opc  # noqa: E402  (CPU baseline leg only)
from qed_splatter_b200 import data_side  # noqa: E402


def timed(fn, reps=20, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def main():
    peak = 6537.0
    try:
        peak = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
    except Exception:
        pass
    dev = torch.device("cuda", 0)
    H, W = 1080, 1920
    g = np.random.default_rng(0)
    yy, xx = np.mgrid[0:H, 0:W]
    depth = (1500 + 800 * np.sin(xx / 97.0) + 600 * np.cos(yy / 61.0) + g.uniform(0, 40, size=(H, W))).astype(np.uint16)
    depth[g.uniform(size=(H, W)) < 0.1] = 0
    K = np.array([[1400.0, 0, W / 2], [0, 1400.0, H / 2], [0, 0, 1]], dtype=np.float32)
    w2c = np.eye(4, dtype=np.float32)
    d_t = torch.from_numpy(depth.astype(np.int16)).view(torch.uint16).to(dev)

    pts = data_side.backproject_frame(d_t, K, w2c, 0.001, 100.0, 1, None)
    n = pts.shape[0]
    ms_bp = timed(lambda: data_side.backproject_frame(d_t, K, w2c, 0.001, 100.0, 1, None))
    bytes_bp = 2 * 2 * H * W + 12 * n
    out = {"kernel": "backproject_depth (flag scan + compacting final phase)", "frame": f"{W}x{H} uint16", "points": n, "ms": ms_bp,
           "algorithmic_bytes": bytes_bp, "achieved_gbps": bytes_bp / ms_bp / 1e6, "peak_gbps": peak, "frac": bytes_bp / ms_bp / 1e6 / peak,
           "note": "includes the host read of the point count (one sync per frame)"}
    t0 = time.perf_counter()
    ref = opc.backproject_depth(depth, K, w2c, 0.001, 100.0, 1)
    out["cpu_oracle_ms"] = (time.perf_counter() - t0) * 1e3
    out["equal_to_oracle"] = bool(np.array_equal(pts.cpu().numpy(), ref))
    print(json.dumps(out), flush=True)

    for vs in (0.05, 0.01):
        res = data_side.voxel_down_sample(pts, vs)
        ms_v = timed(lambda: data_side.voxel_down_sample(pts, vs))
        bytes_v = n * (24 + 8 * 24 + 12 + 12) + 12 * res.shape[0]
        o = {"kernel": "voxel_down_sample (keys + 63-bit radix sort + head scan + run means)", "points": n, "voxel_size": vs, "voxels": int(res.shape[0]),
             "ms": ms_v, "algorithmic_bytes": bytes_v, "achieved_gbps": bytes_v / ms_v / 1e6, "peak_gbps": peak, "frac": bytes_v / ms_v / 1e6 / peak}
        t0 = time.perf_counter()
        r = opc.voxel_down_sample(ref, vs)
        o["cpu_oracle_ms"] = (time.perf_counter() - t0) * 1e3
        o["same_voxel_count"] = bool(r.shape[0] == res.shape[0])
        o["max_abs_diff"] = float(np.abs(r - res.cpu().numpy()).max()) if r.shape == tuple(res.shape) else None
        print(json.dumps(o), flush=True)


if __name__ == "__main__":
    main()
