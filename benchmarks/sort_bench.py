#!/usr/bin/env python
"""This library's stable LSD radix sort (qed_sort_pairs: 64-bit key, 32-bit value pairs, the reference formulation of
gsplat's isect_tiles) against cub::DeviceRadixSort (qed_sort_pairs_cub, the library baseline gsplat calls) on key
distributions shaped like the intersection lists: 13 tile bits | 32 depth bits, 45 key bits sorted.

    python benchmarks/sort_bench.py            # one JSON line per size"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

from qed_splatter_b200 import _lib, ops  # noqa: E402


def main():
    lib = _lib.load()
    g = torch.Generator().manual_seed(0)
    for n in (668_000, 6_570_000, 50_000_000, 150_000_000):
        keys = (torch.randint(0, 1 << 13, (n,), generator=g, dtype=torch.int64) << 32 | torch.randint(0, 1 << 32, (n,), generator=g, dtype=torch.int64)).cuda()
        vals = torch.arange(n, dtype=torch.int32).cuda()
        ref, row = None, {"pairs": n, "key_bits": 45}
        for mode, name in ((0, "own_three_kernel_passes"), (2, "own_lookback_passes"), (1, "own_default"), (-1, "cub")):
            if mode >= 0:
                lib.qed_debug_set_radix_onesweep(mode)
            impl = "cub" if mode < 0 else "own"
            for _ in range(2):
                ko, vo = ops.sort_pairs(keys, vals, 45, impl=impl)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(5):
                ko, vo = ops.sort_pairs(keys, vals, 45, impl=impl)
            e1.record()
            torch.cuda.synchronize()
            if ref is None:
                ref = (ko.clone(), vo.clone())
            row[name + "_ms"] = e0.elapsed_time(e1) / 5
            row[name + "_equal"] = bool(torch.equal(ko, ref[0]) and torch.equal(vo, ref[1]))
            del ko, vo
        lib.qed_debug_set_radix_onesweep(1)
        row["own_over_cub"] = row["own_default_ms"] / row["cub_ms"]
        print(json.dumps(row), flush=True)
        del keys, vals, ref
        torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
