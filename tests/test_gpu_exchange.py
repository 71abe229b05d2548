"""GPU (one device): the arithmetic of the multi-GPU view-colour exchange (include/qed_splat.h, csrc/project_bwd.cu,
csrc/comm.cu) without a second GPU.  "Ranks" are emulated on one device: every emulated rank runs
qed_project_bwd_exchange for ITS view with plain peer pointers (no multicast) into the same exchange buffer; then
  sum_ranks(v_means, v_quats, v_scales, v_opacities)  and  qed_sh_grad_from_view_colors(all slots)
must equal ONE qed_project_bwd over the whole view batch (the single-rank step).  The real N-rank run over NVLink is
checked by bench.py's `multi_gpu_check` (recorded in SCALE_r02.json) and tests/multi_gpu_check.py."""
import ctypes

import pytest
import torch

from helpers import assert_close_frac
from qed_splatter_b200 import _lib
from qed_splatter_b200.pipeline import FusedSplatStep
from qed_splatter_b200.scenes import scene_s0

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("deg,C", [(3, 3), (1, 2), (0, 1)])
def test_exchange_equals_batched_projection_backward(cuda, deg, C):
    lib = _lib.load()
    s = scene_s0(N=5000, C=C, size=96).to(cuda)
    bg = torch.tensor([0.3, 0.2, 0.1], device=cuda)
    fs = FusedSplatStep(cuda)
    whole = fs.step(s.means, s.quats, s.scales, s.opacities, s.sh, s.viewmats, s.Ks, s.width, s.height, deg, s.gt_rgb, s.gt_depth, bg)
    ref = {k: v.clone() for k, v in whole.grads.items()}
    N, K = s.N, s.sh.shape[1]
    xch = torch.zeros((C * N + C) * 4, device=cuda)
    peers = (ctypes.c_void_p * 1)(xch.data_ptr())
    tag = 7.0
    acc = {k: torch.zeros_like(ref[k]) for k in ("means", "quats", "scales", "opacities")}
    for r in range(C):  # emulated rank r renders view r; gradients pre-scaled by 1 / total views as in the sharded step
        sl = slice(r, r + 1)
        fs.step(s.means, s.quats, s.scales, s.opacities, s.sh, s.viewmats[sl].contiguous(), s.Ks[sl].contiguous(), s.width, s.height, deg,
                s.gt_rgb[sl].contiguous(), s.gt_depth[sl].contiguous(), bg, grad_scale=1.0 / C)
        f = fs._fwd
        packed = fs._get("packed", (N, 12))
        out = {k: torch.empty_like(ref[k]) for k in acc}
        _lib.check(lib.qed_project_bwd_exchange(1, N, _lib.ptr(s.means), _lib.ptr(s.quats), _lib.ptr(s.scales), _lib.ptr(s.opacities), 0,
                                                _lib.ptr(s.sh), K, deg, _lib.ptr(s.viewmats[sl].contiguous()), _lib.ptr(s.Ks[sl].contiguous()),
                                                s.width, s.height, 0.3, 0, 1, _lib.ptr(f["radii"]), _lib.ptr(f["conics"]), None, _lib.ptr(packed),
                                                _lib.ptr(out["means"]), _lib.ptr(out["quats"]), _lib.ptr(out["scales"]), _lib.ptr(out["opacities"]),
                                                None, peers, 1, r, C, tag, _lib.current_stream()), "qed_project_bwd_exchange")
        for k in acc:
            acc[k] += out[k]
    v_sh = torch.full_like(ref["sh"], float("nan"))
    _lib.check(lib.qed_sh_grad_from_view_colors(C, N, K, deg, _lib.ptr(s.means), _lib.ptr(xch), tag, _lib.ptr(v_sh), _lib.current_stream()),
               "qed_sh_grad_from_view_colors")
    acc["sh"] = v_sh
    for k in ref:
        scale = float(ref[k].abs().mean()) + 1e-20
        assert_close_frac(acc[k], ref[k], 1e-4, 1e-5 * scale, 1e-3, f"exchange v_{k}")
    assert bool((v_sh[:, (deg + 1) ** 2:, :] == 0).all())  # unused coefficient rows are written as zeros
    # records of a previous step (other tag) are ignored: nothing visible -> zero gradient
    _lib.check(lib.qed_sh_grad_from_view_colors(C, N, K, deg, _lib.ptr(s.means), _lib.ptr(xch), tag + 1.0, _lib.ptr(v_sh), _lib.current_stream()),
               "qed_sh_grad_from_view_colors")
    assert bool((v_sh == 0).all())
