"""CPU oracle for the depth-supervised splat render/train hot path.

TEST INFRASTRUCTURE ONLY.  Nothing in the product package (`qed_splatter_b200`)
may import this.  Only `tests/`, `__graft_entry__.smoke()` and the
`cpu_baseline` / `--impl reference` legs of `bench.py` use it, and only as the
checker / the CPU arm, never as the thing shipped.

PARITY UNPINNED: the reference (`/root/reference`, leggedrobotics/qed-splatter)
holds none of this arithmetic.  It calls `gsplat.rendering.rasterization`
(`qed_splatter/model.py:267-288`); gsplat is an un-vendored, un-pinned
third-party dependency (effective version 1.4.0 via nerfstudio 1.1.5, see
SURVEY.md §0.2) that is absent from this image and cannot be installed (no
network).  The reference ships no tests, fixtures or golden vectors.  The
oracle therefore restates gsplat 1.4.0's published reference algorithm
(`gsplat/cuda/_torch_impl.py`: `_quat_scale_to_covar_preci`, `_world_to_cam`,
`_persp_proj`, `_fully_fused_projection`, `_spherical_harmonics`,
`_isect_tiles`, `_isect_offset_encode`, `_rasterize_to_pixels`) plus the
reference's own call-site arithmetic (`qed_splatter/model.py:22-38`,
`:295-306`, `:87-116`).  It is self-checked instead (tests/test_oracle_*.py):
float64 autograd vs finite differences, the explicit compositing backward vs
autograd, SH-basis orthonormality, and compositing invariants.
"""
from . import pointcloud, torch_impl  # noqa: F401
from .torch_impl import *  # noqa: F401,F403
