"""Camera / data side of the hot path, CUDA only, same names as the reference's functions.

* `get_viewmat(optimized_camera_to_world)` — /root/reference/qed_splatter/model.py:22-38 (called at model.py:246):
  nerfstudio camera-to-world -> gsplat world-to-camera.  One launch of `qed_viewmat_from_c2w` instead of six tiny torch
  kernels; rebind it with `qed_splatter.model.get_viewmat = qed_splatter_b200.get_viewmat`.
* `backproject_frame(...)`, `voxel_down_sample(...)`, `merge_pointclouds(...)`, `create_pointcloud_from_frames(...)` —
  the arithmetic of `qed-init-pc` (/root/reference/qed_splatter/create_init_pointcloud.py:148-196, :85-91, :100-145,
  :199-261), which the reference runs on the host through Open3D with every intermediate cloud written to disk as PLY.
  Here the frames' depth images go through `qed_backproject_depth` / `qed_voxel_downsample` and the pairwise tree merge
  stays on the device.  File handling (transforms.json, PLY, the on-disk cache, colourising from RGB) stays the
  reference's: these functions take arrays, not dataset paths.

No CPU fallback: CPU tensors raise.
"""
from __future__ import annotations

from typing import List, Optional, Sequence

import numpy as np
import torch
from torch import Tensor

from . import _lib
from ._lib import check, current_stream, ptr


def get_viewmat(optimized_camera_to_world: Tensor) -> Tensor:
    """model.py:22-38: c2w [C,3|4,4] (OpenGL camera axes) -> gsplat world-to-camera [C,4,4]."""
    c2w = optimized_camera_to_world
    _lib.require_cuda(c2w)
    _lib.require_dtype(c2w, (torch.float32,), "optimized_camera_to_world")
    if c2w.dim() != 3 or c2w.shape[1] not in (3, 4) or c2w.shape[2] != 4:
        raise ValueError("optimized_camera_to_world must be [C,3,4] or [C,4,4]")
    if c2w.requires_grad and torch.is_grad_enabled():
        # camera_optimizer.mode != "off": the projection has no pose gradient (rasterization() raises for it too)
        raise NotImplementedError("get_viewmat: no gradient with respect to the camera pose (camera_optimizer must be 'off')")
    c2w = c2w.detach().contiguous()
    C = c2w.shape[0]
    out = torch.empty(C, 4, 4, dtype=torch.float32, device=c2w.device)
    check(_lib.load().qed_viewmat_from_c2w(C, int(c2w.shape[1]), ptr(c2w), ptr(out), current_stream()), "qed_viewmat_from_c2w")
    return out


def _host_f32(x, shape, what: str) -> np.ndarray:
    a = np.ascontiguousarray(x.detach().cpu().numpy() if isinstance(x, Tensor) else np.asarray(x), dtype=np.float32)
    if a.shape != shape:
        raise ValueError(f"{what} must be {shape}, got {a.shape}")
    return a


def opengl_c2w_to_opencv_w2c(c2w_opengl) -> np.ndarray:
    """create_init_pointcloud.py:59-70 `_opengl_c2w_to_opencv_w2c` (host, float64 inverse, float32 result)."""
    c2w = np.array(c2w_opengl, dtype=np.float64).copy()
    c2w[:3, 1:3] *= -1
    return np.linalg.inv(c2w).astype(np.float32)


def backproject_frame(depth: Tensor, intrinsic, w2c, depth_unit_scale_factor: float = 0.001, depth_max: float = 100.0, stride: int = 1,
                      frame_voxel_size: Optional[float] = 0.05) -> Optional[Tensor]:
    """create_init_pointcloud.py:148-196: one depth frame -> world points [n,3] (float32, CUDA), or None when no pixel
    survives.  depth [H,W] float32 or the raw uint16 image; intrinsic [3,3]; w2c [4,4] OpenCV world-to-camera
    (`opengl_c2w_to_opencv_w2c` of the frame's transform_matrix)."""
    _lib.require_cuda(depth)
    if depth.dim() == 3:
        depth = depth[..., 0]
    if depth.dtype == torch.int16:
        depth = depth.view(torch.uint16)
    _lib.require_dtype(depth, (torch.float32, torch.uint16), "depth")
    depth = depth.contiguous()
    H, W = depth.shape
    K = _host_f32(intrinsic, (3, 3), "intrinsic")
    E = _host_f32(w2c, (4, 4), "w2c")
    lib = _lib.load()
    dev = depth.device
    cap = ((W + stride - 1) // stride) * ((H + stride - 1) // stride)
    points = torch.empty(cap, 3, dtype=torch.float32, device=dev)
    count = torch.zeros(1, dtype=torch.int64, device=dev)
    ws = torch.empty(max(int(lib.qed_backproject_workspace_bytes(W, H, stride)), 8), dtype=torch.uint8, device=dev)
    check(lib.qed_backproject_depth(W, H, ptr(depth), 1 if depth.dtype == torch.uint16 else 0, float(depth_unit_scale_factor), float(depth_max),
                                    int(stride), K.ctypes.data, E.ctypes.data, ptr(points), ptr(count), ptr(ws), ws.numel(), current_stream()),
          "qed_backproject_depth")
    if frame_voxel_size is not None and frame_voxel_size > 0:
        # the count stays on the device between the two passes; the only host read is the final size
        return _voxel_down_sample(points, float(frame_voxel_size), n_dev=count, empty_is_none=True)
    n = int(count.item())
    return points[:n] if n > 0 else None


def _voxel_down_sample(points: Tensor, voxel_size: float, n_dev: Optional[Tensor] = None, empty_is_none: bool = False):
    lib = _lib.load()
    n = points.shape[0]
    dev = points.device
    out = torch.empty(max(n, 1), 3, dtype=torch.float32, device=dev)
    n_out = torch.zeros(1, dtype=torch.int64, device=dev)
    ws = torch.empty(max(int(lib.qed_voxel_downsample_workspace_bytes(n)), 8), dtype=torch.uint8, device=dev)
    check(lib.qed_voxel_downsample(n, ptr(n_dev), ptr(points), float(voxel_size), ptr(out), ptr(n_out), ptr(ws), ws.numel(), current_stream()),
          "qed_voxel_downsample")
    m = int(n_out.item())
    if m == 0 and empty_is_none:
        return None
    return out[:m]


def voxel_down_sample(points: Tensor, voxel_size: float) -> Tensor:
    """Open3D `PointCloud.voxel_down_sample(voxel_size)` (create_init_pointcloud.py:89, :194, :260): mean of the points
    of every occupied voxel floor(p / voxel_size); rows sorted by voxel index."""
    _lib.require_cuda(points)
    _lib.require_dtype(points, (torch.float32,), "points")
    if points.dim() != 2 or points.shape[1] != 3:
        raise ValueError("points must be [n,3]")
    if not voxel_size > 0:
        raise ValueError("voxel_size must be positive")
    return _voxel_down_sample(points.contiguous(), float(voxel_size))


def merge_pointclouds(clouds: Sequence[Tensor], voxel_size: float = 0.03, max_points: int = 2_000_000) -> Tensor:
    """create_init_pointcloud.py:100-145 `tree_merge_pointclouds_on_disk` without the disk: pairwise merge level by level,
    a merged pair is voxel-downsampled only when it exceeds max_points (:85-91), an odd cloud is carried forward."""
    current: List[Tensor] = list(clouds)
    if not current:
        raise RuntimeError("No valid point clouds could be generated from the dataset.")  # :246
    while len(current) > 1:
        nxt: List[Tensor] = []
        for i in range(0, len(current), 2):
            if i + 1 < len(current):
                merged = torch.cat([current[i], current[i + 1]], dim=0)
                if merged.shape[0] > max_points:
                    merged = voxel_down_sample(merged, voxel_size)
                nxt.append(merged)
            else:
                nxt.append(current[i])
        current = nxt
    return current[0]


def create_pointcloud_from_frames(depths: Sequence[Tensor], intrinsics: Sequence, c2w_opengl: Sequence, depth_unit_scale_factor: float = 0.001,
                                  voxel_size: float = 0.05, merge_voxel_size: float = 0.03, frame_voxel_size: Optional[float] = 0.05,
                                  max_points: int = 2_000_000, depth_max: float = 100.0, stride: int = 1) -> Tensor:
    """create_init_pointcloud.py:199-261 `create_pointcloud_from_transforms` on arrays: back-project every frame (frames
    without a valid pixel are skipped, :169-171 / :238), tree-merge, final voxel_down_sample(voxel_size)."""
    clouds = []
    for depth, K, c2w in zip(depths, intrinsics, c2w_opengl):
        pcd = backproject_frame(depth, K, opengl_c2w_to_opencv_w2c(c2w), depth_unit_scale_factor, depth_max, stride, frame_voxel_size)
        if pcd is not None:
            clouds.append(pcd)
    merged = merge_pointclouds(clouds, voxel_size=merge_voxel_size, max_points=max_points)
    return voxel_down_sample(merged, voxel_size)
