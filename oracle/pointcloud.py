"""CPU oracle of the camera / data side rows.  TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

  get_viewmat_pinned    <- /root/reference/qed_splatter/model.py:22-38 with the float32 op order written out
                           (the reference's torch.bmm leaves it to the BLAS; pinned against the reference's own function
                           executed by tests/test_reference_model_cpu.py to 1 ulp, and against oracle.get_viewmat)
  backproject_depth     <- /root/reference/qed_splatter/create_init_pointcloud.py:165-185 (depth cleaning + Open3D
                           `PointCloud.create_from_depth_image(depth, intrinsic, extrinsic, depth_scale=1, depth_max, stride)`)
  voxel_down_sample     <- Open3D `PointCloud.voxel_down_sample(voxel_size)` (create_init_pointcloud.py:89, :194, :260)
  tree_merge            <- create_init_pointcloud.py:100-145 without the disk round trips
  create_pointcloud     <- create_init_pointcloud.py:199-261 on arrays

PARITY UNPINNED for the two Open3D calls: Open3D 0.18 is an un-vendored dependency of the reference (pyproject: `open3d`,
unpinned), absent from this image.  Restated from its published algorithm: `create_from_depth_image` keeps pixel (u, v) of
the strided grid iff 0 < d < depth_max, d = depth / depth_scale, un-projects it with the pinhole intrinsics and maps it
through inverse(extrinsic); `voxel_down_sample` bins points by floor(p / voxel_size) and returns the mean of every occupied
voxel.  Point order is unspecified in Open3D (atomic counters / hash map); here it is pixel order and voxel-key order, and
parity is defined on the SET of points (tests sort both sides).  Everything the reference itself writes around those calls
(depth cleaning :165-168, axis flip + inverse :59-70, merge schedule :100-145, the max_points rule :85-91) is followed
line by line and IS pinned by tests/test_data_side_cpu.py executing the reference's own helper functions.
"""
from __future__ import annotations

from typing import List, Optional, Sequence

import numpy as np

f32 = np.float32


def get_viewmat_pinned(c2w: np.ndarray) -> np.ndarray:
    """model.py:22-38 in float32 with the sum order ((r0 t0 + r1 t1) + r2 t2), no FMA."""
    c2w = np.asarray(c2w, dtype=f32)
    C = c2w.shape[0]
    R = c2w[:, :3, :3] * np.array([[[1, -1, -1]]], dtype=f32)
    T = c2w[:, :3, 3]
    out = np.zeros((C, 4, 4), dtype=f32)
    out[:, 3, 3] = 1.0
    Rinv = np.transpose(R, (0, 2, 1))
    out[:, :3, :3] = Rinv
    acc = (Rinv[:, :, 0] * T[:, 0:1]).astype(f32)
    acc = (acc + (Rinv[:, :, 1] * T[:, 1:2]).astype(f32)).astype(f32)
    acc = (acc + (Rinv[:, :, 2] * T[:, 2:3]).astype(f32)).astype(f32)
    out[:, :3, 3] = -acc
    return out


def opengl_c2w_to_opencv_w2c(c2w_opengl) -> np.ndarray:
    """create_init_pointcloud.py:59-70."""
    c2w = np.array(c2w_opengl, dtype=np.float64).copy()
    c2w[:3, 1:3] *= -1
    return np.linalg.inv(c2w).astype(f32)


def invert_extrinsic(E: np.ndarray) -> np.ndarray:
    """inverse(extrinsic) as [3,4] float32 (cofactor inverse of the 3x3 block in float64, like the library's host code)."""
    E = np.asarray(E, dtype=np.float64)
    Ri = np.linalg.inv(E[:3, :3])
    P = np.zeros((3, 4), dtype=np.float64)
    P[:, :3] = Ri
    P[:, 3] = -(Ri @ E[:3, 3])
    return P.astype(f32)


def backproject_depth(depth_raw: np.ndarray, intrinsic, w2c, depth_unit_scale_factor: float, depth_max: float, stride: int = 1) -> np.ndarray:
    """create_init_pointcloud.py:165-185: world points [n,3] float32 in pixel (row-major) order."""
    depth = np.asarray(depth_raw).astype(f32) * depth_unit_scale_factor  # :165, float32 product (numpy keeps float32)
    depth = depth.astype(f32)
    depth[~np.isfinite(depth)] = 0.0  # :166
    depth[depth <= 0.0] = 0.0  # :167
    K = np.asarray(intrinsic, dtype=f32)
    fx, fy, cx, cy = K[0, 0], K[1, 1], K[0, 2], K[1, 2]
    P = invert_extrinsic(w2c)
    H, W = depth.shape
    vs, us = np.meshgrid(np.arange(0, H, stride), np.arange(0, W, stride), indexing="ij")
    d = depth[vs, us]
    keep = (d > 0) & (d < f32(depth_max))
    u = us[keep].astype(f32)
    v = vs[keep].astype(f32)
    d = d[keep]
    xc = (((u - cx).astype(f32) * d).astype(f32) / fx).astype(f32)
    yc = (((v - cy).astype(f32) * d).astype(f32) / fy).astype(f32)
    out = np.empty((d.shape[0], 3), dtype=f32)
    for r in range(3):
        acc = (P[r, 0] * xc).astype(f32)
        acc = (acc + (P[r, 1] * yc).astype(f32)).astype(f32)
        acc = (acc + (P[r, 2] * d).astype(f32)).astype(f32)
        out[:, r] = (acc + P[r, 3]).astype(f32)
    return out


def voxel_keys(points: np.ndarray, voxel_size: float) -> np.ndarray:
    return np.floor((np.asarray(points, dtype=f32) / f32(voxel_size)).astype(f32)).astype(np.int64)


def voxel_down_sample(points: np.ndarray, voxel_size: float) -> np.ndarray:
    """Mean of the points of every occupied voxel floor(p / voxel_size); rows sorted by voxel index (x, y, z)."""
    points = np.asarray(points, dtype=f32)
    if points.shape[0] == 0:
        return points.reshape(0, 3)
    keys = voxel_keys(points, voxel_size)
    uniq, inv = np.unique(keys, axis=0, return_inverse=True)
    inv = inv.reshape(-1)
    sums = np.zeros((uniq.shape[0], 3), dtype=np.float64)
    np.add.at(sums, inv, points.astype(np.float64))
    cnt = np.bincount(inv, minlength=uniq.shape[0]).astype(np.float64)
    return (sums / cnt[:, None]).astype(f32)


def tree_merge(clouds: Sequence[np.ndarray], voxel_size: float = 0.03, max_points: int = 2_000_000) -> np.ndarray:
    """create_init_pointcloud.py:100-145 (pairwise levels; :85-91 downsample only above max_points; odd one carried)."""
    current: List[np.ndarray] = list(clouds)
    if not current:
        raise RuntimeError("No valid point clouds could be generated from the dataset.")
    while len(current) > 1:
        nxt = []
        for i in range(0, len(current), 2):
            if i + 1 < len(current):
                merged = np.concatenate([current[i], current[i + 1]], axis=0)
                if merged.shape[0] > max_points:
                    merged = voxel_down_sample(merged, voxel_size)
                nxt.append(merged)
            else:
                nxt.append(current[i])
        current = nxt
    return current[0]


def create_pointcloud(depths, intrinsics, c2w_opengl, depth_unit_scale_factor: float = 0.001, voxel_size: float = 0.05,
                      merge_voxel_size: float = 0.03, frame_voxel_size: Optional[float] = 0.05, max_points: int = 2_000_000,
                      depth_max: float = 100.0, stride: int = 1) -> np.ndarray:
    """create_init_pointcloud.py:199-261 on arrays."""
    clouds = []
    for depth, K, c2w in zip(depths, intrinsics, c2w_opengl):
        pts = backproject_depth(depth, K, opengl_c2w_to_opencv_w2c(c2w), depth_unit_scale_factor, depth_max, stride)
        if pts.shape[0] == 0:
            continue
        if frame_voxel_size is not None and frame_voxel_size > 0:
            pts = voxel_down_sample(pts, frame_voxel_size)
        clouds.append(pts)
    return voxel_down_sample(tree_merge(clouds, merge_voxel_size, max_points), voxel_size)
