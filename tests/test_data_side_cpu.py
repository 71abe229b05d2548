"""CPU: the camera / data-side oracle (oracle/pointcloud.py).

What the reference itself writes is pinned by EXECUTING its lines: `get_viewmat` (model.py:22-38) through the reference
model harness, `_opengl_c2w_to_opencv_w2c` / `_frame_intrinsics` (create_init_pointcloud.py:50-70) extracted from the
reference source (the module itself cannot be imported: open3d / tyro are not in this image).  The two Open3D calls are
restated from their published algorithm (parity unpinned, oracle/pointcloud.py) and checked through their defining
properties."""
import ast
import os

import numpy as np
import pytest
import torch

import reference_model as rm
from oracle import pointcloud as opc


def _rand_c2w(C, seed=0, rows=3):
    g = np.random.default_rng(seed)
    out = np.zeros((C, rows, 4), dtype=np.float32)
    for c in range(C):
        q, _ = np.linalg.qr(g.normal(size=(3, 3)))
        if np.linalg.det(q) < 0:
            q[:, 0] *= -1
        out[c, :3, :3] = q
        out[c, :3, 3] = g.uniform(-5, 5, size=3)
        if rows == 4:
            out[c, 3, 3] = 1.0
    return out


def _reference_function(name):
    """A pure-numpy helper of create_init_pointcloud.py, compiled from the reference's own source text."""
    root = rm.reference_root()
    if root is None:
        pytest.skip("reference package not available")
    src = open(os.path.join(root, "qed_splatter", "create_init_pointcloud.py")).read()
    tree = ast.parse(src)
    fn = [n for n in tree.body if isinstance(n, ast.FunctionDef) and n.name == name]
    assert fn, name
    ns = {"np": np}
    exec(compile(ast.Module(body=fn, type_ignores=[]), f"<reference:{name}>", "exec"), ns)
    return ns[name]


def test_get_viewmat_pinned_equals_reference_lines():
    c2w = _rand_c2w(7, seed=1)
    import oracle

    pinned = opc.get_viewmat_pinned(c2w)
    ours = oracle.get_viewmat(torch.from_numpy(c2w)).numpy()
    np.testing.assert_allclose(pinned, ours, rtol=0, atol=2e-6)
    assert np.array_equal(pinned[:, :3, :3], ours[:, :3, :3]) and np.array_equal(pinned[:, 3], ours[:, 3])
    # inverse of the flipped pose
    flipped = np.concatenate([c2w, np.tile(np.array([[[0, 0, 0, 1]]], np.float32), (7, 1, 1))], axis=1).astype(np.float64)
    flipped[:, :3, 1:3] *= -1
    np.testing.assert_allclose(pinned.astype(np.float64) @ flipped, np.tile(np.eye(4), (7, 1, 1)), atol=5e-6)
    if rm.reference_root() is not None:
        with rm.reference_modules("oracle") as mod:
            ref = mod.get_viewmat(torch.from_numpy(c2w)).numpy()
        np.testing.assert_allclose(pinned, ref, rtol=0, atol=2e-6)
        assert np.array_equal(pinned[:, :3, :3], ref[:, :3, :3])


def test_axis_flip_and_intrinsics_equal_reference_lines():
    ref_flip = _reference_function("_opengl_c2w_to_opencv_w2c")
    ref_K = _reference_function("_frame_intrinsics")
    for c2w in _rand_c2w(4, seed=2, rows=4):
        assert np.array_equal(opc.opengl_c2w_to_opencv_w2c(c2w), ref_flip(c2w.astype(np.float64)))
    K = ref_K({"fl_x": 500.0, "cx": 320.0, "cy": 240.0}, {"fl_y": 510.0})
    assert K.dtype == np.float32 and K[0, 0] == 500.0 and K[1, 1] == 510.0 and K[0, 2] == 320.0 and K[1, 2] == 240.0


def _frame(seed=0, H=48, W=64):
    g = np.random.default_rng(seed)
    depth = g.uniform(400, 6000, size=(H, W)).astype(np.uint16)
    depth[g.uniform(size=(H, W)) < 0.2] = 0  # sensor holes
    K = np.array([[60.0, 0, W / 2 - 0.5], [0, 62.0, H / 2 + 0.25], [0, 0, 1]], dtype=np.float32)
    c2w = _rand_c2w(1, seed=seed + 100, rows=4)[0]
    return depth, K, c2w


def test_backproject_reprojects_onto_its_pixels():
    depth, K, c2w = _frame(3)
    w2c = opc.opengl_c2w_to_opencv_w2c(c2w)
    for stride in (1, 3):
        pts = opc.backproject_depth(depth, K, w2c, 0.001, depth_max=5.0, stride=stride)
        d = depth[::stride, ::stride].astype(np.float32) * np.float32(0.001)
        keep = (d > 0) & (d < 5.0)
        assert pts.shape == (int(keep.sum()), 3)
        cam = (w2c[:3, :3].astype(np.float64) @ pts.T.astype(np.float64)).T + w2c[:3, 3]
        u = cam[:, 0] / cam[:, 2] * K[0, 0] + K[0, 2]
        v = cam[:, 1] / cam[:, 2] * K[1, 1] + K[1, 2]
        vs, us = np.nonzero(keep)
        np.testing.assert_allclose(u, us * stride, atol=2e-3)
        np.testing.assert_allclose(v, vs * stride, atol=2e-3)
        np.testing.assert_allclose(cam[:, 2], d[keep], rtol=2e-5)
    # non-finite and non-positive depths are dropped (create_init_pointcloud.py:166-167)
    bad = depth.astype(np.float32)
    bad[0, 0], bad[0, 1], bad[0, 2] = np.nan, np.inf, -3.0
    n_ref = opc.backproject_depth(np.where(np.isfinite(bad) & (bad > 0), bad, 0), K, w2c, 0.001, 100.0).shape[0]
    assert opc.backproject_depth(bad, K, w2c, 0.001, 100.0).shape[0] == n_ref


def test_voxel_down_sample_properties():
    g = np.random.default_rng(5)
    pts = g.normal(scale=0.4, size=(5000, 3)).astype(np.float32)
    vs = 0.05
    out = opc.voxel_down_sample(pts, vs)
    keys = opc.voxel_keys(pts, vs)
    uniq, cnt = np.unique(keys, axis=0, return_counts=True)
    assert out.shape[0] == uniq.shape[0]
    # the mean of a voxel's points lies in (the closure of) that voxel, rows come in voxel-index order
    assert np.all(np.abs(out / vs - (uniq + 0.5)) <= 0.5 + 1e-4)
    # the count-weighted mean of the voxel means is the mean of the points
    np.testing.assert_allclose((out.astype(np.float64) * cnt[:, None]).sum(0) / cnt.sum(), pts.astype(np.float64).mean(0), atol=1e-6)
    assert opc.voxel_down_sample(pts[:0], vs).shape == (0, 3)
    one = opc.voxel_down_sample(np.tile(np.float32([[0.31, -0.22, 0.05]]), (9, 1)), vs)
    np.testing.assert_allclose(one, [[0.31, -0.22, 0.05]], rtol=1e-6)


def test_tree_merge_schedule_and_pipeline():
    g = np.random.default_rng(7)
    clouds = [g.normal(scale=0.3, size=(n, 3)).astype(np.float32) for n in (300, 200, 250, 100, 50)]
    plain = opc.tree_merge(clouds, voxel_size=0.03, max_points=10**9)
    assert plain.shape[0] == 900  # nothing exceeds max_points: pure concatenation, level by level
    assert np.array_equal(plain[:500], np.concatenate(clouds[:2])) and np.array_equal(plain[-50:], clouds[4])
    small = opc.tree_merge(clouds, voxel_size=0.1, max_points=400)  # pairs above 400 points are merged by voxel (:85-91)
    assert small.shape[0] < 900
    with pytest.raises(RuntimeError):
        opc.tree_merge([])
    frames = [_frame(s) for s in range(3)]
    pc = opc.create_pointcloud([f[0] for f in frames], [f[1] for f in frames], [f[2] for f in frames], 0.001, voxel_size=0.2,
                               merge_voxel_size=0.1, frame_voxel_size=0.2, max_points=1000, depth_max=5.0, stride=2)
    assert pc.ndim == 2 and pc.shape[1] == 3 and pc.shape[0] > 10
    assert np.unique(opc.voxel_keys(pc, 0.2), axis=0).shape[0] >= int(0.95 * pc.shape[0])  # (a mean may sit on a voxel face)
