"""GPU: camera / data-side kernels (rows a1, f4) through the C-ABI against oracle/pointcloud.py.

Integer work (which pixels survive, voxel membership, counts) is exact; the float work mirrors the oracle's pinned
float32 op order, so points are compared bit for bit where the op order is pinned (get_viewmat, back-projection) and to
1e-6 relative where the oracle accumulates in float64 through numpy (voxel means)."""
import numpy as np
import pytest
import torch

from oracle import pointcloud as opc
from test_data_side_cpu import _frame, _rand_c2w

pytestmark = pytest.mark.gpu


def test_get_viewmat_bit_exact(cuda):
    from qed_splatter_b200 import get_viewmat

    for rows in (3, 4):
        c2w = _rand_c2w(9, seed=11, rows=rows)
        got = get_viewmat(torch.from_numpy(c2w).to(cuda)).cpu().numpy()
        assert np.array_equal(got, opc.get_viewmat_pinned(c2w))
    import oracle

    s_c2w = torch.from_numpy(_rand_c2w(3, seed=12))
    torch.testing.assert_close(get_viewmat(s_c2w.to(cuda)).cpu(), oracle.get_viewmat(s_c2w), rtol=0, atol=2e-6)
    assert get_viewmat(torch.zeros(0, 3, 4, device=cuda)).shape == (0, 4, 4)
    with pytest.raises(RuntimeError):
        get_viewmat(s_c2w)  # CPU tensor: no fallback
    with pytest.raises(NotImplementedError):
        get_viewmat(s_c2w.to(cuda).requires_grad_(True))


@pytest.mark.parametrize("stride,as_u16", [(1, True), (3, True), (2, False)])
def test_backproject_bit_exact(cuda, stride, as_u16):
    from qed_splatter_b200 import data_side

    depth, K, c2w = _frame(21, H=71, W=93)  # ragged against the scan tile and the stride
    w2c = opc.opengl_c2w_to_opencv_w2c(c2w)
    if as_u16:
        d_t = torch.from_numpy(depth.astype(np.int16)).view(torch.uint16).to(cuda)
        d_ref = depth
    else:
        d_ref = depth.astype(np.float32)
        d_ref[0, 0], d_ref[1, 0], d_ref[2, 0] = np.nan, np.inf, -5.0
        d_t = torch.from_numpy(d_ref).to(cuda)
    got = data_side.backproject_frame(d_t, K, w2c, 0.001, depth_max=5.0, stride=stride, frame_voxel_size=None)
    ref = opc.backproject_depth(d_ref, K, w2c, 0.001, 5.0, stride)
    assert got.shape == ref.shape
    assert np.array_equal(got.cpu().numpy(), ref)  # same pixels, same order, same float32 op order


def test_backproject_empty_frame(cuda):
    from qed_splatter_b200 import data_side

    z = torch.zeros(20, 30, device=cuda)
    assert data_side.backproject_frame(z, np.eye(3, dtype=np.float32), np.eye(4, dtype=np.float32), 1.0, 10.0, 1, None) is None
    assert data_side.backproject_frame(z, np.eye(3, dtype=np.float32), np.eye(4, dtype=np.float32), 1.0, 10.0, 1, 0.05) is None


@pytest.mark.parametrize("n,vs", [(1, 0.05), (5000, 0.05), (200_000, 0.03), (70_001, 0.5)])
def test_voxel_down_sample_matches_oracle(cuda, n, vs):
    from qed_splatter_b200 import data_side

    g = np.random.default_rng(n)
    pts = g.normal(scale=0.6, size=(n, 3)).astype(np.float32)
    pts[: n // 10] = pts[n // 10: 2 * (n // 10)][: n // 10]  # exact duplicates
    got = data_side.voxel_down_sample(torch.from_numpy(pts).to(cuda), vs).cpu().numpy()
    ref = opc.voxel_down_sample(pts, vs)
    assert got.shape == ref.shape  # number of occupied voxels: exact
    np.testing.assert_allclose(got, ref, rtol=1e-6, atol=1e-7)  # both sorted by voxel index
    assert data_side.voxel_down_sample(torch.zeros(0, 3, device=cuda), vs).shape == (0, 3)


def test_pipeline_matches_oracle(cuda):
    """create_init_pointcloud.py:199-261 on arrays: per-frame back-projection + frame voxel merge (device-side count in
    between), pairwise tree merge with the max_points rule, final voxel merge."""
    from qed_splatter_b200 import data_side

    frames = [_frame(s, H=60, W=80) for s in range(5)]
    kw = dict(depth_unit_scale_factor=0.001, voxel_size=0.2, merge_voxel_size=0.1, frame_voxel_size=0.15, max_points=1500, depth_max=5.0, stride=2)
    ref = opc.create_pointcloud([f[0] for f in frames], [f[1] for f in frames], [f[2] for f in frames], **kw)
    got = data_side.create_pointcloud_from_frames([torch.from_numpy(f[0].astype(np.int16)).view(torch.uint16).to(cuda) for f in frames],
                                                  [f[1] for f in frames], [f[2] for f in frames], **kw).cpu().numpy()
    assert got.shape == ref.shape
    np.testing.assert_allclose(got, ref, rtol=2e-6, atol=2e-7)
