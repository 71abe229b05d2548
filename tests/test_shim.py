"""CPU: the import the reference performs (qed_splatter/model.py:7) resolves through the shim."""
import importlib
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_import_resolves_to_b200_rasterization():
    shim = os.path.join(ROOT, "qed_splatter_b200", "shim")
    sys.path.insert(0, shim)
    try:
        for m in [k for k in sys.modules if k == "gsplat" or k.startswith("gsplat.")]:
            del sys.modules[m]
        mod = importlib.import_module("gsplat.rendering")
        from qed_splatter_b200 import rasterization

        assert mod.rasterization is rasterization
        import inspect

        params = inspect.signature(mod.rasterization).parameters
        # every keyword the reference passes at qed_splatter/model.py:267-288
        for kw in ("means", "quats", "scales", "opacities", "colors", "viewmats", "Ks", "width", "height", "tile_size", "packed",
                   "near_plane", "far_plane", "render_mode", "sh_degree", "sparse_grad", "absgrad", "rasterize_mode"):
            assert kw in params, kw
    finally:
        sys.path.remove(shim)
        for m in [k for k in sys.modules if k == "gsplat" or k.startswith("gsplat.")]:
            del sys.modules[m]


def test_lazy_info_builds_gsplat_lists_on_first_access():
    """`rasterization()` composites on exact tile lists; gsplat's own lists in `info` are built on demand, once."""
    from qed_splatter_b200.rendering import LazyInfo

    calls = []

    def build():
        calls.append(1)
        return "ids", "flat", "off"

    info = LazyInfo({"radii": 1, "width": 5}, build)
    assert "isect_ids" in info and set(info.keys()) == {"radii", "width", "isect_ids", "flatten_ids", "isect_offsets"} and not calls
    assert info["radii"] == 1 and info.get("width") == 5 and info.get("nope", 7) == 7 and not calls
    assert info["flatten_ids"] == "flat" and info["isect_ids"] == "ids" and info.get("isect_offsets") == "off" and calls == [1]
    assert dict(LazyInfo({"radii": 1}, build))["isect_ids"] == "ids"
    assert {**LazyInfo({"radii": 1}, build)}["isect_offsets"] == "off"
    assert "flat" in list(LazyInfo({"radii": 1}, build).values())
    assert LazyInfo({"radii": 1}, build).copy()["flatten_ids"] == "flat"


def test_packed_record_views_are_recognised():
    """Host logic of the autograd path: the compositor backward returns strided views of ONE packed [C*N,12] record and
    the projection backward hands that record to the kernel only if its four incoming gradients are exactly those views."""
    import torch

    from qed_splatter_b200 import ops

    C, N, D = 2, 7, 4
    packed = torch.zeros(C * N, ops.GRAD_FLOATS)
    P = packed.view(C, N, ops.GRAD_FLOATS)
    views = (P[..., 0:2], P[..., 4:7], P[..., 8:8 + D], P[..., 7])
    assert ops._packed_record_of(*views, C, N, D) == packed.data_ptr()
    v1 = (P[..., 0:2], P[..., 4:7], P[..., 8:9], P[..., 7])  # one channel ("D" / "ED" renders)
    assert ops._packed_record_of(*v1, C, N, 1) == packed.data_ptr()
    # anything autograd summed, copied or re-laid-out is NOT the record: fall back to the one-by-one arguments
    assert ops._packed_record_of(views[0].clone(), *views[1:], C, N, D) is None
    assert ops._packed_record_of(views[0], views[1].contiguous(), views[2], views[3], C, N, D) is None
    assert ops._packed_record_of(views[0], views[1], views[2], None, C, N, D) is None
    assert ops._packed_record_of(P[..., 2:4], views[1], views[2], views[3], C, N, D) is None  # wrong slot
    other = torch.zeros(C * N, ops.GRAD_FLOATS).view(C, N, ops.GRAD_FLOATS)
    assert ops._packed_record_of(views[0], other[..., 4:7], views[2], views[3], C, N, D) is None  # another record
    assert ops._packed_record_of(*views, C, N, 3) is None  # channel count of the render differs
    assert ops._packed_record_of(views[0].double(), *views[1:], C, N, D) is None
