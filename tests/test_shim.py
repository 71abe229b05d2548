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
