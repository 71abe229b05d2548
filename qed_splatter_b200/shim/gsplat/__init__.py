"""Import shim: put `qed_splatter_b200/shim` on PYTHONPATH and the reference's
`from gsplat.rendering import rasterization` (/root/reference/qed_splatter/model.py:7) resolves to the
B200-native implementation with no change to the reference.  Only the symbol the reference imports is
provided; this is NOT a gsplat re-implementation (nerfstudio's own imports of gsplat, e.g.
`gsplat.strategy.DefaultStrategy`, still need the real package — see INTEGRATION.md for the one-line
monkeypatch that is used when gsplat itself is installed)."""
from . import rendering  # noqa: F401

__version__ = "1.4.0+qed_splatter_b200"
