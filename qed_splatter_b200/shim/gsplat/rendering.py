"""`gsplat.rendering.rasterization` -> qed_splatter_b200.rasterization (same keyword surface, same returns)."""
from qed_splatter_b200.rendering import rasterization  # noqa: F401

__all__ = ["rasterization"]
