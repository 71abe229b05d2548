from oracle import rasterization  # noqa: F401  (same keyword surface as the call at qed_splatter/model.py:267-288)
