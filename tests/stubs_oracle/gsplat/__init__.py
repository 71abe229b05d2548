"""Test-only `gsplat` whose rasterization is the CPU oracle (see tests/stubs/README.md)."""
