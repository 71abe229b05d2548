"""Stand-in for torchmetrics (test infrastructure, see tests/stubs/README.md): /root/reference/qed_splatter/metrics.py
imports three image metrics at module scope; the hot-path tests never evaluate them (metrics.py is out of scope,
SURVEY.md section 2a #2)."""
