"""Shared helpers for the parity tests."""
import torch


def mismatch_fraction(a: torch.Tensor, b: torch.Tensor, rtol: float, atol: float) -> float:
    """Fraction of elements outside |a-b| <= atol + rtol*|b|."""
    a = a.detach().double().cpu()
    b = b.detach().double().cpu()
    bad = (a - b).abs() > atol + rtol * b.abs()
    return float(bad.double().mean()) if bad.numel() else 0.0


def assert_close_frac(a, b, rtol, atol, max_frac, what=""):
    """allclose with a bounded mismatch fraction: threshold tests (alpha < 1/255, T <= 1e-4) flip on 1-ulp
    differences between exp and ex2.approx and move a pixel by up to ~4e-3 (SURVEY.md §7 hard part a)."""
    frac = mismatch_fraction(a, b, rtol, atol)
    err = float((a.detach().double().cpu() - b.detach().double().cpu()).abs().max()) if a.numel() else 0.0
    assert frac <= max_frac, f"{what}: {frac:.3e} of elements outside rtol={rtol} atol={atol} (max abs err {err:.3e})"
    return frac, err


def scene_args(s, device=None):
    d = dict(means=s.means, quats=s.quats, scales=s.scales, opacities=s.opacities, colors=s.sh, viewmats=s.viewmats, Ks=s.Ks)
    if device is not None:
        d = {k: v.to(device) for k, v in d.items()}
    return d
