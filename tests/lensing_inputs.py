"""Analytic CMB-like spectra for the lensing parity tests and the golden fixture (no data files needed)."""
import numpy as np

from camb_b200.workloads import analytic_cmb_cls as analytic_cls  # noqa: F401  (the fixture was made from this function)


def cl_scalar_from_analytic(cls):
    """CAMB's Cl_scalar convention: C_Phi column holds l^4 C_phi (CPP = Cl_scalar(:, C_Phi) (l+1)^2 / l^2 / 2pi)."""
    out = cls.copy()
    l = np.arange(len(cls), dtype=float)
    out[2:, 3] = cls[2:, 3] * 2 * np.pi * l[2:] ** 2 / (l[2:] + 1) ** 2
    return out
