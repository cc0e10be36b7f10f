"""Consumers of `MT%TransferData` on the host side: sigma_R and the growth quantity
f*sigma_8 exactly as `Transfer_Get_sigmas` / `Transfer_Get_SigmaR` form them (fortran/results.f90:3950-4010,
4115-4182): a trapezoid in ln k over the transfer wavenumbers of (W(kR) k^2)^2 P_R(k) T_a(k) T_b(k).
These run on the host in the reference too; they are here so that the transfer outputs of loops A / A' can be
checked against the reference's published sigma_8 values."""
import numpy as np

TRANSFER_KH, TRANSFER_TOT, TRANSFER_NEWT_VEL_CDM = 0, 6, 10   # 0-based columns (results.f90:3432-3438)


def sigma_r(hs, transfer, radius=8.0, var1=TRANSFER_TOT, var2=TRANSFER_TOT, root=True):
    """`transfer`: (n_z, n_k, 13) TransferData over all of MT%q_trans (source k first, then the extra transfer k),
    k ascending.  Returns one value per redshift: sigma_R (root) or the integral itself."""
    tr = np.asarray(transfer, dtype=np.float64)
    kh = tr[0, :, TRANSFER_KH]
    keep = kh != 0
    tr, kh = tr[:, keep, :], kh[keep]
    h = hs.params.H0 / 100.0
    k = kh * h
    x = kh * radius
    win = np.where(x < 1e-2, 1.0 - x ** 2 / 10, 3 * (np.sin(x) - x * np.cos(x)) / np.where(x < 1e-2, 1.0, x) ** 3)
    lnk = np.log(k)
    dlnk = np.empty_like(lnk)
    dlnk[0] = 0.5
    dlnk[1:] = lnk[1:] - lnk[:-1]
    powers = (win * k ** 2) ** 2 * hs.scalar_power_at(k)
    d = powers[None, :] * tr[:, :, var1] * tr[:, :, var2]
    prev = np.concatenate([np.zeros((d.shape[0], 1)), d[:, :-1]], axis=1)
    sig = ((d + prev) * dlnk[None, :] / 2).sum(axis=1)
    return np.sqrt(sig) if root else sig


def sigma8_and_fsigma8(hs, transfer):
    """sigma_8(z) and f sigma_8(z) = sigma2_vdelta_8 / sigma_8 (camb/results.py get_fsigma8)."""
    s8 = sigma_r(hs, transfer, 8.0)
    sv = sigma_r(hs, transfer, 8.0, TRANSFER_TOT, TRANSFER_NEWT_VEL_CDM, root=False)
    return s8, sv / s8


def matter_power(hs, transfer, var1=TRANSFER_TOT, var2=TRANSFER_TOT, hubble_units=True):
    """`Transfer_GetUnsplinedPower` (fortran/results.f90:3472-3513): P(k) = 2 pi^2/k^3 T_1 T_2 P_R(k) at the transfer
    wavenumbers, in (Mpc/h)^3 when hubble_units.  Returns (kh [n_k], P [n_z, n_k]) in the order of `transfer`."""
    tr = np.asarray(transfer, dtype=np.float64)
    kh = tr[0, :, TRANSFER_KH]
    h = hs.params.H0 / 100.0
    k = kh * h
    pk = tr[:, :, var1] * tr[:, :, var2] * (k * np.pi * 2 * np.pi * hs.scalar_power_at(k))[None, :]
    return kh, pk * h ** 3 if hubble_units else pk


def matter_power_interp(hs, transfer, minkh=1e-4, maxkh=1.0, npoints=100):
    """`Transfer_GetMatterPowerData` + `Transfer_GetMatterPowerD` (fortran/results.f90:3559-3616, 3700-3790): natural cubic
    spline of log P in log(k/h), evaluated on a regular grid in log(k/h) -- what `get_matter_power_spectrum` returns."""
    from scipy.interpolate import CubicSpline
    kh, pk = matter_power(hs, transfer)
    lk = np.log(kh)
    out_kh = minkh * np.exp(np.arange(npoints) * (np.log(maxkh / minkh) / (npoints - 1)))
    res = np.empty((pk.shape[0], npoints))
    for i in range(pk.shape[0]):
        res[i] = np.exp(CubicSpline(lk, np.log(pk[i]), bc_type="natural")(np.log(out_kh)))
    return out_kh, res
