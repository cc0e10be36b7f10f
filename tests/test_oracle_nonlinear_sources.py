"""MakeNonlinearSources (cmbmain.f90:1190-1247): the oracle restatement against an independent numpy / scipy statement of the
same operation -- per wavenumber above Min_kh_nonlinear, a NATURAL cubic spline of the non-linear ratios over the transfer
output times (forutils cubic_spline_second_derivs with both ends omitted), evaluated at the time steps from the first
transfer time to the last-but-one step, multiplying sources 3.. only."""
import numpy as np
from scipy.interpolate import CubicSpline

import oracle_lib as O
from camb_b200.workloads import planck2018_params
from conftest import host_state


def nl_params():
    """planck_2018 cosmology with lensing and a ladder of transfer output redshifts down to z = 0 (what
    Transfer_SetForNonlinearLensing, model.f90, asks for when do_nonlinear >= 2)."""
    p = planck2018_params(lmax=1200, lensing=True, massless=True, want_transfer=True)
    zs = [10.0, 6.0, 4.0, 2.5, 1.5, 0.8, 0.4, 0.15, 0.0]
    p.transfer_PK_num_redshifts = len(zs)
    for i, z in enumerate(zs):
        p.transfer_PK_redshifts[i] = z
    p.transfer_kmax = 2.0
    return p


def synthetic_ratio(hs, q):
    """A smooth stand-in for halofit's sqrt(P_nl / P_lin): grows with k and towards late times, exactly 1 at low k."""
    m = hs.model
    tt = np.ctypeslib.as_array(m.transfer_times, shape=(m.num_transfer_redshifts,)).copy()
    growth = (tt / tt[-1]) ** 1.5
    kh = q / (m.H0 / 100)
    boost = 0.6 * np.tanh(np.maximum(kh - 0.05, 0.0) * 3.0) ** 2
    return tt, 1.0 + growth[:, None] * boost[None, :]          # (n_z, n_k)


def test_oracle_matches_independent_statement():
    hs = host_state("nl_sources", nl_params)
    q = np.asarray(hs.source_q())
    m = hs.model
    assert m.n_src == 3 and m.num_transfer_redshifts == 9
    rng = np.random.default_rng(5)
    src = rng.normal(size=(m.n_tau, m.n_src, len(q)))
    tt, ratio = synthetic_ratio(hs, q)
    out = O.make_nonlinear_sources(hs, q, src, ratio, min_kh=0.005)
    tau = np.ctypeslib.as_array(m.tau, shape=(m.n_tau,)).copy()
    assert abs(tt[-1] - m.tau0) < 1e-6 * m.tau0
    steps = np.nonzero(tau >= tt[0])[0]
    steps = steps[steps < m.n_tau - 1]                      # do i = first_step, npoints - 1
    exp = src.copy()
    kh = q / (m.H0 / 100)
    n_scaled = 0
    for ik in range(len(q)):
        if not kh[ik] > 0.005 or np.all(np.abs(ratio[:, ik] - 1) < 5e-4):
            continue
        n_scaled += 1
        exp[steps, 2, ik] *= CubicSpline(tt, ratio[:, ik], bc_type="natural")(tau[steps])
    assert 20 < n_scaled < len(q)                           # both branches of the k cut are exercised
    assert np.array_equal(out[:, :2, :], src[:, :2, :])     # T and E are never scaled
    assert np.array_equal(out[:steps[0]], src[:steps[0]]) and np.array_equal(out[-1], src[-1])
    assert np.max(np.abs(out - exp)) < 1e-12 * np.max(np.abs(src))
    # ratios of one are a no-op, bit for bit
    assert np.array_equal(O.make_nonlinear_sources(hs, q, src, np.ones_like(ratio)), src)
