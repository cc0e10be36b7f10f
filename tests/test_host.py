"""Host-side logic of the stand-in for the reference's Fortran host: k / l / tau grids and
thermo tables have the reference's structural properties (cmbmain.f90:864-936, 1267-1349;
results.f90:1505-1704, 2658-2779)."""
import numpy as np

from camb_b200.host import CAMBError, HostState, default_params
from camb_b200.workloads import params_ini_config1, planck2018_params, synthetic_lcdm_batch
from conftest import host_state


def test_grids_are_ordered_and_consistent():
    hs = host_state("c1", params_ini_config1)
    q = hs.source_q()
    assert np.all(np.diff(q) > 0) and np.all(np.diff(hs.q_int) > 0) and np.all(np.diff(hs.tau) > 0)
    assert hs.q_int[-1] <= q[-1] * (1 + 1e-12) and hs.n_src == 2
    m = hs.model
    # TimeSteps regions tile [tau_lowest, tau_highest] and dpoints are the trapezoid weights
    assert abs(hs.dtau.sum() - (hs.tau[-1] - hs.tau[0])) < 1e-6 * hs.tau[-1]
    assert m.tau0 > hs.tau[-1] * 0.99 and m.taurst < m.tau_maxvis < m.taurend
    md = hs.modes(False)
    ts = np.ctypeslib.as_array(md.tight_switchoff, shape=(md.n_k,))
    assert np.all(ts > 0) and np.all(ts <= m.tight_tau * (1 + 1e-12))


def test_thermo_tables_are_physical():
    hs = host_state("c1", params_ini_config1)
    m = hs.model
    th = np.ctypeslib.as_array(m.thermo_horner, shape=(m.nthermo, 16))
    emmu = np.ctypeslib.as_array(m.emmu, shape=(m.nthermo,))
    assert np.all(np.diff(emmu) >= -1e-12) and 0 <= emmu[0] <= 1e-30 and abs(emmu[-1] - 1) < 1e-6
    assert np.all(th[:-1, 4] > 0)                    # opacity
    a_over_tau = th[:-1, 8]
    tau = m.tauminn * np.exp(np.arange(m.nthermo - 1) * m.dlntau)
    assert np.all(np.diff(a_over_tau * tau) > 0)     # a(tau) monotone
    assert 1e-5 < hs.x_e(1.0) < 1e-3 and 0.05 < hs.x_e(1 / 1101.0) < 0.35 and hs.x_e(1 / 3000.0) > 0.99  # recombination-only x_e


def test_synthetic_batch_is_reproducible_and_valid():
    a = synthetic_lcdm_batch(3, seed=7)
    b = synthetic_lcdm_batch(3, seed=7)
    assert [p.H0 for p in a] == [p.H0 for p in b]
    hs = HostState(a[0])
    assert len(hs.ls) == 89 and hs.n_src == 3
    ext = synthetic_lcdm_batch(2, seed=7, extended=True)
    hs = HostState(ext[0])
    assert hs.model.Num_Nu_massive == 1 and hs.model.de_is_cosmological_constant == 0


def test_bad_parameters_raise_camb_errors():
    import pytest
    p = planck2018_params()
    p.optical_depth = 0.9  # cannot be reached below max_redshift
    with pytest.raises(CAMBError):
        HostState(p)
