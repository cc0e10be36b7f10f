"""Small end-to-end invocations of every kernel for compute-sanitizer (memcheck / racecheck): one low-resolution
cosmology per model kind.  Usage: compute-sanitizer --tool memcheck python scripts/sanitize.py"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from camb_b200.gpu import GpuContext
from camb_b200.host import HostState
from camb_b200.workloads import (isocurvature_params, lensing_window_params, params_counts_config4, planck2018_params,
                                 tensor_params)

ctx = GpuContext(0)
ctx.set_option("reg_min_modes", 0)   # small cases must run the register kernel, not the few-modes route


def scalar(hs, lanes):
    ctx.set_option("lanes_per_mode", lanes)
    ctx.set_bessel(hs)
    dm = ctx.create_model(hs)
    ctx.evolve([hs], [dm])
    if hs.model.WantTransfer:
        ctx.evolve([hs], [dm], transfer_only=True)
    ctx.project([hs], [dm])
    if hs.model.num_redshiftwindows:
        ctx.cl_windows([hs], [dm])
    else:
        ctx.cl([hs], [dm])
        if hs.n_src == 3 and hs.model.WantScalars:
            ctx.cl_array([hs], [dm])
    ctx.destroy_model(dm)
    print("ok", hs.n_src, lanes, flush=True)


NEW_ONLY = "--round2-only" in sys.argv     # the kernels added in round 2 alone
if NEW_ONLY:
    import numpy as np
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    hs = HostState(planck2018_params(lmax=150, lensing=True, massless=True))
    ctx.build_bessel(hs, want_table=True)                      # bessel_table.cu
    dm = ctx.create_model(hs, thermo_on_device=True)           # thermo_pack.cu
    ctx.evolve([hs], [dm])                                     # register kernel (three layouts), stage loop
    ctx.project([hs], [dm])                                    # lanes-on-l line-of-sight kernel on the device-built table
    ctx.set_option("los_kernel", 0)
    ctx.project([hs], [dm])                                    # the gather kernel
    ctx.set_option("los_kernel", 1)
    ctx.cl([hs], [dm])
    ctx.destroy_model(dm)
    pn = planck2018_params(lmax=150, lensing=True, massless=True, want_transfer=True)
    zs = [6.0, 2.0, 0.8, 0.3, 0.0]
    pn.transfer_PK_num_redshifts = len(zs)
    for i, z in enumerate(zs):
        pn.transfer_PK_redshifts[i] = z
    pn.transfer_kmax = 0.5
    hn = HostState(pn)
    ctx.set_bessel(hn)
    dm = ctx.create_model(hn)
    ctx.evolve([hn], [dm])
    q = np.asarray(hn.source_q())
    ratio = 1.0 + 0.3 * np.outer(np.linspace(0.2, 1.0, len(zs)), np.tanh(q * 10.0))
    ctx.make_nonlinear_sources(dm, ratio)                      # cb_nonlinear_sources_kernel
    ctx.set_option("async_calls", 1); ctx.set_option("async_outputs", 1)
    out = [np.zeros((len(hn.q_int), len(hn.ls), hn.n_src))]
    ctx.project([hn], [dm], want_delta=True, out=out)
    ctx.cl([hn], [dm])
    ctx.wait_outputs()
    ctx.set_option("async_outputs", 0); ctx.set_option("async_calls", 0)
    ctx.destroy_model(dm)
    ctx.close()
    print("sanitize run complete (round 2 kernels)")
    sys.exit(0)
LENS_ONLY = "--lens-only" in sys.argv      # racecheck of the lensing kernels alone (the whole script takes > 10 min under racecheck)
p = planck2018_params(lmax=150, lensing=True, massless=False, want_transfer=True)
p.transfer_kmax = 0.5
if not LENS_ONLY:
    scalar(HostState(p), 8)
if not LENS_ONLY:
    scalar(HostState(planck2018_params(lmax=150, lensing=True, massless=True)), 32)
if not LENS_ONLY:
    scalar(HostState(planck2018_params(lmax=150, lensing=False, massless=True)), 4)
if not LENS_ONLY:
    scalar(HostState(isocurvature_params(vector=[1, 0.2, 0.1, 0.1, 0.1], lmax=150)), 16)
if not LENS_ONLY:
    scalar(HostState(tensor_params(lmax_tensor=150, k_eta_max_tensor=400, massless=False)), 0)
if not LENS_ONLY:
    scalar(HostState(params_counts_config4(lmax=100)), 8)
if not LENS_ONLY:
    scalar(HostState(lensing_window_params(lmax=100, limber=False)), 8)
# lensing of the C_l: both methods, lmin 1 and 2, a batch of two, a theta slice
import numpy as np
from camb_b200 import lensing as L
from camb_b200.workloads import analytic_cmb_cls
for variant in (L.LENS_APODIZED, L.LENS_FULL):
    for lmin in (2, 1):
        plan = L.LensPlan(400, np.array(list(range(2, 400, 25)) + [350, 400]), lmin=lmin, variant=variant)
        cls = analytic_cmb_cls(plan.lmax)
        cs = cls.copy()
        ell = np.arange(len(cls), dtype=float)
        cs[2:, 3] = cls[2:, 3] * 2 * np.pi * ell[2:] ** 2 / (ell[2:] + 1) ** 2
        sp = np.ascontiguousarray(L.prepare_spectra(cs, cls[:, 3], lmin, plan.lmax, plan.lmax, plan.ee_taper, None))
        ctx.lens_contrib(plan, [sp, sp])
        ctx.lens_contrib(plan.theta_slice(1, 3), [sp])
        print("ok lens", variant, lmin, flush=True)
ctx.close()
print("sanitize run complete")
