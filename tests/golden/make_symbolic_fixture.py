#!/usr/bin/env python
"""Per-evaluation known answers for `derivs` and the temperature / polarisation sources, authored by the
reference: /root/reference/camb/symbolic.py (the sympy statement of the scalar equations that the reference's own
`testSymbolic`, camb/tests/camb_test_slow.py:179-215, compiles into CAMB and compares with the built-in sources).

What is taken from the reference (imported stand-alone and EXECUTED -- untrusted public code: run this script in a
sandbox / throw-away container, never on a machine with credentials):
  * the constraint equations `constraints` (symbolic.py:360-364) solved for z, sigma, phi, hdot (`var_subs`, :367-373);
  * the fluid equations `delta_eqs`, `vel_eqs` (:537-572), d eta/dt (`deta`, :382), the totals `tot_pert_subs`,
    `tot_subs` (:576-600), background `background_eqs`, `density_eqs` (:357, 528-535);
  * the Boltzmann hierarchies `J_eq`, `G_eq`, `E_eq` (:619-652) for photons, massless neutrinos, E polarisation;
  * the line-of-sight sources `get_scalar_temperature_sources()` (:674-722) and `scalar_E_source` (:671).
Everything is projected to the synchronous (CDM-frame) gauge with the reference's `cdm_gauge` (:486-497), K = 0.

What is taken from this repository: the STATES.  There is no CAMB binary here, so the points (k, tau, y) at which the
equations are evaluated come from the CPU oracle's own evolution of two cosmologies (oracle_debug_states), and the
background / ionisation values at tau (a, densities, opacity, visibility and their derivatives) from the stand-in host's
tables.  They are INPUT DATA: the fixture asserts "given this state and this background, the reference's equations
give these derivatives and these sources", evaluated with 40-digit mpmath arithmetic.  The tests then require the
oracle's `derivs` (tests/test_symbolic_fixture.py) and the CUDA `derivs` (-m gpu) to reproduce them.

Every time derivative inside the sources (phi', sigma', v_b', polter', polter'', Pi') is eliminated with the reference's
own first-order equations, so the expected source is a function of the state and the background only -- nothing of
the oracle's intermediate algebra enters.

One deliberate, documented difference: the reference's Fortran drops the baryon-pressure term of d Delta_b/dt
(equations.f90:2446 `clxbdot=-k*(z+vb)`) that symbolic.py keeps (`(p_b/rho_b - csq_b)*3*H*Delta_b`, :541); the
fixture stores the symbolic value with that term removed and says so in `notes`.

Usage:  python tests/golden/make_symbolic_fixture.py   ->  tests/golden/symbolic_derivs.json
"""
import ctypes as C
import importlib.util
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

REF = "/root/reference/camb/symbolic.py"
LMAX_SYM = 6   # hierarchies l = 2..5 from the reference's generators (they need l+1 as input)


def load_symbolic():
    spec = importlib.util.spec_from_file_location("camb_symbolic_ref", REF)
    s = importlib.util.module_from_spec(spec)
    sys.modules["camb_symbolic_ref"] = s
    spec.loader.exec_module(s)
    return s


def build_expressions(s):
    """Returns (input symbols, {name: sympy expression of the inputs})."""
    import sympy
    from sympy import Derivative, Eq, Symbol, diff

    t, k = s.t, s.k
    J = {l: s._make_index_func("J", l) for l in range(3, LMAX_SYM + 1)}
    G = {l: s._make_index_func("G", l) for l in range(3, LMAX_SYM + 1)}
    E = {l: s._make_index_func("E", l) for l in range(2, LMAX_SYM + 1)}
    hier = []
    for l in range(2, LMAX_SYM):
        hier += [s.J_eq(l), s.G_eq(l), s.E_eq(l)]
    # first-order system in the CDM frame: A = 0, v_c = 0
    first_order = s.cdm_gauge(s.delta_eqs + s.vel_eqs) + hier + s.density_eqs + s.background_eqs + \
        [Eq(diff(s.eta, t), s.cdm_gauge(s.deta))]
    first_order = [e for e in first_order if isinstance(e, sympy.Equality)]
    dmap = {e.lhs: e.rhs for e in first_order}
    # w = -1 dark energy without perturbations, no massive neutrinos, no baryon pressure (as the Fortran)
    zero = {s.rho_nu: 0, s.p_nu: 0, s.Delta_nu: 0, s.q_nu: 0, s.pi_nu: 0, s.Delta_P_nu: 0, s.Delta_de: 0, s.v_de: 0,
            s.p_b: 0, s.w_de: -1, s.A: 0, s.v_c: 0, s.K: 0, s.K_fac: 1}
    for l in range(1, LMAX_SYM + 2):
        zero[s.Kf[l]] = 1

    def close(expr):
        """Eliminate time derivatives of perturbations / background with the first-order system, then the
        constrained variables and the totals."""
        expr = sympy.sympify(expr)
        for _ in range(4):
            expr = expr.doit()
            ders = [d for d in expr.atoms(Derivative)]
            rep = {}
            for d in ders:
                if d in dmap:
                    rep[d] = dmap[d]
            if not rep:
                break
            expr = expr.subs(rep)
        for _ in range(3):
            expr = s.subs(s.var_subs, expr)                     # z, sigma, phi, hdot from the constraints
            expr = s.subs([s.polter_sub], expr)
            expr = s.subs(s.tot_pert_subs + s.tot_subs, expr)   # delta, q, Pi, delta_P, rho, P
            expr = expr.doit()
            ders = [d for d in expr.atoms(Derivative) if d in dmap]
            if ders:
                expr = expr.subs({d: dmap[d] for d in ders})
        return expr.subs(zero).doit()

    # --- numeric input symbols ---
    names = ("k tau tau0 a H rho_b rho_c rho_g rho_r rho_de csq_b opacity dopacity ddopacity visibility dvisibility "
             "ddvisibility exptau eta Delta_c Delta_b v_b Delta_g q_g pi_g Delta_r q_r pi_r").split()
    names += [f"J_{l}" for l in range(3, LMAX_SYM + 1)] + [f"G_{l}" for l in range(3, LMAX_SYM + 1)] + \
        [f"E_{l}" for l in range(2, LMAX_SYM + 1)]
    sym = {n: Symbol(n) for n in names}
    to_num = {s.a: sym["a"], s.H: sym["H"], s.rho_b: sym["rho_b"], s.rho_c: sym["rho_c"], s.rho_g: sym["rho_g"],
              s.rho_r: sym["rho_r"], s.rho_de: sym["rho_de"], s.csq_b: sym["csq_b"], s.opacity: sym["opacity"],
              s.visibility: sym["visibility"], s.exptau: sym["exptau"], s.eta: sym["eta"], s.Delta_c: sym["Delta_c"],
              s.Delta_b: sym["Delta_b"], s.v_b: sym["v_b"], s.Delta_g: sym["Delta_g"], s.q_g: sym["q_g"],
              s.pi_g: sym["pi_g"], s.Delta_r: sym["Delta_r"], s.q_r: sym["q_r"], s.pi_r: sym["pi_r"],
              s.kappa: 1, s.tau0: sym["tau0"]}
    for l in range(3, LMAX_SYM + 1):
        to_num[J[l]] = sym[f"J_{l}"]
        to_num[G[l]] = sym[f"G_{l}"]
    for l in range(2, LMAX_SYM + 1):
        to_num[E[l]] = sym[f"E_{l}"]
    der_num = {Derivative(s.opacity, t): sym["dopacity"], Derivative(s.opacity, (t, 2)): sym["ddopacity"],
               Derivative(s.visibility, t): sym["dvisibility"], Derivative(s.visibility, (t, 2)): sym["ddvisibility"]}

    def numeric(expr):
        expr = close(expr)
        expr = expr.subs(der_num).subs(to_num)
        expr = expr.subs({s.f_K(sym["tau0"] - t): sym["tau0"] - sym["tau"]}).subs(t, sym["tau"])
        left = [str(x) for x in expr.atoms(sympy.Function)] + [str(x) for x in expr.atoms(Derivative)]
        if left:
            raise RuntimeError(f"unresolved symbols: {left}")
        return expr

    out = {}
    out["z"] = numeric(s.z)
    out["sigma"] = numeric(s.sigma)
    out["phi"] = numeric(s.phi)
    out["d_eta"] = numeric(diff(s.eta, t))
    out["d_Delta_c"] = numeric(diff(s.Delta_c, t))
    # equations.f90:2446 drops the baryon-pressure term that symbolic.py:541 keeps
    out["d_Delta_b"] = numeric(dmap[Derivative(s.Delta_b, t)] - (s.p_b / s.rho_b - s.csq_b) * 3 * s.H * s.Delta_b)
    out["d_v_b"] = numeric(diff(s.v_b, t))
    out["d_Delta_g"] = numeric(diff(s.Delta_g, t))
    out["d_q_g"] = numeric(diff(s.q_g, t))
    out["d_pi_g"] = numeric(diff(s.pi_g, t))
    out["d_Delta_r"] = numeric(diff(s.Delta_r, t))
    out["d_q_r"] = numeric(diff(s.q_r, t))
    out["d_pi_r"] = numeric(diff(s.pi_r, t))
    for l in range(3, LMAX_SYM):
        out[f"d_J_{l}"] = numeric(diff(J[l], t))
        out[f"d_G_{l}"] = numeric(diff(G[l], t))
    for l in range(2, LMAX_SYM):
        out[f"d_E_{l}"] = numeric(diff(E[l], t))
    mono, isw, dop, quad = s.get_scalar_temperature_sources()
    out["src_monopole"] = numeric(mono)
    out["src_ISW"] = numeric(isw)
    out["src_doppler"] = numeric(dop)
    out["src_quadrupole"] = numeric(quad)
    out["src_E"] = numeric(s.scalar_E_source)
    return [sym[n] for n in names], names, out


def main():
    import mpmath
    import sympy

    import oracle_lib as O
    from camb_b200._cdefs import c_double_p
    from camb_b200.host import HostState
    from camb_b200.workloads import planck2018_params, synthetic_lcdm_batch

    s = load_symbolic()
    syms, names, exprs = build_expressions(s)
    mpmath.mp.dps = 40
    funcs = {nm: sympy.lambdify(syms, e, modules="mpmath") for nm, e in exprs.items()}

    lib = O.load()
    c_int_p = C.POINTER(C.c_int)
    lib.oracle_debug_states.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, c_double_p, C.c_int, c_double_p, c_int_p]
    lib.oracle_debug_derivs.argtypes = [C.c_void_p, C.c_double, C.c_double, c_int_p, c_double_p, C.c_int, c_double_p,
                                        c_double_p, c_double_p, c_int_p]

    cosmos = [("planck2018_massless_lcdm", planck2018_params(lmax=2500, lensing=True, massless=True)),
              ("box_sample_massless_lcdm", synthetic_lcdm_batch(3, seed=5, lmax=2500, lensing=True, massless=True)[2])]
    rng = np.random.default_rng(20261018)
    states = []
    for cname, p in cosmos:
        hs = HostState(p)
        m = hs.model
        md = hs.modes(False)
        qs = np.array([md.q[i] for i in range(md.n_k)])
        tau0, taurend = m.tau0, m.taurend
        # physical trajectory states: modes whose full hierarchies are alive between the end of tight coupling and the
        # k*tau switch; times around recombination, the visibility tail, reionisation and late ISW
        for k_target in (2e-4, 1e-3, 5e-3, 1.5e-2, 4e-2):
            q_ix = int(np.argmin(np.abs(qs - k_target)))
            times = np.array([230.0, 262.0, 280.0, 295.0, 320.0, 380.0, 600.0, 1500.0, 4500.0, 9000.0, 13000.0])
            times = times[times < 0.98 * tau0]
            ld = 512
            y = np.zeros((len(times), ld))
            lay = np.zeros((len(times), 16), dtype=np.int32)
            e = lib.oracle_debug_states(C.byref(m), C.byref(md), q_ix, len(times), times.ctypes.data_as(c_double_p), ld,
                                        y.ctypes.data_as(c_double_p), lay.ctypes.data_as(c_int_p))
            assert e == 0, e
            for i, tau in enumerate(times):
                L = lay[i]
                if L[9] or L[10] or L[11] or L[12]:
                    continue   # an approximation is active: the reference's symbolic module does not describe it
                states.append(dict(cosmo=cname, kind="trajectory", q=float(qs[q_ix]), tau=float(tau),
                                   y=y[i, :L[0]].tolist()))
        # off-trajectory states at high k (on the trajectory the neutrino hierarchy is already truncated there):
        # smooth random multipoles of realistic size around recombination
        for q in (0.08, 0.2, 0.45):
            for tau in (255.0, 283.0, 310.0):
                flags = (C.c_int * 4)(0, 0, 0, 0)
                lay = np.zeros(16, dtype=np.int32)
                dummy = np.zeros(1)
                lib.oracle_debug_derivs(C.byref(m), q, tau, flags, dummy.ctypes.data_as(c_double_p), -1,
                                        dummy.ctypes.data_as(c_double_p), dummy.ctypes.data_as(c_double_p), None,
                                        lay.ctypes.data_as(c_int_p))
                n = int(lay[0])
                yv = rng.normal(size=n) * np.exp(-0.35 * np.arange(n) % 7)
                yv[:4] = rng.normal(size=4) * np.array([q * 2, 3.0, 1.5, 0.2])
                states.append(dict(cosmo=cname, kind="random", q=float(q), tau=float(tau), y=yv.tolist()))
    print(len(states), "states")

    hs_by_name = {cname: HostState(p) for cname, p in cosmos}
    out_states = []
    for st in states:
        hs = hs_by_name[st["cosmo"]]
        m = hs.model
        yv = np.array(st["y"])
        n = len(yv)
        flags = (C.c_int * 4)(0, 0, 0, 0)
        yp = np.zeros(n); src = np.zeros(3); aux = np.zeros(16); lay = np.zeros(16, dtype=np.int32)
        e = lib.oracle_debug_derivs(C.byref(m), st["q"], st["tau"], flags, yv.ctypes.data_as(c_double_p), n,
                                    yp.ctypes.data_as(c_double_p), src.ctypes.data_as(c_double_p),
                                    aux.ctypes.data_as(c_double_p), lay.ctypes.data_as(c_int_p))
        assert e == 0, (e, n, lay)
        g_ix, polind, r_ix = int(lay[2]), int(lay[3]), int(lay[4])
        a = aux[0]
        k = st["q"]
        Y = lambda i: yv[i - 1]   # noqa: E731  (1-based like the reference)
        v = dict(k=k, tau=st["tau"], tau0=m.tau0, a=a, H=aux[1], rho_b=aux[2] / a**2, rho_c=aux[3] / a**2,
                 rho_g=aux[4] / a**2, rho_r=aux[5] / a**2, rho_de=aux[6] / a**2, csq_b=aux[7], opacity=aux[8],
                 dopacity=aux[9], ddopacity=aux[10], visibility=aux[11], dvisibility=aux[12], ddvisibility=aux[13],
                 exptau=aux[14], eta=-2 * Y(1) / k, Delta_c=Y(2), Delta_b=Y(3), v_b=Y(4), Delta_g=Y(g_ix), q_g=Y(g_ix + 1),
                 pi_g=Y(g_ix + 2), Delta_r=Y(r_ix), q_r=Y(r_ix + 1), pi_r=Y(r_ix + 2))
        lmaxg, lmaxgpol, lmaxnr = int(lay[6]), int(lay[7]), int(lay[8])
        # multipoles beyond a (shallow, low-k) hierarchy do not exist: zero, and the equations that would read them
        # (and the reference's truncated last multipole, which symbolic.py does not describe) are not compared
        for l in range(3, LMAX_SYM + 1):
            v[f"J_{l}"] = Y(g_ix + l) if l <= lmaxg else 0.0
            v[f"G_{l}"] = Y(r_ix + l) if l <= lmaxnr else 0.0
        for l in range(2, LMAX_SYM + 1):
            v[f"E_{l}"] = Y(polind + l) if l <= lmaxgpol else 0.0
        args = [mpmath.mpf(float(v[nm])) for nm in names]
        exp = {nm: float(f(*args)) for nm, f in funcs.items()}
        # the symbolic answers translated to the reference's Fortran variables (symbolic.py camb_sub attributes:
        # eta = -2 etak / k, everything else one to one)
        expected = dict(
            z=exp["z"], sigma=exp["sigma"], phi=exp["phi"],
            yprime={"1": -k * exp["d_eta"] / 2, "2": exp["d_Delta_c"], "3": exp["d_Delta_b"], "4": exp["d_v_b"],
                    str(g_ix): exp["d_Delta_g"], str(g_ix + 1): exp["d_q_g"], str(g_ix + 2): exp["d_pi_g"],
                    str(r_ix): exp["d_Delta_r"], str(r_ix + 1): exp["d_q_r"], str(r_ix + 2): exp["d_pi_r"]},
            src_T=exp["src_monopole"] + exp["src_ISW"] + exp["src_doppler"] + exp["src_quadrupole"],
            src_T_terms=[exp["src_monopole"], exp["src_ISW"], exp["src_doppler"], exp["src_quadrupole"]],
            src_E=exp["src_E"])
        for l in range(3, LMAX_SYM):
            if l < lmaxg:
                expected["yprime"][str(g_ix + l)] = exp[f"d_J_{l}"]
            if l < lmaxnr:
                expected["yprime"][str(r_ix + l)] = exp[f"d_G_{l}"]
        for l in range(2, LMAX_SYM):
            if l < lmaxgpol:
                expected["yprime"][str(polind + l)] = exp[f"d_E_{l}"]
        if lmaxg < 4 or lmaxgpol < 4:   # polter'' reads J_3', E_3': truncated multipoles there
            for key in ("src_T", "src_T_terms"):
                expected.pop(key)
        # report how the oracle does against it right now (informational; the tests enforce it)
        err_y = max(abs(yp[int(i) - 1] - val) / max(abs(val), np.abs(yp).max() * 1e-3) for i, val in expected["yprime"].items())
        err_T = float("nan")
        if "src_T" in expected:
            scale_T = max(abs(x) for x in expected["src_T_terms"])
            err_T = abs(src[0] - expected["src_T"]) / scale_T
        err_E = abs(src[1] - expected["src_E"]) / max(abs(expected["src_E"]), 1e-300)
        print(f"{st['cosmo'][:8]} {st['kind'][:4]} k={k:.4g} tau={st['tau']:.0f} n={n}  yprime {err_y:.1e}  T {err_T:.1e}  E {err_E:.1e}")
        out_states.append(dict(cosmo=st["cosmo"], kind=st["kind"], q=k, tau=st["tau"], y=st["y"],
                               layout=dict(neq=n, g_ix=g_ix, polind=polind, r_ix=r_ix, lmaxg=int(lay[6]), lmaxgpol=int(lay[7]),
                                           lmaxnr=int(lay[8])),
                               background={nm: float(v[nm]) for nm in names[2:18]}, expected=expected))
    doc = dict(
        source="/root/reference/camb/symbolic.py: constraints :360-373, delta_eqs/vel_eqs :537-572, J_eq/G_eq/E_eq :619-652, "
               "get_scalar_temperature_sources :674-722, scalar_E_source :671; cdm_gauge :486-497",
        generator="tests/golden/make_symbolic_fixture.py (mpmath, 40 digits)",
        notes=["states and background values are inputs produced by the repository's oracle / stand-in host; expected values "
               "are the reference's symbolic equations evaluated on them",
               "d Delta_b/dt: baryon-pressure term of symbolic.py:541 removed, as equations.f90:2446 does",
               "yprime keys are 1-based indices into y (the reference's layout, equations.f90:585-701)"],
        cosmologies=[c[0] for c in cosmos], states=out_states)
    path = os.path.join(ROOT, "tests", "golden", "symbolic_derivs.json")
    with open(path, "w") as f:
        json.dump(doc, f)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
