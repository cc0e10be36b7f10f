"""ctypes wrapper of libcambhost.so, the stand-in for the reference's Fortran host.

It produces the boundary data (`cb_model`, `cb_modes`, `cb_bessel`, integration k grid,
l samples, primordial power) that the reference's `cmbmain` would hand to the GPU path:
`InitVars` (fortran/cmbmain.f90:785-862), `SetkValuesForSources` (:864-936),
`InitTransfer` (:564-704), `SetkValuesForInt` (:1267-1349), `lSamples_init`
(fortran/results.f90:1505-1704), `InitSpherBessels` (fortran/bessels.f90:64-113).
"""
import ctypes as C
import os

import numpy as np

from ._cdefs import (cb_bessel, cb_model, cb_modes, cb_params, c_double_p, c_int_p)

_HERE = os.path.dirname(os.path.abspath(__file__))
_lib = None


class CAMBError(Exception):
    """Mirror of camb.baseconfig.CAMBError; `.code` is the CAMB error id (config.f90:62-72)."""

    def __init__(self, msg, code=-1):
        super().__init__(msg)
        self.code = code


def load_host_lib():
    global _lib
    if _lib is None:
        path = os.path.join(_HERE, "libcambhost.so")
        if not os.path.exists(path):
            raise ImportError(f"{path} not built; run `make host` (or __graft_entry__.build())")
        lib = C.CDLL(path)
        lib.cb_params_set_defaults.argtypes = [C.POINTER(cb_params)]
        lib.cbh_create.restype = C.c_void_p
        lib.cbh_create.argtypes = [C.POINTER(cb_params), c_int_p, C.c_char_p, C.c_int]
        lib.cbh_destroy.argtypes = [C.c_void_p]
        lib.cbh_get_model.argtypes = [C.c_void_p, C.POINTER(cb_model)]
        lib.cbh_get_modes.argtypes = [C.c_void_p, C.POINTER(cb_modes), C.c_int]
        lib.cbh_get_bessel.argtypes = [C.c_void_p, C.POINTER(cb_bessel)]
        lib.cbh_get_qint.argtypes = [C.c_void_p, c_int_p, C.POINTER(c_double_p), C.POINTER(c_double_p)]
        lib.cbh_get_ls.argtypes = [C.c_void_p, c_int_p, C.POINTER(c_int_p)]
        lib.cbh_get_pows.argtypes = [C.c_void_p, C.c_int, c_double_p]
        lib.cbh_scalar_power_at.argtypes = [C.c_void_p, C.c_int, c_double_p, c_double_p]
        lib.cbh_omnuh2_from_mass.restype = C.c_double
        lib.cbh_omnuh2_from_mass.argtypes = [C.c_double] * 4
        lib.cbh_get_derived.argtypes = [C.c_void_p, C.c_double * 13]
        lib.cbh_get_background.argtypes = [C.c_void_p, C.c_double * 16]
        lib.cbh_interp_cl.argtypes = [C.c_void_p, c_double_p, c_double_p]
        lib.cbh_lens_spectra.argtypes = [C.c_void_p, c_double_p, C.c_int, c_double_p, C.c_int, C.c_int, c_double_p, c_double_p]
        lib.cbh_lens_spectra_batch.argtypes = [C.c_int, C.POINTER(C.c_void_p), C.c_void_p, C.POINTER(c_double_p), C.c_int, c_double_p, C.c_int,
                                               C.c_int, c_double_p, c_double_p, C.c_int]
        lib.cbh_lens_assemble_batch.argtypes = [C.c_int, C.c_void_p, c_double_p, c_double_p, C.c_int, c_double_p, C.c_int]
        lib.cbh_dtauda.restype = C.c_double
        lib.cbh_dtauda.argtypes = [C.c_void_p, C.c_double]
        lib.cbh_xe.restype = C.c_double
        lib.cbh_xe.argtypes = [C.c_void_p, C.c_double]
        _lib = lib
    return _lib


def default_params():
    p = cb_params()
    load_host_lib().cb_params_set_defaults(C.byref(p))
    return p


def _dp(a):
    return a.ctypes.data_as(c_double_p)


def set_neutrino_masses(p, mnu, nnu=3.044, num_massive_neutrinos=1, standard_neutrino_neff=3.044):
    """CAMBparams_SetNeutrinoMasses (fortran/model.f90:364-405) with the degenerate hierarchy
    (CAMBparams_SetNeutrinoHierarchyBase :260-298): what `set_cosmology(mnu=...)` does in the reference's Python API."""
    if mnu == 0:
        p.omnuh2, p.Num_Nu_massive, p.Nu_mass_eigenstates, p.Num_Nu_massless, p.share_delta_neff = 0.0, 0, 0, nnu, 0
        return p
    neff_massive = num_massive_neutrinos * standard_neutrino_neff / 3
    p.Num_Nu_massive, p.Nu_mass_eigenstates, p.share_delta_neff = num_massive_neutrinos, 1, 0
    if nnu > neff_massive:
        p.Num_Nu_massless = nnu - neff_massive
    else:
        p.Num_Nu_massless, neff_massive = 0.0, nnu
    p.Nu_mass_numbers[0] = num_massive_neutrinos
    p.Nu_mass_degeneracies[0] = neff_massive
    mass = mnu / num_massive_neutrinos
    p.omnuh2 = load_host_lib().cbh_omnuh2_from_mass(mass, neff_massive, float(num_massive_neutrinos), p.TCMB)
    p.Nu_mass_fractions[0] = 1.0
    return p


DERIVED_NAMES = ["age", "zstar", "rstar", "thetastar", "DAstar", "zdrag", "rdrag", "kd", "thetad", "zeq", "keq",
                 "thetaeq", "thetarseq"]


class HostState:
    """One cosmology's host-side state (the part of CAMBdata the hot path reads)."""

    def __init__(self, params: cb_params):
        lib = load_host_lib()
        err = C.c_int(0)
        msg = C.create_string_buffer(1024)
        self._h = lib.cbh_create(C.byref(params), C.byref(err), msg, 1024)
        if not self._h:
            raise CAMBError(msg.value.decode() or "host setup failed", err.value)
        self._lib = lib
        self.params = params
        self.model = cb_model()
        lib.cbh_get_model(self._h, C.byref(self.model))
        m = self.model
        self.n_tau, self.n_src, self.nthermo = m.n_tau, m.n_src, m.nthermo
        self.tau = np.ctypeslib.as_array(m.tau, shape=(m.n_tau,))
        self.dtau = np.ctypeslib.as_array(m.dtau, shape=(m.n_tau,))
        self._bessel = None
        self._pows = [None, None]
        nl = C.c_int()
        lp = c_int_p()
        lib.cbh_get_ls(self._h, C.byref(nl), C.byref(lp))
        self.ls = np.ctypeslib.as_array(lp, shape=(nl.value,)).copy() if nl.value else np.zeros(0, np.int32)
        nq = C.c_int()
        q = c_double_p()
        dq = c_double_p()
        lib.cbh_get_qint(self._h, C.byref(nq), C.byref(q), C.byref(dq))
        if nq.value:
            self.q_int = np.ctypeslib.as_array(q, shape=(nq.value,))
            self.dq_int = np.ctypeslib.as_array(dq, shape=(nq.value,))
        else:
            self.q_int = self.dq_int = np.zeros(0)

    def modes(self, transfer_only=False):
        md = cb_modes()
        self._lib.cbh_get_modes(self._h, C.byref(md), 1 if transfer_only else 0)
        return md

    def source_q(self):
        md = self.modes(False)
        return np.ctypeslib.as_array(md.q, shape=(md.n_k,))

    def transfer_q(self):
        md = self.modes(True)
        return np.ctypeslib.as_array(md.q, shape=(md.n_k,)) if md.n_k else np.zeros(0)

    def bessel(self):
        if self._bessel is None:
            b = cb_bessel()
            e = self._lib.cbh_get_bessel(self._h, C.byref(b))
            if e:
                raise CAMBError("Bessel table generation failed", e)
            self._bessel = b
        return self._bessel

    def scalar_power(self):
        """InitPower%ScalarPower on the integration grid (cmbmain.f90:3199); computed once per cosmology."""
        if self._pows[0] is None:
            out = np.zeros(len(self.q_int))
            self._lib.cbh_get_pows(self._h, 0, _dp(out))
            out.flags.writeable = False
            self._pows[0] = out
        return self._pows[0]

    def scalar_power_at(self, k):
        k = np.ascontiguousarray(k, dtype=np.float64)
        out = np.zeros(k.shape)
        self._lib.cbh_scalar_power_at(self._h, k.size, _dp(k), _dp(out))
        return out

    def limber_power(self):
        """P_R(k = (l + 1/2)/(tau0 - tau_n)) on the (l sample, time step) grid of GetLimberTransfers
        (cmbmain.f90:499-501), shape (n_l_all, n_tau); zero where tau_n = tau0."""
        chi = self.model.tau0 - self.tau
        k = np.where(chi[None, :] > 0, (self.ls[:, None] + 0.5) / np.where(chi > 0, chi, 1.0)[None, :], 1.0)
        return np.where(chi[None, :] > 0, self.scalar_power_at(k), 0.0)

    def tensor_power(self):
        if self._pows[1] is None:
            out = np.zeros(len(self.q_int))
            self._lib.cbh_get_pows(self._h, 1, _dp(out))
            out.flags.writeable = False
            self._pows[1] = out
        return self._pows[1]

    def derived(self):
        d = (C.c_double * 13)()
        if self._lib.cbh_get_derived(self._h, d):
            return {}
        return dict(zip(DERIVED_NAMES, list(d)))

    def background(self):
        b = (C.c_double * 16)()
        self._lib.cbh_get_background(self._h, b)
        names = ["tau0", "taurst", "taurend", "tau_maxvis", "dtaurec", "scale", "tight_tau", "matter_verydom_tau",
                 "zre", "reion_opt_depth", "nu_mass0", "adotrad", "qmin", "qmax", "max_etak_scalar", "nthermo"]
        return dict(zip(names, list(b)))

    def interp_cl(self, iCl):
        """results.f90:1706-1743 InterpolateClArr for each column of iCl (n_l, n_cl) -> (lmax+1, n_cl)."""
        iCl = np.asarray(iCl, dtype=np.float64)
        lmin, lmax = int(self.params.Min_l), int(self.ls[-1])
        out = np.zeros((lmax + 1, iCl.shape[1]))
        for c in range(iCl.shape[1]):
            col = np.ascontiguousarray(iCl[:, c])
            tmp = np.zeros(lmax - lmin + 1)
            self._lib.cbh_interp_cl(self._h, _dp(col), _dp(tmp))
            out[lmin:, c] = tmp
        return out

    def lens_spectra(self, iCl, template, lmax, ee_taper=True):
        """InterpolateCls (templated for TT, EE, TE when a template is given) + PrepareLensedCLSpectra (lensing.f90:280-336):
        iCl (n_l, >=4) -> (Cl_scalar (Max_l+1, 4), spectra (4, lmax+1) for cb_lens_contrib)."""
        iCl = np.ascontiguousarray(iCl, dtype=np.float64)
        Max_l = int(self.ls[-1])
        cl = np.empty((Max_l + 1, 4))
        sp = np.empty((4, lmax + 1))
        if template is not None:
            assert template.flags.c_contiguous and template.shape[1] == 4 and template.shape[0] > lmax
        err = self._lib.cbh_lens_spectra(self._h, _dp(iCl), iCl.shape[1], _dp(template) if template is not None else None, int(lmax),
                                         int(bool(ee_taper)), _dp(cl), _dp(sp))
        if err:
            raise CAMBError({1: "lens_spectra: bad arguments", 2: "You need to normalize realistically to use lensing",
                             3: "lens_spectra: lmax > Max_l needs the high-L template",
                             4: "You need to normalize the high-L template so it is dimensionless"}.get(err, "lens_spectra"), err)
        return cl, sp

    def dtauda(self, a):
        return self._lib.cbh_dtauda(self._h, float(a))

    def x_e(self, a):
        return self._lib.cbh_xe(self._h, float(a))

    def close(self):
        if getattr(self, "_h", None):
            self._lib.cbh_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
