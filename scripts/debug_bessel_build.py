"""Compare the device-built Bessel table with the host table (diagnostics for tests/test_gpu_parity.py::test_bessel_table_built_on_device)."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from camb_b200.gpu import GpuContext
from camb_b200.host import HostState
from camb_b200.workloads import planck2018_params
hs = HostState(planck2018_params(lmax=2500, lensing=True, massless=True))
b = hs.bessel()
H = np.ctypeslib.as_array(b.horner, shape=(b.n_l, b.nx - 1, 4)).copy()
x = np.ctypeslib.as_array(b.x, shape=(b.nx,)).copy()
ls = np.asarray(hs.ls)[:b.n_l]
ctx = GpuContext(0)
D = ctx.build_bessel(hs, want_table=True)
print("finite", np.isfinite(D).all(), "nan count", int(np.isnan(D).sum()))
zh, zd = (H == 0).all(axis=2), (D == 0).all(axis=2)
bad = np.argwhere(zh != zd)
print("zero-pattern mismatches", len(bad))
for j, i in bad[:12]:
    print("  l", ls[j], "interval", i, "x", x[i], x[i + 1], "host", H[j, i], "dev", D[j, i])
peak = np.abs(H[:, :, 0]).max(axis=1)[:, None, None]
err = np.abs(D - H) / peak
err = np.nan_to_num(err, nan=1e9)
j, i, c = np.unravel_index(np.argmax(err), err.shape)
print("max err / peak", err.max(), "at l", ls[j], "interval", i, "x", x[i], "coef", c, "host", H[j, i], "dev", D[j, i])
for j in (0, 5, 30, 60, b.n_l - 1):
    print("l", ls[j], "max err", err[j].max(), "first nonzero host", int(np.argmax(~zh[j])), "dev", int(np.argmax(~zd[j])))
