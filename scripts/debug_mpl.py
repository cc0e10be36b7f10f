"""Compare the mode-per-lane evolve kernel with the register kernel on one cosmology: per-mode source differences."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from camb_b200.gpu import GpuContext
from camb_b200.host import HostState
from camb_b200.workloads import planck2018_params
hs = HostState(planck2018_params(lmax=2500, lensing=True, massless=True))
ctx = GpuContext(0)
ctx.set_bessel(hs)
dm = ctx.create_model(hs)
res = {}
for mode in (1, 3):
    ctx.set_option("use_reg_kernel", mode)
    src, tr, st = ctx.evolve([hs], [dm])
    res[mode] = (src[0].copy(), st[0].copy())
q = np.array(hs.source_q()); tau = np.array(hs.tau)
a, b = res[1][0], res[3][0]          # (n_tau, n_src, n_k)
sa, sb = res[1][1], res[3][1]
scale = np.abs(a).max(axis=0)         # (n_src, n_k)
d = np.abs(a - b) / np.maximum(scale[None], 1e-300)
print("n_k", len(q), "max rel diff per source", d.max(axis=(0, 2)))
bad = np.where(d.max(axis=(0, 1)) > 1e-9)[0]
print("modes deviating:", len(bad), "of", len(q))
for k in list(bad[:6]) + list(bad[-6:]):
    j = np.argmax(d[:, :, k].max(axis=1) > 1e-9)
    print("  q=%.5g k_ix=%d first bad tau index %d (tau %.2f) maxdiff %.2e  derivs %d vs %d steps %d vs %d" % (
        q[k], k, j, tau[j], d[:, :, k].max(), sa[k, 0], sb[k, 0], sa[k, 1], sb[k, 1]))
for k in (0, 60, 150, 224):
    for sidx in range(a.shape[1]):
        dd = np.abs(a[:, sidx, k] - b[:, sidx, k])
        j = int(np.argmax(dd))
        nz = np.where(dd > 1e-13 * np.abs(a[:, sidx, k]).max())[0]
        print("   mode %d q=%.4g src %d: largest abs diff at tau[%d]=%.1f: ref %.6e mpl %.6e; first differing tau index %s (tau %.2f); last %s" % (
            k, q[k], sidx, j, tau[j], a[j, sidx, k], b[j, sidx, k], nz[0] if len(nz) else None, tau[nz[0]] if len(nz) else -1, nz[-1] if len(nz) else None))
print("tight switchoff of those modes:", [float(hs.modes(False).tight_switchoff[k]) for k in (0, 60, 150, 224)], "tau[1]", tau[1], "taurend", hs.model.taurend)
good = np.where(d.max(axis=(0, 1)) <= 1e-9)[0]
print("good modes q range:", (q[good].min(), q[good].max()) if len(good) else None)
for k in (0, 60, 150, 224):
    print("   mode", k, "stats reg", sa[k], "mpl", sb[k])
print("stats identical:", (sa[:, 0] == sb[:, 0]).mean())
ctx.destroy_model(dm); ctx.close()
