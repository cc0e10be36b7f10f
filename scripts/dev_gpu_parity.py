"""First GPU parity run: GPU path vs oracle on the same boundary arrays."""
import sys, time
import numpy as np
sys.path.insert(0,'.'); sys.path.insert(0,'tests')
from camb_b200.host import HostState, default_params
from camb_b200.gpu import GpuContext
import oracle_lib as O

def planck(p, lens=0, massless=False):
    p.H0=67.32117; p.ombh2=0.0223828; p.omch2=0.1201075; p.omnuh2=0.6451439E-03; p.Yhe=0.2454006
    p.Num_Nu_massless=2.046; p.Num_Nu_massive=1; p.Nu_mass_eigenstates=1; p.share_delta_neff=1
    p.Nu_mass_fractions[0]=1.0; p.Nu_mass_numbers[0]=0
    p.As=2.100549e-9; p.ns=0.9660499
    p.use_optical_depth=1; p.optical_depth=0.05430842
    p.DoLensing=lens; p.Max_l=2500; p.Max_eta_k=6250
    if massless:
        p.omnuh2=0; p.Num_Nu_massive=0; p.Nu_mass_eigenstates=0; p.Num_Nu_massless=3.046
    return p

def rel(a,b):
    s=np.abs(b).max()
    return np.abs(a-b).max()/s

for name,kw in [('massless',dict(massless=True)),('massive',dict()),('lensing',dict(lens=1))]:
    p=planck(default_params(), **kw)
    hs=HostState(p)
    t=time.time(); icl_o, aux = O.cls_from_params(hs, nthreads=8); t_or=time.time()-t
    ctx=GpuContext(0); ctx.set_bessel(hs); dm=ctx.create_model(hs)
    t=time.time()
    src,tr,st = ctx.evolve([hs],[dm]); t1=time.time()-t
    delta = ctx.project([hs],[dm]); icl = ctx.cl([hs],[dm])
    t2=time.time()-t
    print(f'== {name}: oracle {t_or:.2f}s  gpu evolve {t1:.3f}s total {t2:.3f}s kernels {ctx.timings_ms()}')
    s_g, s_o = src[0], aux['src']
    print('  sources  max|d|/max|ref| per src:', [rel(s_g[:,s,:], s_o[:,s,:]) for s in range(s_o.shape[1])])
    print('  stats: derivs gpu', st[0][:,0].sum(), 'oracle', aux['stats'][:,0].sum(), 'modes with different eval count', int((st[0][:,0]!=aux['stats'][:,0]).sum()), 'of', len(st[0]))
    print('  delta rel:', [rel(delta[0][:,:,s], aux['delta'][:,:,s]) for s in range(s_o.shape[1])])
    for c in range(icl_o.shape[1]):
        d = np.abs(icl[0][:,c]-icl_o[:,c]); 
        if c in (2,4,5): nrm = None
        print(f'  iCl col {c}: max rel {np.max(d/np.maximum(np.abs(icl_o[:,c]),1e-300)):.3e}  max |d|/max|ref| {d.max()/np.abs(icl_o[:,c]).max():.3e}')
    ctx.destroy_model(dm); ctx.close()
