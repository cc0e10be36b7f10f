import sys, time
import numpy as np
sys.path.insert(0,'.'); sys.path.insert(0,'tests')
from camb_b200.host import HostState, default_params
import oracle_lib as O
p = default_params()
# planck_2018.ini cosmology
p.H0=67.32117; p.ombh2=0.0223828; p.omch2=0.1201075; p.omnuh2=0.6451439E-03; p.Yhe=0.2454006
p.Num_Nu_massless=2.046; p.Num_Nu_massive=1; p.Nu_mass_eigenstates=1; p.share_delta_neff=1
p.Nu_mass_fractions[0]=1.0; p.Nu_mass_numbers[0]=0
p.As=2.100549e-9; p.ns=0.9660499
p.use_optical_depth=1; p.optical_depth=0.05430842; p.delta_redshift=0.5; p.helium_delta_redshift=0.5; p.helium_redshiftstart=3.5+5*0.5
p.RECFAST_fudge=1.125; p.RECFAST_fudge_He=0.86; p.RECFAST_He_rate_correction=0
p.AGauss1=-0.14; p.AGauss2=0.079; p.zGauss1=7.28; p.zGauss2=6.73; p.wGauss1=0.18; p.wGauss2=0.33
p.de_model=1
p.DoLensing = int(sys.argv[1]) if len(sys.argv)>1 else 0
p.Max_l=2500; p.Max_eta_k=6250
if len(sys.argv)>2 and sys.argv[2]=='massless':
    p.omnuh2=0; p.Num_Nu_massive=0; p.Nu_mass_eigenstates=0; p.Num_Nu_massless=3.046
t=time.time(); hs=HostState(p); print('host', time.time()-t, hs.background()); print(hs.derived())
t=time.time(); icl, aux = O.cls_from_params(hs, nthreads=8); print('oracle chain', time.time()-t)
st=aux['stats']; print('n_k',len(st),'derivs evals total', st[:,0].sum(), 'max', st[:,0].max())
allcl = hs.interp_cl(icl)
T = np.loadtxt('/root/reference/fortran/HighLExtrapTemplate_lenspotentialCls.dat')
L=T[:,0].astype(int)
for name,col,tc in [('TT',0,1),('EE',1,2),('TE',2,4)]:
    ls=np.array([2,10,30,100,220,500,800,1000,1500,2000,2400])
    mine=allcl[ls,col]; ref=T[ls-2,tc]
    if name=='TE':
        nrm=np.sqrt(T[ls-2,1]*T[ls-2,2]); print(name,'(d/sqrt(TT EE))', np.array2string((mine-ref)/nrm,precision=4))
    else: print(name, np.array2string(mine/ref-1,precision=4))
ls=np.arange(30,2400)
print('max |dTT/TT| 30..2400', np.abs(allcl[ls,0]/T[ls-2,1]-1).max(), ' EE', np.abs(allcl[ls,1]/T[ls-2,2]-1).max())
if icl.shape[1]>3:
    ls=np.array([2,10,30,100,400,1000,2000]); 
    # template PP = [l(l+1)]^2 C_phi/2pi ; iCl PP col 3 = l^4 C_phi *4pi... compare after conversion
    print('PP raw ratio', allcl[ls,3]/T[ls-2,5])
np.save('/tmp/icl.npy', icl)
ls_s = hs.ls
sel = ls_s>=2
print('at sampled l: max|dTT/TT|', np.abs(icl[sel,0]/T[ls_s[sel]-2,1]-1).max(), 'EE (l>=30)', np.abs(icl[ls_s>=30,1]/T[ls_s[ls_s>=30]-2,2]-1).max())
print(np.array2string(np.c_[ls_s, icl[:,0]/T[ls_s-2,1]-1, icl[:,1]/T[ls_s-2,2]-1][::4], precision=5, suppress_small=True))
