import ctypes as C, sys, time
import numpy as np
sys.path.insert(0,'.')
from camb_b200._cdefs import *
lib = C.CDLL('camb_b200/libcambhost.so')
lib.cbh_create.restype = C.c_void_p
lib.cbh_create.argtypes = [C.POINTER(cb_params), c_int_p, C.c_char_p, C.c_int]
p = cb_params(); lib.cb_params_set_defaults(C.byref(p))
p.H0=67.5; p.ombh2=0.022; p.omch2=0.122
# mnu=0.06 one massive
p.Num_Nu_massive=1; p.Nu_mass_eigenstates=1; p.share_delta_neff=1; p.Num_Nu_massless=2.044
p.Nu_mass_fractions[0]=1.0; p.Nu_mass_numbers[0]=1
p.omnuh2=0.06/93.14*( (3.044/3)**0.75)  # approx
p.Yhe=0.2453
p.use_optical_depth=0; p.redshift=10.0  # python default Reion? unknown
p.DoLensing=0
err=C.c_int(0); msg=C.create_string_buffer(512)
t=time.time()
st = lib.cbh_create(C.byref(p), C.byref(err), msg, 512)
print('create', time.time()-t, err.value, msg.value)
d=(C.c_double*13)(); lib.cbh_get_derived.argtypes=[C.c_void_p, C.c_double*13]
lib.cbh_get_derived(st,d)
names=['age','zstar','rstar','thetastar','DAstar','zdrag','rdrag','kd','thetad','zeq','keq','thetaeq','thetarseq']
ref={'age': 13.74045123110015, 'zstar': 1090.5554322640728, 'rstar': 144.21085063851663, 'thetastar': 1.0446452473184842, 'DAstar': 13.804767791620522, 'zdrag': 1059.215089050609, 'rdrag': 146.98865612136981, 'kd': 0.14069733804243129, 'thetad': 0.1617464978660653, 'zeq': 3441.8515054149993, 'keq': 0.010503792649954168, 'thetaeq': 0.8079815083469327, 'thetarseq': 0.4469753307013929}
for n,v in zip(names,d): print(f'{n:10s} {v:14.7f} ref {ref[n]:14.7f} rel {v/ref[n]-1:+.2e}')
b=(C.c_double*16)(); lib.cbh_get_background.argtypes=[C.c_void_p, C.c_double*16]; lib.cbh_get_background(st,b)
print('tau0,taurst,taurend,tau_maxvis,dtaurec,scale,tight_tau,verydom,zre,optdepth,numass,adotrad,qmin,qmax,maxetak,nthermo'); print(list(b))
m=cb_model(); lib.cbh_get_model.argtypes=[C.c_void_p, C.POINTER(cb_model)]; lib.cbh_get_model(st, C.byref(m))
print('n_tau', m.n_tau, 'nthermo', m.nthermo, 'nqmax', m.nqmax)
md=cb_modes(); lib.cbh_get_modes.argtypes=[C.c_void_p, C.POINTER(cb_modes), C.c_int]; lib.cbh_get_modes(st, C.byref(md), 0)
print('n_k', md.n_k)
nq=C.c_int(); q=c_double_p(); dq=c_double_p(); lib.cbh_get_qint.argtypes=[C.c_void_p, c_int_p, C.POINTER(c_double_p), C.POINTER(c_double_p)]
lib.cbh_get_qint(st, C.byref(nq), C.byref(q), C.byref(dq)); print('n_q', nq.value)
nl=C.c_int(); lp=c_int_p(); lib.cbh_get_ls.argtypes=[C.c_void_p, c_int_p, C.POINTER(c_int_p)]; lib.cbh_get_ls(st,C.byref(nl),C.byref(lp)); print('n_l', nl.value, [lp[i] for i in range(nl.value)])
t=time.time(); bs=cb_bessel(); lib.cbh_get_bessel.argtypes=[C.c_void_p, C.POINTER(cb_bessel)]; e=lib.cbh_get_bessel(st, C.byref(bs)); print('bessel', e, time.time()-t, bs.nx, bs.n_l)
# check bessel table vs scipy
from scipy.special import spherical_jn
H=np.ctypeslib.as_array(bs.horner, shape=(bs.n_l, bs.nx-1, 4)); X=np.ctypeslib.as_array(bs.x, shape=(bs.nx,))
for j in [0,5,30,60,88]:
    l=lp[j]; ex=spherical_jn(l, X[1:]); got=H[j,:,0]
    pk=np.abs(ex).max(); print('l',l,'max abs err/peak', np.abs(got-ex).max()/pk, 'zeroed below x=', X[1:][np.nonzero(got)[0][0]])
