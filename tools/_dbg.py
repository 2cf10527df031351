import ctypes, numpy as np, sys, os
ROOT=os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0,ROOT)
import torch as T
from gaussian_process_edge_trace_b200 import _cabi, _gp_host as H
from gaussian_process_edge_trace_b200._cabi import call, ptr
lib=_cabi.load()
P=lambda a: a.ctypes.data_as(ctypes.c_void_p)
lo, hi = H.FINAL_BOUNDS[:, 0].copy(), H.FINAL_BOUNDS[:, 1].copy()
E=500
rng=np.random.RandomState(5)
x0=rng.uniform(lo,hi,size=(E,3)); cen=rng.uniform(lo-3,hi+3,size=(E,3)); sc=np.exp(rng.uniform(-2,1,size=(E,3)))
def fg(x):
    d=(x-cen)*sc
    return 0.5*(d*d).sum(axis=1)+np.cos(1.3*x).sum(axis=1), sc*d-1.3*np.sin(1.3*x)
nd,ni=lib.gpet_lbfgsb_state_doubles(),lib.gpet_lbfgsb_state_ints()
dev=T.device("cuda"); st=T.cuda.current_stream().cuda_stream

def run(fill_d, fill_i):
    d_state=T.full((nd,E),fill_d,dtype=T.float64,device=dev)
    if fill_i=="rand": i_state=T.randint(-3,5,(ni,E),dtype=T.int32,device=dev)
    else: i_state=T.full((ni,E),fill_i,dtype=T.int32,device=dev)
    d_lo,d_hi,d_x0=(T.from_numpy(a.copy()).to(dev) for a in (lo,hi,x0))
    d_tr=T.arange(E,dtype=T.int32,device=dev); d_theta=T.zeros((E,3),dtype=T.float64,device=dev)
    d_f=T.zeros(E,dtype=T.float64,device=dev); d_g=T.zeros((E,3),dtype=T.float64,device=dev)
    d_ev=T.full((E,),-1,dtype=T.int32,device=dev); d_n=T.zeros(1,dtype=T.int32,device=dev)
    call("gpet_lbfgsb_init_f64",ptr(d_state),ptr(i_state),E,ptr(d_x0),ptr(d_lo),ptr(d_hi),st)
    hs,hi_=np.zeros((E,nd)),np.zeros((E,ni),dtype=np.int32); need=np.zeros(E,dtype=np.int32); hx=np.zeros((E,3))
    lib.gpet_lbfgsb_host_init(P(hs),P(hi_),E,P(x0),P(lo),P(hi))
    give,f,g=np.zeros(E,dtype=np.int32),np.zeros(E),np.zeros((E,3))
    first=1
    for r in range(60):
        call("gpet_lbfgsb_advance_f64",ptr(d_state),ptr(i_state),E,first,ptr(d_tr),ptr(d_f),ptr(d_g),ptr(d_theta),ptr(d_ev),ptr(d_n),st)
        lib.gpet_lbfgsb_host_advance(P(hs),P(hi_),E,P(give),P(f),P(g),P(need),P(hx))
        th=d_theta.cpu().numpy(); ev=d_ev.cpu().numpy(); a_=need.astype(bool)
        if not np.array_equal(ev>=0,a_): print("  fill",fill_d,fill_i,"round",r,"need mismatch", np.nonzero((ev>=0)!=a_)[0][:5]); return
        if not a_.any(): print("  fill",fill_d,fill_i,"identical through",r,"rounds"); return
        dv=np.abs(th[a_]-hx[a_]).max(axis=1)
        if dv.max()>0:
            e=np.nonzero(a_)[0][np.argmax(dv)]
            print("  fill",fill_d,fill_i,"round",r,"theta differs, run",e,th[e],hx[e]); return
        f[:],g[:]=fg(hx); give[:]=need
        d_f.copy_(T.from_numpy(f)); d_g.copy_(T.from_numpy(g)); first=0
for fd in (0.0, float("nan"), 1e300, -1e300, float("inf"), 1e-310):
    for fi in (0, "rand", -1, 2147483647):
        run(fd, fi)
