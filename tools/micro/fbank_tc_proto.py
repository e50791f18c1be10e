"""Numpy prototype of the folded, split-f16 DFT that fbank_tc_kernel runs on the tensor cores: checks the folding
identities and compares the power spectrum against an f64 DFT, next to torch's f32 FFT (run on CPU)."""
import numpy as np
rng=np.random.default_rng(0)
N=400
n=np.arange(N)
win=(0.54-0.46*np.cos(2*np.pi*n/N))
def make_W():
    m=np.arange(104)[:,None]; k=np.arange(112)[None,:]
    Ce=np.cos(np.pi*m*k/100.0); Ce[101:]=0; Ce[:,101:]=0
    Co=np.cos(np.pi*(2*m+1)*k/200.0); Co[100:]=0; Co[:,100:]=0
    Se=np.sin(np.pi*m*k/100.0); Se[101:]=0; Se[:,100:]=0; Se[:,0]=0
    So=np.sin(np.pi*(2*m+1)*k/200.0); So[100:]=0; So[:,101:]=0; So[:,0]=0
    return Ce,Co,Se,So
def fold(xw):
    # xw: windowed frame [400] float32
    xw=xw.astype(np.float32)
    p=np.zeros(112,np.float32); q=p.copy(); r=p.copy(); t=p.copy()
    for k in range(101):
        x0=xw[k]; x1=xw[400-k] if k>0 else np.float32(0)
        x2=xw[200-k]; x3=xw[200+k] if k>0 else np.float32(0)
        a_n=x0+x1; a_m=x2+x3; b_n=x0-x1; b_m=x2-x3
        if k==0: b_n=np.float32(0); b_m=np.float32(0)
        if k<100:
            p[k]=a_n+a_m; q[k]=a_n-a_m; r[k]=b_n-b_m; t[k]=b_n+b_m
        else:
            p[k]=a_n; q[k]=0; r[k]=0; t[k]=b_n
    return p,q,r,t
def split(v, s=2048.0):
    h=v.astype(np.float16); l=((v-h.astype(np.float32))*np.float32(s)).astype(np.float16)
    return h,l
Ws=make_W()
Wsp=[split(w.astype(np.float32)) for w in Ws]
def power_tc(x):
    mx=np.abs(x).max()
    e=np.frexp(mx)[1] if mx>0 else 0   # mx = f*2^e, f in [0.5,1)
    sc=np.float32(2.0**(1-e)) if mx>0 else np.float32(1)
    xw=(x.astype(np.float32)*win.astype(np.float32))*sc
    vs=fold(xw)
    out=[]
    for (wh,wl),v in zip(Wsp,vs):
        vh,vl=split(v)
        main=wh.astype(np.float32)@vh.astype(np.float32)
        corr=wh.astype(np.float32)@vl.astype(np.float32)+wl.astype(np.float32)@vh.astype(np.float32)
        out.append((main+corr*np.float32(1/2048.0)).astype(np.float32))
    ce,co,se,so=out
    P=np.zeros(201)
    P[0::2]=ce[:101].astype(np.float64)**2+se[:101].astype(np.float64)**2
    P[1::2]=co[:100].astype(np.float64)**2+so[:100].astype(np.float64)**2
    return P/ (float(sc)**2)
def power_ref(x):
    X=np.fft.rfft(x.astype(np.float64)*win)
    return np.abs(X)**2
def power_f32fft(x):
    import torch
    X=torch.fft.rfft(torch.from_numpy((x.astype(np.float32)*win.astype(np.float32))))
    return (X.real.double()**2+X.imag.double()**2).numpy()
for name,x in [('white',rng.standard_normal(400)*0.1),
               ('tone+weak', 0.5*np.sin(2*np.pi*0.05*n)+1e-4*rng.standard_normal(400)),
               ('quiet',1e-4*rng.standard_normal(400)),
               ('tilt', np.cumsum(rng.standard_normal(400))*0.01)]:
    pr=power_ref(x); pt=power_tc(x); pf=power_f32fft(x)
    rel_t=np.abs(pt-pr)/pr; rel_f=np.abs(pf-pr)/pr
    print(f'{name:10s} dyn range {10*np.log10(pr.max()/pr.min()):6.1f} dB | TC max rel {rel_t.max():.2e} med {np.median(rel_t):.2e} | f32 FFT max rel {rel_f.max():.2e} med {np.median(rel_f):.2e}')
