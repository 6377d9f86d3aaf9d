"""Fits the polynomial of the GEMM epilogue GELU (csrc/gemm_tcgen05.cu gelu2): GELU(x) = max(x,0) - a 2^q(a), a = min(|x|, 6), q of degree 6
fitted to log2 Phi(-a) by iteratively re-weighted least squares on the error of a 2^q(a); prints the float32 coefficients and the error of
a float32 Horner/FMA evaluation against the exact erf form."""
import numpy as np
from scipy.special import log_ndtr, erf
A=6.0
def q_true(a): return log_ndtr(-a)/np.log(2.0)
deg=6
a=np.cos(np.linspace(0,np.pi,6001))*A/2+A/2
w=np.ones_like(a)
for it in range(120):
    sens=a*np.exp2(q_true(a))*np.log(2)+1e-12
    V=np.vander(a/A,deg+1,increasing=True)
    W=(w*sens)[:,None]
    c,*_=np.linalg.lstsq(V*W,q_true(a)*W[:,0],rcond=None)
    err=np.abs(a*np.exp2(V@c)-a*np.exp2(q_true(a)))
    w=w*(1+4*err/err.max()); w/=w.mean()
ca=np.array([c[k]/A**k for k in range(deg+1)])
cf=ca.astype(np.float32)
print("coef (float32, in a):", [repr(float(v)) for v in cf])
x=np.linspace(-8,8,400001).astype(np.float32)
af=np.minimum(np.abs(x),np.float32(6))
p=np.full_like(af,cf[-1])
for k in range(deg-1,-1,-1):
    p=(p.astype(np.float64)*af.astype(np.float64)+cf[k].astype(np.float64)).astype(np.float32)   # fma: one rounding
e=np.exp2(p.astype(np.float64)).astype(np.float32)
y=(-(af.astype(np.float64))*e.astype(np.float64)+np.maximum(x,0).astype(np.float64)).astype(np.float32)
xt=x.astype(np.float64); yt=0.5*xt*(1+erf(xt/np.sqrt(2)))
print("max abs err GELU:",np.abs(y-yt).max(), "at x=",x[np.abs(y-yt).argmax()])
m=np.abs(xt)>1e-3
print("max rel err (|x|>1e-3, x>-3):",(np.abs(y-yt)/np.abs(yt))[m&(xt>-3)].max())
