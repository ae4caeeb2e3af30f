import numpy as np
from scipy.optimize import least_squares
from scipy.special import erf
R=6.0
x=np.linspace(-R,R,24001)
ref=x*0.5*(1+erf(x/np.sqrt(2)))
def model(c,x):
    u=x*x
    p=x*(c[0]+u*(c[1]+u*(c[2]+ (u*c[3] if len(c)>3 else 0))))
    return x/(1+np.exp(-p))
for deg in (3,4):
    c0=np.array([1.5957691216,0.0713548163,0.0,0.0][:deg])
    best=None
    w=np.ones_like(x)
    for it in range(60):
        res=least_squares(lambda c:(model(c,x)-ref)*w,c0,xtol=1e-15,ftol=1e-15,gtol=1e-15)
        c0=res.x
        err=np.abs(model(c0,x)-ref)
        if best is None or err.max()<best[0]: best=(err.max(),c0.copy())
        w=w*(1+4*err/err.max()); w/=w.mean()
    print(deg,best[0],list(best[1]))
    c=best[1]
    xx=np.linspace(-12,12,100001)
    xc=np.clip(xx,-R,R)
    u=xc*xc
    p=xc*(c[0]+u*(c[1]+u*(c[2]+(u*c[3] if len(c)>3 else 0))))
    y=xx/(1+np.exp(-p))
    r=xx*0.5*(1+erf(xx/np.sqrt(2)))
    print('  max abs err on [-12,12] with clamp:',np.abs(y-r).max(), 'at',xx[np.abs(y-r).argmax()])
    # float32 emulation
    c32=c.astype(np.float32); x32=xx.astype(np.float32); xc=np.clip(x32,-R,R).astype(np.float32); u=xc*xc
    L=np.float32(-1.4426950408889634)
    d=(c32*L).astype(np.float32)
    q=xc*(d[0]+u*(d[1]+u*(d[2]+(u*d[3] if len(c)>3 else np.float32(0)))))
    y32=x32/(np.float32(1)+np.exp2(q.astype(np.float32)))
    print('  fp32 eval max abs err:',np.abs(y32.astype(np.float64)-r).max(), 'coeffs*-log2e:',[float(v) for v in d])
