import sys, numpy as np, ctypes as C
sys.path.insert(0,'.')
import svgdcpp_b200 as sv
from svgdcpp_b200 import synth, _capi
n,d=65536,64
x0,means,covs=synth.mvn_problem(n,d)
model=sv.MultivariateNormal(means[0],covs[0])
s=sv.SVGD(d,1,x0,sv.GaussianRBFKernel(x0,sv.ScaleMethod.Median,model),model,sv.Adam(d,n,0.1,0.9,0.999),precision=1)
s.Initialize()
lib=_capi.load(); s._upload()
prev=None; out=[]
for it in range(40):
    lib.svgdb_step(s._ctx,1)
    st=s.Stats(); a=st['last_scale']; med2=np.log(n)/a
    out.append(med2)
    print(it, "med2=%.6f"%med2, "rel change %.3e"%((med2-prev)/prev if prev else 0), "extrap resid %.3e"%(((med2-(2*out[-2]-out[-3]))/med2) if len(out)>2 else 0), "passes",st['median_passes'],"hits",st['median_bracket_hits'])
    prev=med2
