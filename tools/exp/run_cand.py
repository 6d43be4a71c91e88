import os, sys, ctypes as C, subprocess
import numpy as np
HERE = os.path.dirname(os.path.abspath(__file__)); ROOT = os.path.dirname(os.path.dirname(HERE)); sys.path.insert(0, ROOT)
from oracle import vxo_py as O
so = "/tmp/libexp_cand.so"
subprocess.run(["g++", "-O2", "-ffp-contract=off", "-fopenmp", "-shared", "-fPIC", "-o", so, os.path.join(HERE, "exp_cand.cpp")], check=True)
L = C.CDLL(so)
d = np.load("/tmp/cfg3_scene.npz")
vol = np.ascontiguousarray(d["volume"]); sz, sy, sx = vol.shape
gbd = dict(depth24=d["depth24"], normal=d["normal"], material=d["material"], noise=d["noise"])
gb = O._gb(gbd); view = O._view(d["view"])
out = np.zeros(1024, np.float64)
bstep = int(sys.argv[1]) if len(sys.argv) > 1 else 8
L.exp_cand(vol.ctypes.data_as(C.c_void_p), sx, sy, sz, O._p(view), C.byref(gb), 16, bstep, out.ctypes.data_as(C.c_void_p))
c = out[:29 * 8].reshape(29, 8)
nr, nh = out[29 * 8], out[29 * 8 + 1]
print("rays", nr, "hit frac", nh / nr)
print(" k  performed  cell4   texel   hit    need32  need48  need64   (per ray)")
for k in range(29):
    print(f"{k:2d} {c[k,0]/nr:9.3f} {c[k,1]/nr:7.3f} {c[k,2]/nr:7.3f} {c[k,3]/nr:7.3f} {c[k,4]/nr:7.3f} {c[k,5]/nr:7.3f} {c[k,6]/nr:7.3f}")
s = c.sum(axis=0) / nr
print("sum", np.round(s, 3))
print("phase1 sum", np.round(c[:6].sum(axis=0) / nr, 3), "phase2 sum", np.round(c[6:].sum(axis=0) / nr, 3))
sp = out[29 * 8 + 2: 29 * 8 + 34]
print("block spread hist (4-voxel bins):", np.round(sp / sp.sum(), 3))
