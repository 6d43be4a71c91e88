import os, sys, ctypes as C, subprocess
import numpy as np
HERE = os.path.dirname(os.path.abspath(__file__)); ROOT = os.path.dirname(os.path.dirname(HERE)); sys.path.insert(0, ROOT)
from oracle import vxo_py as O
so = "/tmp/libexp_clear.so"
subprocess.run(["g++", "-O2", "-ffp-contract=off", "-fopenmp", "-shared", "-fPIC", "-o", so, os.path.join(HERE, "exp_clear.cpp")], check=True)
L = C.CDLL(so)
d = np.load("/tmp/cfg3_scene.npz")
vol = np.ascontiguousarray(d["volume"]); sz, sy, sx = vol.shape
gbd = dict(depth24=d["depth24"], normal=d["normal"], material=d["material"], noise=d["noise"])
gb = O._gb(gbd); view = O._view(d["view"])
out = np.zeros(1024, np.float64)
bstep = int(sys.argv[1]) if len(sys.argv) > 1 else 8
L.exp_run(vol.ctypes.data_as(C.c_void_p), sx, sy, sz, O._p(view), C.byref(gb), 16, bstep, out.ctypes.data_as(C.c_void_p))
o = 0
for kind in ("sun", "ao"):
    for ph in (1, 2):
        c = out[o:o+9]; o += 9
        print(f"{kind} phase{ph}: probes {c[0]:.0f}  plain-clear 4/8/16/32: {c[1]/c[0]:.3f} {c[2]/c[0]:.3f} {c[3]/c[0]:.3f} {c[4]/c[0]:.3f}   dilated-clear: {c[5]/c[0]:.3f} {c[6]/c[0]:.3f} {c[7]/c[0]:.3f} {c[8]/c[0]:.3f}")
nr = out[o:o+2]; ts = out[o+2:o+4]; la = out[o+4:o+6]; lb = out[o+6:o+8]; o += 8; print('probes per ray', ts/nr)
print("rays sun/ao", nr, " lookups per ray (A):", la/nr, " candidates per ray:", lb/nr)
h = out[o:o+128].reshape(2, 64)
for k in range(2):
    print("hist lookups", ("sun","ao")[k], np.round(h[k]/h[k].sum(), 3)[:48])
