import os, sys
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/tests")
os.environ["VBNMF_NO_GRAPH"] = "1"
import scipy.sparse as sp
from ccfindr_b200 import synth
from ccfindr_b200.engine import Engine
hyper = dict(aw=1.0, bw=1.0, ah=1.0, bh=1.0)
x = sp.csc_matrix(synth.simulate_whx(1000, 200, 3, seed=1)["x"])
w0, h0 = synth.random_init(*x.shape, 3, hyper, seed=3)
with Engine(x) as eng:
    eng.set_state(w0, h0)
    eng.run(hyper, Itmax=40, Tol=0.0)
