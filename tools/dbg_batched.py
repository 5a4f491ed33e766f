import importlib, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from tests import helpers as Hp
from oracle import oracle as O
pvt = importlib.import_module("parallel-video-object-tracker_b200")
g = Hp.golden("maps.npz")
f, t = g["frame"], g["templ"]
ref = O.ncc_match_cpu(f, t)
# what an EMA-updated template would give
b, bx, by = O.max_loc(ref)
t2 = O.add_weighted(t, f[by:by+13, bx:bx+17].copy(), 0.10)
ref_ema = O.ncc_match_cpu(f, t2)
print("peak", b, bx, by, "ema-map vs ref", float(np.abs(ref_ema - ref).max()))
for rep in range(8):
    m = pvt.ncc_match_naive_cuda(f, t)
    d = np.abs(m - ref)
    bad = d > 1e-4
    print("single", rep, "max", float(d.max()), "nbad", int(bad.sum()), "vs ema-map", float(np.abs(m - ref_ema).max()),
          "bad rows", np.unique(np.nonzero(bad)[0])[:10], "bad cols", np.unique(np.nonzero(bad)[1])[:12])
