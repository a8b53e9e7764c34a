import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import numpy as np
from lwfbrook90.jl_b200 import capi
import oracle as orc
orc.use("elem")
rng = np.random.default_rng(0)
n = 20000
cases = {0: rng.uniform(-700, 700, n), 1: 10 ** rng.uniform(-300, 300, n), 3: rng.uniform(-50, 50, n), 4: rng.uniform(-50, 50, n),
         5: rng.uniform(-1.5, 1.5, n), 6: rng.uniform(-1, 1, n), 7: rng.uniform(-1, 1, n)}
for fn, x in cases.items():
    g = capi.elem_eval(fn, x); h = orc.elem(fn, x)
    bad = ~((g == h) | (np.isnan(g) & np.isnan(h)))
    print("fn", fn, "mismatches", bad.sum(), (x[bad][:3], g[bad][:3], h[bad][:3]) if bad.any() else "")
x = 10 ** rng.uniform(-6, 0, n); y = rng.uniform(-25, 25, n)
for fn in (2, 8):
    if fn == 8: x = 10 ** rng.uniform(-8, 3, n); y = np.where(rng.random(n) < 0.5, 0.14, 0.08)
    g = capi.elem_eval(fn, x, y); h = orc.elem(fn, x, y)
    bad = ~((g == h) | (np.isnan(g) & np.isnan(h)))
    print("fn", fn, "mismatches", bad.sum(), (x[bad][:3], y[bad][:3], g[bad][:3], h[bad][:3]) if bad.any() else "")
