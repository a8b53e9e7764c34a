"""Writes the input file of tests/c/multi_driver.c: a DAV2020-full ensemble and the final states of a single-handle run.
usage (on a GPU box): python tests/c/make_blocks.py out.bin [members] [days]"""
import os, sys, warnings
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
warnings.simplefilter("ignore")
import numpy as np
import lwfbrook90.jl_b200 as lw
from lwfbrook90.jl_b200 import capi, ensemble as en
out = sys.argv[1]
M = int(sys.argv[2]) if len(sys.argv) > 2 else 20000
D = int(sys.argv[3]) if len(sys.argv) > 3 else 30
sim = lw.setup(lw.loadSPAC(os.path.join(ROOT, "tests", "golden", "inputs", "DAV2020-full"), "DAV2020-full", simulate_isotopes=True))
ens = en.make_ensemble(sim, M)
h = capi.Handle(ens.N, True, 0, D, record_daily=False)
h.set_sites(ens.forcing); h.set_members(ens.site_of_member, ens.scal, ens.lay, ens.u0); h.run()
fin = h.get_final(); h.close()
with open(out, "wb") as f:
    np.array([ens.N, 1, D, ens.forcing.shape[1], M], dtype=np.int32).tofile(f)
    for a in (ens.forcing[0], ens.scal, ens.lay, ens.u0, fin):
        np.ascontiguousarray(a, dtype=np.float64).tofile(f)
print("wrote", out)
