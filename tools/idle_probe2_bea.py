"""Profiling helper: per-member per-day attempted steps (needs -DLWFB_DEBUG_ITERS: written to aux.theta[0])."""
import os, sys, warnings
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
warnings.simplefilter("ignore")
import numpy as np
import lwfbrook90.jl_b200 as lw
from lwfbrook90.jl_b200 import capi, ensemble as en, layout as lo
M, D = int(sys.argv[1]), int(sys.argv[2])
G = os.path.join(ROOT, "tests", "golden", "inputs", "BEA2016-reset-FALSE-input")
sim = lw.setup(lw.loadSPAC(G, "BEA2016-reset-FALSE", simulate_isotopes=False))
ens = en.make_ensemble(sim, M)
h = capi.Handle(ens.N, False, 0, D, record_daily=True)
h.set_sites(ens.forcing); h.set_members(ens.site_of_member, ens.scal, ens.lay, ens.u0); h.run()
ix = lo.StateIndex(ens.N, False)
out = np.zeros((M, D), dtype=np.int16)
for a in range(0, M, 960):
    r = h.get_daily(a, min(960, M - a))
    out[a:a + r.shape[0]] = r[:, 1:, ix.aux_theta.start].astype(np.int16)
np.save(os.path.join(ROOT, "gpurun_out", "steps_per_day_bea.npy"), out)
print("saved", out.shape, out.mean())
