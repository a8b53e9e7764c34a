"""Small fixed workload for ncu: DAV2020-full isotope ensemble, M members x D days, one launch."""
import os, sys, warnings
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
warnings.simplefilter("ignore")
import numpy as np
import lwfbrook90.jl_b200 as lw
from lwfbrook90.jl_b200 import capi, ensemble as en
M = int(sys.argv[1]) if len(sys.argv) > 1 else 37888      # 148 SMs x 256 threads
D = int(sys.argv[2]) if len(sys.argv) > 2 else 20
G = os.path.join(ROOT, "tests", "golden", "inputs", "DAV2020-full")
sim = lw.setup(lw.loadSPAC(G, "DAV2020-full", simulate_isotopes=True))
ens = en.make_ensemble(sim, M)
h = capi.Handle(ens.N, True, 0, D)
h.set_sites(ens.forcing); h.set_members(ens.site_of_member, ens.scal, ens.lay, ens.u0)
h.run()
n, ms = h.last_run_stats()
st, cnt = h.get_status()
print(f"M={M} D={D} kernel {ms:.1f} ms -> {M*D/ms*1e3:.3e} site-days/s; steps/day {cnt[:,0].mean()/D:.2f}; nf/day {cnt[:,2].mean()/D:.1f}")
