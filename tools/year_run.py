"""Full-record timing: first run (given order) and second run (adaptive placement) of the same handle."""
import os, sys, warnings
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
warnings.simplefilter("ignore")
import numpy as np
import lwfbrook90.jl_b200 as lw
from lwfbrook90.jl_b200 import capi, ensemble as en
M = int(sys.argv[1]) if len(sys.argv) > 1 else 28416
D = int(sys.argv[2]) if len(sys.argv) > 2 else 365
G = os.path.join(ROOT, "tests", "golden", "inputs", "DAV2020-full")
sim = lw.setup(lw.loadSPAC(G, "DAV2020-full", simulate_isotopes=True))
ens = en.make_ensemble(sim, M)
h = capi.Handle(ens.N, True, 0, D, record_daily=os.environ.get("LWFB_RECORD", "1") == "1")
h.set_sites(ens.forcing); h.set_members(ens.site_of_member, ens.scal, ens.lay, ens.u0)
for k in range(3):
    h.run(); ms = h.last_run_stats()[1]
    print(f"run {k}: M={M} D={D} kernel {ms:.1f} ms -> {M*D/ms*1e3:.3e} site-days/s", flush=True)
