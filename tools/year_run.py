"""Full-record timing: three runs of one handle (work quanta, adaptive placement).
usage: year_run.py [M] [D] [config: dav|bea|isobea]      env LWFB_RECORD=0: no daily record"""
import os, sys, warnings
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
warnings.simplefilter("ignore")
import numpy as np
import lwfbrook90.jl_b200 as lw
from lwfbrook90.jl_b200 import capi, ensemble as en
M = int(sys.argv[1]) if len(sys.argv) > 1 else 28416
D = int(sys.argv[2]) if len(sys.argv) > 2 else 365
cfg = sys.argv[3] if len(sys.argv) > 3 else "dav"
name, iso = {"dav": ("DAV2020-full", True), "bea": ("BEA2016-reset-FALSE", False), "isobea": ("isoBEA2016-reset-FALSE", True)}[cfg]
G = os.path.join(ROOT, "tests", "golden", "inputs", name + "-input" if "BEA" in name else name)
sim = lw.setup(lw.loadSPAC(G, name, simulate_isotopes=iso))
ens = en.make_ensemble(sim, M)
h = capi.Handle(ens.N, iso, 0, D, record_daily=os.environ.get("LWFB_RECORD", "1") == "1")
h.set_sites(ens.forcing); h.set_members(ens.site_of_member, ens.scal, ens.lay, ens.u0)
for k in range(int(os.environ.get("LWFB_RUNS", "3"))):
    h.set_placement(1)            # cold: nothing learned from the previous run is reused
    h.run(); n, ms = h.last_run_stats()
    print(f"{cfg} run {k}: M={M} D={D} kernel {ms:.1f} ms ({n} launches) -> {M*D/ms*1e3:.3e} site-days/s", flush=True)
