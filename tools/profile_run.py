"""Small fixed workload for ncu: ensemble of M members x D days, one launch.
usage: profile_run.py [M] [D] [config: dav|bea|isobea]"""
import os, sys, warnings
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
warnings.simplefilter("ignore")
import numpy as np
import lwfbrook90.jl_b200 as lw
from lwfbrook90.jl_b200 import capi, ensemble as en
M = int(sys.argv[1]) if len(sys.argv) > 1 else 37888      # 148 SMs x 256 threads
D = int(sys.argv[2]) if len(sys.argv) > 2 else 20
cfg = sys.argv[3] if len(sys.argv) > 3 else "dav"
name, iso = {"dav": ("DAV2020-full", True), "davnoiso": ("DAV2020-full", False), "bea": ("BEA2016-reset-FALSE", False), "isobea": ("isoBEA2016-reset-FALSE", True)}[cfg]
folder = name + "-input" if "BEA" in name else name
G = os.path.join(ROOT, "tests", "golden", "inputs", folder)
sim = lw.setup(lw.loadSPAC(G, name, simulate_isotopes=iso))
ens = en.make_ensemble(sim, M)
h = capi.Handle(ens.N, iso, 0, D, record_daily=os.environ.get("LWFB_RECORD", "1") == "1")
h.set_sites(ens.forcing); h.set_members(ens.site_of_member, ens.scal, ens.lay, ens.u0)
h.run()
n, ms = h.last_run_stats()
st, cnt = h.get_status()
print(f"{cfg} M={M} D={D} kernel {ms:.1f} ms -> {M*D/ms*1e3:.3e} site-days/s; steps/day {cnt[:,0].mean()/D:.2f}; nf/day {cnt[:,2].mean()/D:.1f}")
