"""Cold and warm run time of lwfb_run against the work quantum (LWFB_QUANTUM; 0 = single launch).
usage: quantum_sweep.py dav|bea|isobea M q1,q2,... [record]"""
import os, sys, warnings
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
warnings.simplefilter("ignore")
import numpy as np
import lwfbrook90.jl_b200 as lw
from lwfbrook90.jl_b200 import capi, ensemble as en
cfg, M = sys.argv[1], int(sys.argv[2])
qs = [tuple(int(x) for x in (q + ":50").split(":")[:2]) for q in sys.argv[3].split(",")]   # quantum[:prio_pct]
record = len(sys.argv) > 4 and sys.argv[4] == "record"
name, iso = {"dav": ("DAV2020-full", True), "bea": ("BEA2016-reset-FALSE", False), "isobea": ("isoBEA2016-reset-FALSE", True)}[cfg]
folder = name + "-input" if "BEA" in name else name
sim = lw.setup(lw.loadSPAC(os.path.join(ROOT, "tests", "golden", "inputs", folder), name, simulate_isotopes=iso))
ens = en.make_ensemble(sim, M)
D = ens.n_days
ref = None
for q, pct in qs:
    os.environ["LWFB_QUANTUM"] = str(q)
    os.environ["LWFB_PRIO_PCT"] = str(pct)
    h = capi.Handle(ens.N, iso, 0, D, record_daily=record)
    h.set_sites(ens.forcing); h.set_members(ens.site_of_member, ens.scal, ens.lay, ens.u0)
    for k in range(2):
        h.run(); n, ms = h.last_run_stats()
        print(f"{cfg} M={M} D={D} quantum {q:6d} prio {pct:2d}% {'cold' if k == 0 else 'warm'}: {n:4d} launches, device {ms:8.0f} ms -> "
              f"{M*D/ms*1e3:.3e} site-days/s", flush=True)
    fin = h.get_final(); st, cnt = h.get_status()
    if ref is None:
        ref = (fin, st, cnt)
    else:
        ok = np.array_equal(fin, ref[0], equal_nan=True) and np.array_equal(st, ref[1]) and np.array_equal(cnt, ref[2])
        print("   final state, status, counters identical to the first setting:", ok, flush=True)
    h.close()
