"""GPU diagnostic: thread-per-member kernel against the lane kernel and the oracle on isoBEA2016 (40 members, 60 days)."""
import os, sys, warnings
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
warnings.simplefilter("ignore")
import numpy as np
import lwfbrook90.jl_b200 as lw
from lwfbrook90.jl_b200 import capi, ensemble as en
import oracle as orc
G = os.path.join(ROOT, "tests", "golden", "inputs", "isoBEA2016-reset-FALSE-input")
sim = lw.setup(lw.loadSPAC(G, "isoBEA2016-reset-FALSE", simulate_isotopes=True))
ens = en.make_ensemble(sim, 40)
nd = 60
out = {}
for legacy in (False, True):
    h = capi.Handle(ens.N, True, 0, nd)
    if legacy:
        h.force_legacy(True)
    h.set_sites(ens.forcing); h.set_members(ens.site_of_member, ens.scal, ens.lay, ens.u0)
    h.run()
    out[legacy] = (h.get_daily(), *h.get_status())
    h.close()
orc.use("elem")
ref, st_o, cnt_o = orc.run(ens.N, True, 0, nd, ens.forcing, ens.site_of_member, ens.scal, ens.lay, ens.u0, n_threads=8)
for legacy in (False, True):
    rec, st, cnt = out[legacy]
    neq = ~((rec == ref) | (np.isnan(rec) & np.isnan(ref)))
    bad = np.nonzero(neq.any(axis=(1, 2)))[0]
    print("legacy" if legacy else "lanes ", "members differing from the oracle:", bad, "status", st[bad], "oracle status", st_o[bad],
          "counters equal:", np.array_equal(cnt, cnt_o))
    for m in bad[:3]:
        d0 = int(np.nonzero(neq[m].any(axis=1))[0][0])
        cols = np.nonzero(neq[m, d0])[0]
        print(f"   member {m}: first differing record {d0}, columns {cols[:12]} gpu {rec[m, d0, cols[:4]]} oracle {ref[m, d0, cols[:4]]}")
        print(f"   counters gpu {cnt[m]} oracle {cnt_o[m]}; NaN in gpu record: {np.isnan(rec[m]).any()}, in oracle: {np.isnan(ref[m]).any()}")
