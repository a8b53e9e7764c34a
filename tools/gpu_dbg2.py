import os, sys, warnings
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle")); sys.path.insert(0, os.path.join(ROOT, "tests"))
warnings.simplefilter("ignore")
import numpy as np
import lwfbrook90.jl_b200 as lw
from lwfbrook90.jl_b200 import capi, ensemble as en, layout as lo
import oracle as orc
from conftest import inputs
orc.use("elem")
m = lw.loadSPAC(inputs("infiltrationSaturationINFEXP1"), "infiltrationSaturationINFEXP1", simulate_isotopes=False,
                Δz_thickness_m=[0.05] * 42, IC_soil=dict(PSIM_init_kPa=-6.3, delta18O_init_permil=-13.0, delta2H_init_permil=-95.0),
                root_distribution=dict(beta=0.90, z_rootMax_m=None))
s = lw.setup(m, requested_tspan=(0, 300))
e = en.batched_setup(s, params={"DRAIN": np.array([1.0, 0.6, 0.3])})
nd = 260
h = capi.Handle(e.N, e.iso, 0, nd); h.set_sites(e.forcing); h.set_members(e.site_of_member, e.scal, e.lay, e.u0); h.run()
g = h.get_daily(); st, cnt = h.get_status()
r, st_o, cnt_o = orc.run(e.N, e.iso, 0, nd, e.forcing, e.site_of_member, e.scal, e.lay, e.u0)
print("status", st, st_o); print("cnt gpu", cnt.tolist()); print("cnt orc", cnt_o.tolist())
for i in range(3):
    neq = ~((g[i] == r[i]) | (np.isnan(g[i]) & np.isnan(r[i])))
    days = np.nonzero(neq.any(axis=1))[0]
    print("member", i, "first mismatch day", days[:3], "cols", np.nonzero(neq[days[0]])[0][:10] if len(days) else None)
    if len(days):
        d0 = days[0]; c = np.nonzero(neq[d0])[0][:5]; print("   gpu", g[i, d0, c], "orc", r[i, d0, c])
ix = lo.StateIndex(e.N, e.iso)
np.set_printoptions(linewidth=220, precision=4)
print("gpu SWATI day1", g[0, 1, ix.SWATI_mm]); print("orc SWATI day1", r[0, 1, ix.SWATI_mm])
print("gpu TRANI day1", g[0, 1, ix.TRANI_mm]); print("orc TRANI day1", r[0, 1, ix.TRANI_mm])
print("gpu scal", g[0, 1, ix.GWAT_mm:ix.TRANI_mm.start]); print("orc scal", r[0, 1, ix.GWAT_mm:ix.TRANI_mm.start])
print("gpu acc", g[0, 1, ix.accum]); print("orc acc", r[0, 1, ix.accum])
