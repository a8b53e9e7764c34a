"""Deep profiles: the warp-per-member kernel (N > 32) against the thread-per-member kernel on INFEXP1 (N = 42).
usage: wide_run.py [M] [D] [layers, default 42]"""
import os, sys, warnings
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
warnings.simplefilter("ignore")
import numpy as np
import lwfbrook90.jl_b200 as lw
from lwfbrook90.jl_b200 import capi, ensemble as en
M = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
D = int(sys.argv[2]) if len(sys.argv) > 2 else 60
G = os.path.join(ROOT, "tests", "golden", "inputs", "infiltrationSaturationINFEXP1")
NL = int(sys.argv[3]) if len(sys.argv) > 3 else 42
m = lw.loadSPAC(G, "infiltrationSaturationINFEXP1", simulate_isotopes=False, Δz_thickness_m=[2.1 / NL] * NL,
                IC_soil=dict(PSIM_init_kPa=-6.3, delta18O_init_permil=-13.0, delta2H_init_permil=-95.0),
                root_distribution=dict(beta=0.90, z_rootMax_m=None))
sim = lw.setup(m, requested_tspan=(0, D))
ens = en.batched_setup(sim, params={"DRAIN": np.linspace(0.1, 1.0, M)})
print("N =", ens.N)
res = {}
for legacy in ((0, 1) if ens.N <= 64 else (0,)):
    h = capi.Handle(ens.N, ens.iso, 0, D, record_daily=True)
    h.force_legacy(legacy)
    h.set_sites(ens.forcing); h.set_members(ens.site_of_member, ens.scal, ens.lay, ens.u0)
    for k in range(2):
        h.run(); n, ms = h.last_run_stats()
    res[legacy] = h.get_final()
    print(f"{'thread-per-member' if legacy else 'warp-per-member  '} kernel: M={M} D={D} N={ens.N} {ms:.1f} ms -> {M*D/ms*1e3:.3e} site-days/s", flush=True)
    h.close()
if 1 in res:
    print("same bits:", np.array_equal(res[0], res[1], equal_nan=True))
