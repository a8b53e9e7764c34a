"""Distribution of per-member step counts of an ensemble (GPU run, no record)."""
import os, sys, warnings
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
warnings.simplefilter("ignore")
import numpy as np
import lwfbrook90.jl_b200 as lw
from lwfbrook90.jl_b200 import capi, ensemble as en
cfg, M = sys.argv[1], int(sys.argv[2])
name, iso = {"dav": ("DAV2020-full", True), "bea": ("BEA2016-reset-FALSE", False), "isobea": ("isoBEA2016-reset-FALSE", True)}[cfg]
folder = name + "-input" if "BEA" in name else name
sim = lw.setup(lw.loadSPAC(os.path.join(ROOT, "tests", "golden", "inputs", folder), name, simulate_isotopes=iso))
ens = en.make_ensemble(sim, M)
D = ens.n_days
h = capi.Handle(ens.N, iso, 0, D, record_daily=False)
h.set_sites(ens.forcing); h.set_members(ens.site_of_member, ens.scal, ens.lay, ens.u0)
for k in range(2):
    h.run(); ms = h.last_run_stats()[1]
    print(f"run {k}: {cfg} M={M} D={D} kernel {ms:.0f} ms -> {M*D/ms*1e3:.3e} site-days/s", flush=True)
st, cnt = h.get_status()
att = (cnt[:, 0] + cnt[:, 1]) / D
print("status counts", np.bincount(st), "attempted steps/day quantiles [min 10% 50% 90% 99% 99.9% max]:",
      np.round(np.quantile(att, [0, .1, .5, .9, .99, .999, 1]), 1))
print("share of all steps taken by the top 1% members:", np.sort(att)[-max(1, M // 100):].sum() / att.sum())
