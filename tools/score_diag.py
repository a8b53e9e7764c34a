"""GPU diagnostic: scores accumulated during the run (lane kernel, thread-per-member kernel) against the scores formed
from the resident record, column by column."""
import os, sys, warnings
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
warnings.simplefilter("ignore")
import numpy as np
import lwfbrook90.jl_b200 as lw
from lwfbrook90.jl_b200 import capi, ensemble as en
G = os.path.join(ROOT, "tests", "golden", "inputs", "isoBEA2016-reset-FALSE-input")
sim = lw.setup(lw.loadSPAC(G, "isoBEA2016-reset-FALSE", simulate_isotopes=True))
ens = en.make_ensemble(sim, 40)
nd = 60
obs_day = np.array([0, 3, 10, 11, 30, 59, 60], dtype=np.int32)
obs_layer = np.array([0, 3, 6], dtype=np.int32)
rng = np.random.default_rng(11)
th = 0.25 + 0.05 * rng.random((len(obs_day), 3)); d18 = -10 + rng.random((len(obs_day), 3))
res = {}
for name, legacy, record in (("record/lanes", False, True), ("record/legacy", True, True), ("run/lanes", False, False), ("run/legacy", True, False)):
    h = capi.Handle(ens.N, True, 0, nd, record_daily=record)
    if legacy:
        h.force_legacy(True)
    h.set_sites(ens.forcing); h.set_members(ens.site_of_member, ens.scal, ens.lay, ens.u0)
    if not record:
        h.set_observations(obs_day, obs_layer, th, d18)
    h.run()
    res[name] = h.scores(obs_day, obs_layer, th, d18) if record else h.scores()
    h.close()
ref = res["record/lanes"]
for k, v in res.items():
    neq = ~((v == ref) | (np.isnan(v) & np.isnan(ref)))
    print(k, "mismatching entries per column:", neq.sum(axis=0), "max rel diff", np.nanmax(np.abs(v - ref) / np.abs(ref)))
