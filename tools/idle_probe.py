"""Profiling helper: lock-step waiting overhead = block iterations vs the members' own attempted steps.
Needs a library built with -DLWFB_DEBUG_ITERS (the NREJECT counter then carries the block's step iterations)."""
import os, sys, warnings
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
warnings.simplefilter("ignore")
import numpy as np
import lwfbrook90.jl_b200 as lw
from lwfbrook90.jl_b200 import capi, ensemble as en
M, D = int(sys.argv[1]), int(sys.argv[2])
G = os.path.join(ROOT, "tests", "golden", "inputs", "DAV2020-full")
sim = lw.setup(lw.loadSPAC(G, "DAV2020-full", simulate_isotopes=True))
ens = en.make_ensemble(sim, M)
h = capi.Handle(ens.N, True, 0, D, record_daily=False)
h.set_sites(ens.forcing); h.set_members(ens.site_of_member, ens.scal, ens.lay, ens.u0); h.run()
st, cnt = h.get_status(); ms = h.last_run_stats()[1]
nf = cnt[:, 2]; own = (nf - D - 1) / 6.0                 # attempted steps of the member itself (nf = 1 + days + 6 per attempt)
blk = cnt[:, 1]                                          # with -DLWFB_DEBUG_ITERS: step iterations of the member's block
print(f"M={M} D={D} kernel {ms:.0f} ms -> {M*D/ms*1e3:.3e} site-days/s")
print(f"own attempted steps/day: mean {own.mean()/D:.2f} p50 {np.median(own)/D:.2f} p99 {np.quantile(own,0.99)/D:.2f} max {own.max()/D:.2f}")
print(f"block iterations/day: mean {blk.mean()/D:.2f} max {blk.max()/D:.2f}; waiting overhead = {blk.mean()/own.mean():.3f}x")
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
np.save(os.path.join(ROOT, "gpurun_out", f"idle_own_{D}.npy"), own.astype(np.float32)); np.save(os.path.join(ROOT, "gpurun_out", f"idle_blk_{D}.npy"), blk.astype(np.float32))
