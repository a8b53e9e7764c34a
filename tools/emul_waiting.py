"""Lock-step waiting of the production kernel measured in the block emulator (CPU, no GPU needed): passes of the
blocks against the members' own attempted steps, for one launch and for work quanta with regrouping.
usage: emul_waiting.py [members] [days] [quantum,...] [warps_per_block] [config: dav|bea|isobea]
(members: multiple of 48 for DAV2020-full, of 32 for the BEA2016 configurations)"""
import os, sys, time, warnings
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests"), os.path.join(ROOT, "tests", "simt")):
    sys.path.insert(0, p)
warnings.simplefilter("ignore")
import ctypes as C
import numpy as np
import emul
import lwfbrook90.jl_b200 as lw
from lwfbrook90.jl_b200 import ensemble as en
M = int(sys.argv[1]) if len(sys.argv) > 1 else 768
D = int(sys.argv[2]) if len(sys.argv) > 2 else 30
quanta = [int(q) for q in sys.argv[3].split(",")] if len(sys.argv) > 3 else [0, 128]
cfg = sys.argv[5] if len(sys.argv) > 5 else "dav"
name, iso = {"dav": ("DAV2020-full", True), "bea": ("BEA2016-reset-FALSE", False), "isobea": ("isoBEA2016-reset-FALSE", True)}[cfg]
sim = lw.setup(lw.loadSPAC(os.path.join(ROOT, "tests", "golden", "inputs", name + "-input" if "BEA" in name else name), name, simulate_isotopes=iso))
ens = en.make_ensemble(sim, M)
L = emul.lib(); L.lwfb_emul_stage_syncs.restype = C.c_long; L.lwfb_emul_votes.restype = C.c_long
wpb = int(sys.argv[4]) if len(sys.argv) > 4 else 8       # lock-step group: warps that vote and barrier together
mpb = wpb * (32 // ens.N)
for q in quanta:
    L.lwfb_emul_stage_syncs(); L.lwfb_emul_votes()
    t = time.time()
    _, st, cnt = emul.run(emul.LANES, ens.N, ens.iso, ens.t0_day, D, ens.forcing, ens.site_of_member, ens.scal, ens.lay, ens.u0,
                          seg_len=q, resort=2 if q else 0, warps_per_block=wpb, n_threads=os.cpu_count())
    syncs = L.lwfb_emul_stage_syncs(); votes = L.lwfb_emul_votes()
    steps = (cnt[:, 0] + cnt[:, 1]).sum()
    # a step pass makes 4 block votes and 6 stage barriers (7 right after a callback: FSAL re-evaluation), a callback
    # phase 3 votes + 2 block barriers (counted with the votes); with one re-evaluation pass per callback phase:
    # syncs = 6 P_step + P_cb, votes = 4 P_step + 5 P_cb
    cb_passes = (votes - 2.0 * syncs / 3.0) / (13.0 / 3.0)
    passes = (syncs - cb_passes) / 6.0
    print(f"{wpb} warps per block, quantum {q:5d}: block passes x slots = {passes * mpb:.3e}, member steps = {steps:.3e}, "
          f"executed / needed = {passes * mpb / steps:.3f}, callback phases per block and day = "
          f"{cb_passes / (M / mpb) / D:.2f}   ({time.time() - t:.0f} s)", flush=True)
