"""Development helper: kernel text under the SIMT emulator vs the CPU oracle (elem mode), bit-for-bit."""
import os, sys, time, warnings
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests"), os.path.join(ROOT, "tests", "simt")):
    sys.path.insert(0, p)
warnings.simplefilter("ignore")
import numpy as np
import lwfbrook90.jl_b200 as lw
from lwfbrook90.jl_b200 import ensemble as en, layout as lo
import oracle as orc
import emul
from conftest import inputs

def compare(tag, ens, nd, kernel, M=None):
    sub = ens if M is None else ens.subset(np.arange(M))
    orc.use("elem")
    t = time.time()
    ref, st_o, cnt_o = orc.run(sub.N, sub.iso, sub.t0_day, nd, sub.forcing, sub.site_of_member, sub.scal, sub.lay, sub.u0, n_threads=8)
    t1 = time.time()
    g, st, cnt = emul.run(kernel, sub.N, sub.iso, sub.t0_day, nd, sub.forcing, sub.site_of_member, sub.scal, sub.lay, sub.u0)
    t2 = time.time()
    same = (g == ref) | (np.isnan(g) & np.isnan(ref))
    bad = np.nonzero(~same.all(axis=(1, 2)))[0]
    print(f"[{tag}] kernel={kernel} M={sub.M} days={nd}: bit-exact {sub.M - len(bad)}/{sub.M}; status eq {np.array_equal(st, st_o)}; counters eq {np.array_equal(cnt, cnt_o)}; oracle {t1-t:.1f}s emul {t2-t1:.1f}s")
    if len(bad):
        i = bad[0]
        days = np.nonzero((~same[i]).any(axis=1))[0]
        d0 = days[0]; cols = np.nonzero(~same[i, d0])[0]
        print(f"   member {i}: first mismatch day {d0}, cols {cols[:12]} ({len(cols)} cols)")
        print("   emul", g[i, d0, cols[:6]]); print("   orc ", ref[i, d0, cols[:6]])
        print("   counters emul", cnt[i], "orc", cnt_o[i], "status", st[i], st_o[i])
    return len(bad) == 0

if __name__ == "__main__":
    which = sys.argv[1] if len(sys.argv) > 1 else "dav"
    kernel = int(sys.argv[2]) if len(sys.argv) > 2 else 0
    nd = int(sys.argv[3]) if len(sys.argv) > 3 else 5
    M = int(sys.argv[4]) if len(sys.argv) > 4 else 12
    if which == "dav":
        s = lw.setup(lw.loadSPAC(inputs("DAV2020-full"), "DAV2020-full", simulate_isotopes=True))
    elif which == "bea":
        s = lw.setup(lw.loadSPAC(inputs("BEA2016-reset-FALSE-input"), "BEA2016-reset-FALSE", simulate_isotopes=False))
    elif which == "isobea":
        s = lw.setup(lw.loadSPAC(inputs("isoBEA2016-reset-FALSE-input"), "isoBEA2016-reset-FALSE", simulate_isotopes=True))
    ens = en.make_ensemble(s, M)
    compare(which, ens, nd, kernel)
