"""GPU triage (run under gpurun): both kernels vs the oracle (elem mode) on small cases + raw timing."""
import os, sys, time, warnings
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
warnings.simplefilter("ignore")
import numpy as np
import lwfbrook90.jl_b200 as lw
from lwfbrook90.jl_b200 import capi, ensemble as en, layout as lo
import oracle as orc
from conftest import inputs

def compare(tag, ens, nd, legacy, M=None):
    sub = ens if M is None else ens.subset(np.arange(M))
    orc.use("elem")
    ref, st_o, cnt_o = orc.run(sub.N, sub.iso, sub.t0_day, nd, sub.forcing, sub.site_of_member, sub.scal, sub.lay, sub.u0, n_threads=8)
    h = capi.Handle(sub.N, sub.iso, sub.t0_day, nd)
    h.force_legacy(legacy)
    h.set_sites(sub.forcing); h.set_members(sub.site_of_member, sub.scal, sub.lay, sub.u0); h.run()
    g = h.get_daily(); st, cnt = h.get_status(); ms = h.last_run_stats()[1]; h.close()
    same = (g == ref) | (np.isnan(g) & np.isnan(ref))
    bad = np.nonzero(~same.all(axis=(1, 2)))[0]
    print(f"[{tag}] legacy={legacy} M={sub.M} N={sub.N} days={nd}: bit-exact {sub.M - len(bad)}/{sub.M}; status eq {np.array_equal(st, st_o)}; counters eq {np.array_equal(cnt, cnt_o)}; kernel {ms:.1f} ms", flush=True)
    if len(bad):
        i = bad[0]
        days = np.nonzero((~same[i]).any(axis=1))[0]
        d0 = days[0]; cols = np.nonzero(~same[i, d0])[0]
        print(f"   member {i}: first mismatch day {d0}, cols {cols[:12]} ({len(cols)} cols)")
        print("   gpu", g[i, d0, cols[:6]]); print("   orc", ref[i, d0, cols[:6]])
        print("   counters gpu", cnt[i], "orc", cnt_o[i], "status", st[i], st_o[i], flush=True)

def timing(tag, ens, nd, M, legacy=False, record=True):
    reps = -(-M // ens.M)
    site = np.tile(ens.site_of_member, reps)[:M]; scal = np.tile(ens.scal, (reps, 1))[:M]
    lay = np.tile(ens.lay, (reps, 1, 1))[:M]; u0 = np.tile(ens.u0, (reps, 1))[:M]
    h = capi.Handle(ens.N, ens.iso, ens.t0_day, nd, record_daily=record)
    h.force_legacy(legacy)
    h.set_sites(ens.forcing); h.set_members(site, scal, lay, u0)
    h.run(); best = h.last_run_stats()[1]
    h.run(); best = min(best, h.last_run_stats()[1])
    st, cnt = h.get_status()
    print(f"[{tag}] legacy={legacy} M={M} days={nd}: kernel {best:.1f} ms -> {M * nd / best * 1e3:.3e} site-days/s; steps/day {cnt[:,0].mean()/nd:.2f}; status {np.bincount(st)}", flush=True)
    h.close()

if __name__ == "__main__":
    s = lw.setup(lw.loadSPAC(inputs("DAV2020-full"), "DAV2020-full", simulate_isotopes=True))
    ens = en.make_ensemble(s, 4096)
    for legacy in (False, True):
        compare("dav", ens, 5, legacy, 12)
        compare("dav", ens, 120, legacy, 96)
    sb = lw.setup(lw.loadSPAC(inputs("BEA2016-reset-FALSE-input"), "BEA2016-reset-FALSE", simulate_isotopes=False))
    eb = en.make_ensemble(sb, 48)
    for legacy in (False, True):
        compare("bea", eb, 60, legacy)
    timing("dav", ens, 60, 37888)
    timing("dav", ens, 60, 4096, legacy=True)
