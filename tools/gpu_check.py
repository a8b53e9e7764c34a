"""Development helper: first GPU-vs-oracle comparison + raw kernel timing (run under gpurun)."""
import os, sys, time, warnings
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
warnings.simplefilter("ignore")
import numpy as np
import lwfbrook90.jl_b200 as lw
from lwfbrook90.jl_b200 import capi, ensemble as en, layout as lo
import oracle as orc

G = os.path.join(ROOT, "tests", "golden")


def compare(tag, ens, n_days, members, mode="elem"):
    orc.use(mode)
    tag = f"{tag} | oracle={mode}"
    sub = ens.subset(members)
    h = capi.Handle(ens.N, ens.iso, ens.t0_day, n_days)
    h.set_sites(ens.forcing)
    h.set_members(sub.site_of_member, sub.scal, sub.lay, sub.u0)
    h.run()
    gpu = h.get_daily()
    st, cnt = h.get_status()
    ref, st_o, cnt_o = orc.run(ens.N, ens.iso, ens.t0_day, n_days, ens.forcing, sub.site_of_member, sub.scal, sub.lay, sub.u0, n_threads=8)
    ix = lo.StateIndex(ens.N, ens.iso)
    sw_g, sw_o = gpu[:, :, ix.SWATI_mm], ref[:, :, ix.SWATI_mm]
    rel = np.abs(sw_g - sw_o) / np.abs(sw_o)
    per_member = rel.reshape(len(members), -1).max(axis=1)
    print(f"[{tag}] members={len(members)} days={n_days} status gpu={np.bincount(st)} oracle={np.bincount(st_o)}")
    print(f"   SWATI max rel err: median {np.median(per_member):.2e}  p99 {np.quantile(per_member, 0.99):.2e}  max {per_member.max():.2e}  frac<1e-6 {np.mean(per_member < 1e-6):.4f}")
    if ens.iso:
        d = np.abs(gpu[:, :, ix.SWATI_d18O] - ref[:, :, ix.SWATI_d18O]).reshape(len(members), -1).max(axis=1)
        d2 = np.abs(gpu[:, :, ix.SWATI_d2H] - ref[:, :, ix.SWATI_d2H]).reshape(len(members), -1).max(axis=1)
        print(f"   d18O max abs err: median {np.median(d):.2e} max {d.max():.2e} frac<1e-4 {np.mean(d < 1e-4):.4f};  d2H max {d2.max():.2e} frac<1e-4 {np.mean(d2 < 1e-4):.4f}")
    allerr = np.abs(gpu - ref)
    fin = np.isfinite(ref) & np.isfinite(gpu)
    print(f"   whole record: nan pattern equal {np.array_equal(np.isnan(gpu), np.isnan(ref))}, max abs err {allerr[fin].max():.3e}")
    bit = np.array([np.array_equal(gpu[i], ref[i], equal_nan=True) for i in range(len(members))])
    print(f"   BIT-EXACT members: {bit.sum()} / {len(members)}")
    if not bit.all():
        i = int(np.nonzero(~bit)[0][0])
        neq = ~((gpu[i] == ref[i]) | (np.isnan(gpu[i]) & np.isnan(ref[i])))
        d0 = int(np.nonzero(neq.any(axis=1))[0][0])
        cols = np.nonzero(neq[d0])[0][:8]
        print(f"   first mismatch: member {members[i]} day {d0} cols {cols} gpu {gpu[i, d0, cols]} ref {ref[i, d0, cols]}")
    print(f"   steps: gpu naccept mean {cnt[:,0].mean():.1f} nreject {cnt[:,1].mean():.2f}; identical counters {np.mean((cnt == cnt_o).all(axis=1)):.4f}")
    h.close()
    return per_member


def timing(tag, ens, n_days, M, record=True):
    reps = -(-M // ens.M)
    site = np.tile(ens.site_of_member, reps)[:M]
    scal = np.tile(ens.scal, (reps, 1))[:M]
    lay = np.tile(ens.lay, (reps, 1, 1))[:M]
    u0 = np.tile(ens.u0, (reps, 1))[:M]
    h = capi.Handle(ens.N, ens.iso, ens.t0_day, n_days, record_daily=record)
    h.set_sites(ens.forcing)
    t = time.time(); h.set_members(site, scal, lay, u0); t_up = time.time() - t
    h.run()
    best = 1e9
    for _ in range(2):
        h.run(); best = min(best, h.last_run_stats()[1])
    st, cnt = h.get_status()
    print(f"[{tag}] M={M} days={n_days} record={record}: kernel {best:.1f} ms -> {M * n_days / best * 1e3:.3e} site-days/s; upload {t_up*1e3:.0f} ms; steps/day {cnt[:,0].mean()/n_days:.2f}; status {np.bincount(st)}")
    h.close()


if __name__ == "__main__":
    s = lw.setup(lw.loadSPAC(os.path.join(G, "inputs/DAV2020-full"), "DAV2020-full", simulate_isotopes=True))
    ens = en.make_ensemble(s, 4096)
    compare("DAV2020-full iso N=5", ens, 364, np.arange(256))
    compare("DAV2020-full iso N=5", ens, 364, np.arange(256), mode="libm")
    sb = lw.setup(lw.loadSPAC(os.path.join(G, "inputs/BEA2016-reset-FALSE-input"), "BEA2016-reset-FALSE", simulate_isotopes=False))
    ensb = en.make_ensemble(sb, 1024)
    compare("BEA2016 noiso N=7", ensb, sb.n_days, np.arange(128))
    compare("BEA2016 noiso N=7", ensb, sb.n_days, np.arange(128), mode="libm")
    si = lw.setup(lw.loadSPAC(os.path.join(G, "inputs/isoBEA2016-reset-FALSE-input"), "isoBEA2016-reset-FALSE", simulate_isotopes=True))
    ensi = en.make_ensemble(si, 1024)
    compare("isoBEA2016 iso N=7", ensi, si.n_days, np.arange(128))
    m2 = lw.loadSPAC(os.path.join(G, "inputs/infiltrationSaturationINFEXP1"), "infiltrationSaturationINFEXP1", simulate_isotopes=False,
                     Δz_thickness_m=[0.05] * 42, IC_soil=dict(PSIM_init_kPa=-6.3, delta18O_init_permil=-13.0, delta2H_init_permil=-95.0),
                     root_distribution=dict(beta=0.90, z_rootMax_m=None))
    s2 = lw.setup(m2, requested_tspan=(0, 300))
    e2 = en.batched_setup(s2)
    compare("INFEXP1 N=42 (generic kernel), first 231 days", e2, 231, np.arange(1))
    for M in (4096, 32768):
        timing("DAV2020-full", ens, 364, M)
    timing("isoBEA2016", ensi, si.n_days, 8192)
