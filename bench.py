#!/usr/bin/env python
"""bench.py — site-days/s of the batched LWFBrook90 integrator on B200 (contract: see the task brief).

A "step" = one pass of the hot path over one batch: every member of this rank's shard of the ensemble integrated over
the full record, scored on the GPU against synthetic observations (sums accumulated by the integrator itself at the
observation days: no daily record is kept, so the whole 1,022,976-member DAV2020-full ensemble fits ONE GPU), scores
all-gathered across ranks (NCCL) when N > 1.  Default = the metric's own configuration, the FIXED 1,022,976-member
ensemble split over the ranks (strong scaling); `--members M` runs M members per GPU instead (weak scaling).
Every step starts COLD: nothing learned about the members' cost in earlier runs is reused (lwfb_set_placement).

  value  site-days/s with the member blocks already resident in HBM (device-synchronised wall time of K steps)
  e2e    the same through the public API with HOST buffers, every step: the members' knob rows (what remakeSPAC would be
         given) from pinned host memory, batched setup of the member blocks ON THE GPU (lwfb_setup_members), run, D2H
         of scores + status + counters
  e2e_daily      bounded extra leg: the drop-in simulate! contract (daily record kept in HBM and downloaded with
                 lwfb_get_daily into pinned host memory) on a sub-ensemble
  roofline       FP64-pipe roofline of the integrate kernel (the path is FP64-bound)
  roofline_hbm   the secondary HBM figure (algorithmic bytes 8*(L+F) per site-day)
  cpu_baseline   the CPU oracle (restatement of the reference algorithm) on a bounded sample, all cores and one core
  parity         the first 256 members of rank 0 re-integrated by the CPU oracle (device arithmetic): final state,
                 status and step counters must be bit-identical

`--config bea2016 | isobea2016` runs BASELINE.json's configs 3 and 4 (10,000 / 100,000 members on one B200) with the
same line.  `--impl reference` times the CPU oracle (the reference itself is Julia and cannot run here) with all host
threads on the same workload definition.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time
import warnings

ROOT = os.path.dirname(os.path.abspath(__file__))
for p in (ROOT, os.path.join(ROOT, "oracle")):
    if p not in sys.path:
        sys.path.insert(0, p)
warnings.simplefilter("ignore")

import numpy as np  # noqa: E402

METRIC = "site-days/sec, 1M-member DAV2020-full iso ensemble at 1/2/4/8 B200 vs CPU"
# name -> (input folder, prefix, isotopes, members of the whole ensemble, what BASELINE.json calls it)
CONFIGS = {
    # 72 full waves of the lane kernel on one B200 (148 SMs x 2 blocks x 8 warps x 6 members = 14208 members per wave)
    "dav2020": ("DAV2020-full", "DAV2020-full", True, 1022976, "DAV2020-full isotope ensemble (configs[4], the metric's own)"),
    "bea2016": ("BEA2016-reset-FALSE-input", "BEA2016-reset-FALSE", False, 10000, "BEA2016-reset-FALSE parameter ensemble (configs[2])"),
    "isobea2016": ("isoBEA2016-reset-FALSE-input", "isoBEA2016-reset-FALSE", True, 100000, "isoBEA2016 isotope ensemble (configs[3])"),
}

# FLOP-equivalents per site-day (SURVEY 8d convention: add/mul/div/min/max = 1, FMA = 2, sqrt = 4,
# exp/log/trig = 20, pow = 45; only live arithmetic counts)
C_FL, C_FS, C_AL, C_AS = 380.0, 25.0, 145.0, 80.0


def flop_eq(N, iso, n_days, counters):
    nf = counters[:, 2].astype(np.float64)
    nacc = counters[:, 0].astype(np.float64)
    w = nf * (N * C_FL + C_FS) + (nacc * (N * C_AL + C_AS) if iso else 0.0) + n_days * (4000.0 + 60.0 * N)
    return float(w.sum())


class ClockSampler:
    """nvidia-smi clocks + throttle reasons during the timed region."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu, self.rows, self.proc = gpu_index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "200", "-i", str(self.gpu)], stdout=subprocess.PIPE, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if self.proc is None:
            return None
        self.proc.terminate()
        time.sleep(0.05)
        sm = [float(r[1]) for r in self.rows if len(r) >= 9 and r[1].replace(".", "").isdigit()]
        if not sm:
            return None
        mx = [float(r[2]) for r in self.rows if len(r) >= 9 and r[2].replace(".", "").isdigit()]
        reasons = set()
        for r in self.rows:
            if len(r) >= 9:
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)) if mx else None, "reasons": sorted(reasons),
                "samples": len(sm)}


def load_site(config):
    import lwfbrook90.jl_b200 as lw
    folder, prefix, iso, _, _ = CONFIGS[config]
    G = os.path.join(ROOT, "tests", "golden", "inputs", folder)
    return lw.setup(lw.loadSPAC(G, prefix, simulate_isotopes=iso))


def build_workload(config, members, first_member=0):
    """Host-built member blocks (numpy batched setup): the CPU legs and the reference arm."""
    from lwfbrook90.jl_b200 import ensemble as en
    sim = load_site(config)
    ens = en.make_ensemble(sim, members, seed=1234, first_member=first_member)
    return sim, ens


def synthetic_observations(sim):
    """theta / delta18O 'observations' for the scores: weekly, layers 1, 3, 5 (0-based 0, 2, 4), taken from
    the unmodified site definition's initial profile plus a seasonal signal (values only matter for the scores)."""
    nd = sim.n_days
    days = np.arange(7, nd, 7, dtype=np.int32)
    layers = np.array([0, 2, 4], dtype=np.int32)
    th = 0.2 + 0.03 * np.sin(2 * np.pi * days[:, None] / 365.0) + 0.01 * layers[None, :]
    d18 = -12.0 + 3.0 * np.sin(2 * np.pi * (days[:, None] - 60) / 365.0) + 0.2 * layers[None, :]
    return days, layers, th, d18


def cpu_oracle_rate(ens, n_threads, target_s=15.0):
    import oracle as orc
    orc.use("libm")
    probe = min(ens.M, 4 * n_threads)
    orc.run(ens.N, ens.iso, ens.t0_day, 3, ens.forcing, ens.site_of_member[:probe], ens.scal[:probe],
            ens.lay[:probe], ens.u0[:probe], n_threads=n_threads, record=False)      # warm-up (library load, thread start)
    t = time.perf_counter()
    orc.run(ens.N, ens.iso, ens.t0_day, ens.n_days, ens.forcing, ens.site_of_member[:probe], ens.scal[:probe],
            ens.lay[:probe], ens.u0[:probe], n_threads=n_threads, record=False)
    dt = time.perf_counter() - t
    n = int(min(ens.M, max(probe, probe * target_s / max(dt, 1e-3))))
    t = time.perf_counter()
    _, st, cnt = orc.run(ens.N, ens.iso, ens.t0_day, ens.n_days, ens.forcing, ens.site_of_member[:n], ens.scal[:n],
                         ens.lay[:n], ens.u0[:n], n_threads=n_threads, record=False)
    dt = time.perf_counter() - t
    return n * ens.n_days / dt, n, dt


def workload_text(config, ens, M, world, total, scaling, extra=""):
    _, _, iso, _, what = CONFIGS[config]
    L = ens.u0.shape[1]
    return (f"{what}: {total:,} members in all = {M:,} per GPU x {world} GPU(s) ({scaling} scaling), N={ens.N} layers, "
            f"{ens.n_days} days, state length {L}, {'with' if iso else 'without'} isotopes, Philox seed 1234{extra}")


def run_reference(args):
    """--impl reference: the CPU restatement of the reference algorithm on the host cores."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    sample = args.ref_members
    sim, ens = build_workload(args.config, sample)
    import oracle as orc
    orc.use("libm")

    def step():
        t = time.perf_counter()
        orc.run(ens.N, ens.iso, ens.t0_day, ens.n_days, ens.forcing, ens.site_of_member, ens.scal, ens.lay, ens.u0,
                n_threads=cores, record=False)
        return time.perf_counter() - t
    for _ in range(args.warmup):
        step()
    ts = [step() for _ in range(args.steps)]
    ms = 1e3 * float(np.mean(ts))
    val = ens.M * ens.n_days / (ms * 1e-3)
    total = CONFIGS[args.config][3]
    line = {"impl": "reference", "metric": METRIC, "value": val, "unit": "site-days/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True,
            "scaling": "weak" if args.members else "strong",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": workload_text(args.config, ens, total, 1, total, "n/a") +
                                   f"; bounded sample of {ens.M} members per step on the host cores"},
            "cpu_baseline": {"value": val, "unit": "site-days/s", "cores": cores, "kind": "port",
                             "sample": f"first {ens.M} members x {ens.n_days} days, CPU restatement of the reference "
                                       "algorithm (oracle/lwfb90_oracle.c, host libm); the Julia reference cannot run here"},
            "e2e": {"value": val, "unit": "site-days/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


def parity_check(kb, h, n_check=256):
    """First n_check members of this rank: final state, status and counters from the GPU (members set up on the device)
    against the CPU oracle in device arithmetic, from member blocks built by the CPU twin of the device setup
    (bit-identical or not)."""
    import oracle as orc
    n = min(n_check, kb.M)
    fin = h.get_final(0, n)
    st, cnt = h.get_status()
    scal, lay, u0, _ = orc.setup_members(kb.N, kb.iso, kb.base_scal, kb.base_lay, kb.base_u0, kb.knobs[:n], kb.over_idx,
                                         kb.profiles, kb.psim_base, kb.ilayer)
    orc.use("elem")
    ref, st_o, cnt_o = orc.run(kb.N, kb.iso, kb.t0_day, kb.n_days, kb.forcing, np.zeros(n, np.int32), scal, lay, u0,
                               n_threads=os.cpu_count() or 1)
    last = ref[:, -1, :]
    same = np.array_equal(fin, last, equal_nan=True) and np.array_equal(st[:n], st_o) and np.array_equal(cnt[:n], cnt_o)
    return bool(same), n


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200")
    ap.add_argument("--config", default="dav2020", choices=sorted(CONFIGS))
    ap.add_argument("--members", type=int, default=0, help="members PER GPU (weak scaling); default: the config's whole "
                                                            "ensemble split over the GPUs (strong scaling)")
    ap.add_argument("--total-members", type=int, default=0, help="size of the whole ensemble (default: the config's)")
    ap.add_argument("--record", type=int, default=0, help="1: keep the daily record in HBM during the timed steps")
    ap.add_argument("--daily-members", type=int, default=16384, help="members of the e2e_daily leg (0: skip)")
    ap.add_argument("--e2e-steps", type=int, default=2)
    ap.add_argument("--ref-members", type=int, default=2048, help="members per step of the reference arm")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-parity", action="store_true")
    ap.add_argument("--balance", type=int, default=0, help="N > 1, strong scaling: days of a probe run whose RHS counts deal the "
                    "members to the ranks by cost (snake deal of the cost-sorted members) instead of contiguous shards; 0: off")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)

    import torch
    from lwfbrook90.jl_b200 import capi, layout as lo
    from lwfbrook90.jl_b200.parallel import DeviceArray, allgather_by_members, allgather_rows, allgather_scores, balanced_deal, shard_range

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    dist = None
    torch.cuda.set_device(local_rank)
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", local_rank))
    iso = CONFIGS[args.config][2]
    if args.members:                                   # weak scaling: fixed work per GPU
        M, first, scaling = args.members, rank * args.members, "weak"
        total = M * world
    else:                                              # strong scaling: the fixed ensemble, contiguous shards
        total = args.total_members or CONFIGS[args.config][3]
        lo_m, hi_m = shard_range(total, rank, world)
        M, first, scaling = hi_m - lo_m, lo_m, "strong"
    from lwfbrook90.jl_b200 import ensemble as en
    t_gen = time.perf_counter()
    sim = load_site(args.config)
    kb = en.make_knob_block(sim, M, seed=1234, first_member=first)     # each GPU integrates its own slice of the ensemble:
    t_gen = time.perf_counter() - t_gen                                # Philox knob rows on the host, setup on the GPU
    member_ids = None                                                  # (contiguous shard: first .. first + M - 1)
    balance_note = None
    if args.balance > 0 and world > 1 and not args.members:
        # cost-balanced shards: a short probe run of the contiguous shard, its RHS counts all-gathered, the
        # members sorted by cost and dealt to the ranks in a snake.  The knob generator is counter-based: every rank can
        # produce any member, nothing has to move.  Outside the timed region (a calibration loop does it once).
        t_bal = time.perf_counter()
        hp = capi.Handle(kb.N, iso, kb.t0_day, min(args.balance, kb.n_days), device=local_rank, record_daily=False)
        hp.set_sites(kb.forcing); hp.setup_members(kb); hp.run()
        _, cnt_p = hp.get_status()
        hp.close()
        cost_blocks = allgather_rows(torch.from_numpy(cnt_p[:, 2:3].astype(np.float64)).to(f"cuda:{local_rank}"), rank, world)
        cost = torch.cat(cost_blocks, dim=0)[:, 0].cpu().numpy()           # contiguous shards in rank order = global member order
        deal = balanced_deal(cost, world)
        member_ids = deal[rank]
        M = len(member_ids)
        kb = en.make_knob_block(sim, M, seed=1234, members=member_ids)
        balance_note = {"probe_days": int(min(args.balance, kb.n_days)), "seconds": time.perf_counter() - t_bal,
                        "cost_share_of_rank_max_over_mean": float(max(cost[d_].sum() for d_ in deal) / (cost.sum() / world)),
                        "contiguous_would_be": float(max(cost[slice(*shard_range(total, r_, world))].sum() for r_ in range(world)) / (cost.sum() / world))}
    ens = en.make_ensemble(sim, min(M, 8192), seed=1234, first_member=first)   # host-built blocks of the CPU legs
    N, L, nd = kb.N, kb.base_u0.shape[0], kb.n_days
    obs_day, obs_layer, obs_th, obs_d18 = synthetic_observations(sim)
    if not iso:
        obs_d18 = None

    # pinned host copy of the knob rows (the e2e leg uploads them every step)
    knobs_t = torch.from_numpy(np.ascontiguousarray(kb.knobs)).pin_memory()
    kb.knobs = knobs_t.numpy()

    record = bool(args.record)
    h = capi.Handle(N, iso, kb.t0_day, nd, device=local_rank, record_daily=record)
    h.set_sites(kb.forcing)
    h.setup_members(kb)                                         # member blocks built on the device
    h.set_observations(obs_day, obs_layer, obs_th, obs_d18)     # scored by the integrator itself: no record needed
    fp64_peak = capi.fp64_peak(local_rank)

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    def gather(ptr):
        if dist is None:
            return None
        t = torch.as_tensor(DeviceArray(ptr, (M, lo.NSCORES)), device=f"cuda:{local_rank}")
        if member_ids is not None:
            return allgather_by_members(t, member_ids, total, rank, world)
        return allgather_scores(t, total, rank, world)

    # our kernels launched per step: n integrator launches, one placement_keys kernel after each of them (regrouping),
    # one score kernel; cub's radix-sort kernels are not counted
    launches = []

    def step_resident():
        h.set_placement(1)                                    # cold: forget what the previous step learned
        h.run()
        ptr = h.scores_device()
        gather(ptr)
        n, ms = h.last_run_stats()
        launches.append(2 * n + 1 if n > 1 else n + 1)
        return ms

    def step_e2e():
        h.setup_members(kb)                                   # H2D of the knob rows + batched setup on the GPU (cold run)
        h.run()
        sc = h.scores()                                       # D2H of the result
        st, cnt = h.get_status()
        if dist is not None:
            t = torch.from_numpy(sc).to(f"cuda:{local_rank}")
            if member_ids is not None:
                allgather_by_members(t, member_ids, total, rank, world)
            else:
                allgather_scores(t, total, rank, world)
        return sc, st, cnt

    for _ in range(args.warmup):
        step_resident()
    sampler = ClockSampler(local_rank)
    sampler.start()
    barrier()
    t0 = time.perf_counter()
    launches.clear()
    kernel_ms = [step_resident() for _ in range(args.steps)]
    n_launched = int(sum(launches))
    barrier()
    wall = time.perf_counter() - t0
    clocks = sampler.stop()
    st, cnt = h.get_status()
    parity = None
    if rank == 0 and not args.no_parity:
        parity = parity_check(kb, h)

    barrier()
    t0 = time.perf_counter()
    n_e2e = max(1, args.e2e_steps)
    for _ in range(n_e2e):
        sc, st2, cnt2 = step_e2e()
    barrier()
    wall_e2e = (time.perf_counter() - t0) / n_e2e
    h.close()

    # bounded leg with the daily record: what the reference's simulate! returns (saveat = daily) brought to the host
    daily = None
    if args.daily_members > 0:
        Md = min(M, args.daily_members)
        hd = capi.Handle(N, iso, kb.t0_day, nd, device=local_rank, record_daily=True)
        hd.set_sites(kb.forcing)
        kbd = en.KnobBlock(**{**kb.__dict__, "knobs": kb.knobs[:Md]})
        out_t = torch.empty((Md, nd + 1, L), dtype=torch.float64).pin_memory()
        out_h = out_t.numpy()

        def step_daily():
            hd.setup_members(kbd)
            hd.run()
            hd.get_daily(out=out_h)
            return hd.get_status()
        step_daily()
        barrier()
        t0 = time.perf_counter()
        step_daily()
        torch.cuda.synchronize()
        t_daily = time.perf_counter() - t0
        daily = (Md, t_daily)
        hd.close()

    # max over ranks
    def allmax(x):
        if dist is None:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=f"cuda:{local_rank}")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def allsum(x):
        if dist is None:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=f"cuda:{local_rank}")
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return float(t.item())
    wall = allmax(wall)
    wall_e2e = allmax(wall_e2e)
    kms = allmax(float(np.mean(kernel_ms)))
    flops = allsum(flop_eq(N, iso, nd, cnt)) / world          # mean per GPU
    n_ok = allsum(float((st == 0).sum()))
    steps_acc = allsum(float(cnt[:, 0].sum())) / (total * nd)
    steps_att = allsum(float((cnt[:, 0] + cnt[:, 1]).sum())) / (total * nd)
    nf_day = allsum(float(cnt[:, 2].sum())) / (total * nd)
    daily_rate = None
    if daily is not None:
        daily_rate = allsum(daily[0]) * nd / allmax(daily[1])

    if rank == 0:
        site_days = float(total) * nd
        ms_per_step = 1e3 * wall / args.steps
        value = site_days / (ms_per_step * 1e-3)
        e2e_val = site_days / wall_e2e
        h2d = kb.knobs.nbytes + (lo.NPS + lo.NPL * N + L) * 8         # knob rows + the base member
        d2h = M * lo.NSCORES * 8 + M * 4 + M * lo.NCOUNTERS * 8
        ach_tf = flops / (kms * 1e-3) / 1e12
        alg_bytes = float(M) * nd * 8 * (L + lo.NF)
        blocks_mb = M * (lo.NPS + lo.NPL * N + L) * 8 / 1e6
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
        traffic = traffic_detail = None
        try:                                          # DRAM bytes of the committed ncu capture (per site-day), scaled to one step
            tr = json.load(open(os.path.join(ROOT, "profiles", "r02_traffic.json")))
            traffic_detail = tr.get(args.config) or tr.get("dav2020")
            traffic = float(traffic_detail["dram_bytes_per_site_day"]) * float(M) * nd
        except Exception:
            pass
        line = {
            "metric": METRIC, "value": value, "unit": "site-days/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": scaling,
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": workload_text(args.config, ens, M, world, total, scaling,
                                                 "; per step: integrate (scores accumulated at the observation days by the "
                                                 "integrator, " + ("daily record kept in HBM" if record else "no daily record")
                                                 + ") + score finalisation + score all-gather; every step cold (work quanta of "
                                                 "64 block passes, members regrouped by progress between the launches)"),
                       "members_per_gpu": M, "members_total": total, "n_days": nd, "n_layers": N,
                       "l2": "member blocks (%.0f MB per GPU) exceed the 126 MB L2" % blocks_mb,
                       "members_ok": n_ok, "mean_accepted_steps_per_day": steps_acc,
                       "mean_attempted_steps_per_day": steps_att, "mean_rhs_evaluations_per_day": nf_day,
                       "host_knob_generation_s": t_gen,
                       **({"balanced_shards": balance_note} if balance_note else {})},
            "e2e": {"value": e2e_val, "unit": "site-days/s", "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h)},
            "gpu_launches": n_launched,
            "clocks": clocks,
            "roofline": {"bound": "fp64", "achieved": ach_tf, "peak": fp64_peak, "unit": "TFLOP/s",
                         "frac": ach_tf / fp64_peak if fp64_peak else None,
                         "traffic": traffic,
                         "note": "integrate_lanes; achieved = counted FLOP-equivalents of a step (SURVEY 8d) / CUDA-event "
                                 "time from the first to the last integrator launch of the step (includes the "
                                 "regrouping between the launches); peak = DFMA microbenchmark measured in this run "
                                 "(MEASURED_PEAKS.json has no FP64 entry); traffic = bytes per step = DRAM bytes per "
                                 "site-day of the committed ncu --set full capture (profiles/r02_traffic.json, a 20-day "
                                 "launch WITH the daily record; traffic_per_site_day) x the site-days of a step, against "
                                 "8 (L + 16) algorithmic bytes per site-day; the excess is register spills and the "
                                 "members' live state passing through L2, 1.6 % of the HBM peak",
                         "traffic_per_site_day": traffic_detail, "kernel_ms": kms},
            "roofline_hbm": {"bound": "hbm", "achieved": alg_bytes / (kms * 1e-3) / 1e9, "peak": hbm_peak, "unit": "GB/s",
                             "frac": alg_bytes / (kms * 1e-3) / 1e9 / hbm_peak, "traffic": traffic,
                             "peak_source": "measured" if peaks else "fallback"},
        }
        if parity is not None:
            line["parity"] = {"ok": parity[0], "members_checked": parity[1],
                              "what": "final state, status, step counters of the first members of rank 0 vs the CPU oracle "
                                      "(device elementary functions), bit for bit"}
        if daily_rate is not None:
            line["e2e_daily"] = {"value": daily_rate, "unit": "site-days/s", "members_per_gpu": daily[0],
                                 "d2h_bytes_per_step": int(daily[0] * (nd + 1) * L * 8),
                                 "what": "knob rows H2D + setup on the GPU + run with the daily record + lwfb_get_daily of "
                                         "every member into pinned host memory"}
        if not args.no_cpu_baseline and world == 1:
            cores = os.cpu_count() or 1
            rate, n, dt = cpu_oracle_rate(ens, cores)
            rate1, n1, dt1 = cpu_oracle_rate(ens, 1, target_s=6.0)
            line["cpu_baseline"] = {"value": rate, "unit": "site-days/s", "cores": cores, "kind": "port",
                                    "sample": f"first {n} members x {nd} days in {dt:.1f} s; CPU restatement of the "
                                              "reference algorithm (oracle, host libm), not Julia",
                                    "one_core": {"value": rate1, "sample": f"first {n1} members in {dt1:.1f} s"}}
        print(json.dumps(line))
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
