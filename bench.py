#!/usr/bin/env python
"""bench.py — site-days/s of the batched LWFBrook90 integrator on B200 (contract: see the task brief).

A "step" = one pass of the hot path over one batch: every member of this rank's shard of the
DAV2020-full isotope ensemble integrated over the full record (365 days), scored on the GPU against
synthetic observations, scores all-gathered across ranks (NCCL) when N > 1.

  value  site-days/s with the member blocks already resident in HBM (device time of K steps)
  e2e    the same through the public API with HOST buffers: pinned H2D of the member blocks, run,
         score, D2H of scores + status + counters, every step
  roofline       FP64-pipe roofline of the integrate kernel (the path is FP64-bound)
  roofline_hbm   the secondary HBM figure (algorithmic bytes 8*(L+F) per site-day)
  cpu_baseline   the CPU oracle (restatement of the reference algorithm) on a bounded sample

`--impl reference` times the CPU oracle (the reference itself is Julia and cannot run here) with all
host threads on the same workload definition.
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
# 9 full waves of the lane kernel on one B200: 148 SMs x 2 blocks x 8 warps x 6 members = 14208 members per wave.
# 8 x 127872 = 1,022,976 members at N = 8 (weak scaling).
MEMBERS_PER_GPU = 127872

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


def build_workload(members, first_member=0):
    import lwfbrook90.jl_b200 as lw
    from lwfbrook90.jl_b200 import ensemble as en
    G = os.path.join(ROOT, "tests", "golden", "inputs", "DAV2020-full")
    sim = lw.setup(lw.loadSPAC(G, "DAV2020-full", simulate_isotopes=True))
    ens = en.make_ensemble(sim, members, seed=1234, first_member=first_member)
    return sim, ens


def synthetic_observations(sim):
    """theta / delta18O 'observations' for the score kernel: weekly, layers 1, 3, 5 (0-based 0, 2, 4), taken from
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


def run_reference(args):
    """--impl reference: the CPU restatement of the reference algorithm on the host cores."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    sample = args.ref_members
    sim, ens = build_workload(sample)
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
    line = {"impl": "reference", "metric": METRIC, "value": val, "unit": "site-days/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": f"DAV2020-full isotope ensemble (N=5 layers, {ens.n_days} days, Philox seed 1234), "
                                   f"bounded sample of {ens.M} members per step"},
            "cpu_baseline": {"value": val, "unit": "site-days/s", "cores": cores, "kind": "port",
                             "sample": f"first {ens.M} members x {ens.n_days} days, CPU restatement of the reference "
                                       "algorithm (oracle/lwfb90_oracle.c, host libm); the Julia reference cannot run here"},
            "e2e": {"value": val, "unit": "site-days/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200")
    ap.add_argument("--members", type=int, default=MEMBERS_PER_GPU, help="members per GPU")
    ap.add_argument("--ref-members", type=int, default=2048, help="members per step of the reference arm")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)

    import torch
    from lwfbrook90.jl_b200 import capi, layout as lo
    from lwfbrook90.jl_b200.parallel import DeviceArray, allgather_scores

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    dist = None
    torch.cuda.set_device(local_rank)
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", local_rank))
    M = args.members
    sim, ens = build_workload(M, first_member=rank * M)       # each GPU integrates its own slice of the ensemble
    N, L, nd = ens.N, ens.u0.shape[1], ens.n_days
    obs_day, obs_layer, obs_th, obs_d18 = synthetic_observations(sim)

    # pinned host copies of the member blocks (the e2e leg copies them every step)
    def pin(a):
        t = torch.from_numpy(np.ascontiguousarray(a)).pin_memory()
        return t, t.numpy()
    keep = [pin(ens.scal), pin(ens.lay), pin(ens.u0)]
    scal_h, lay_h, u0_h = (k[1] for k in keep)
    site_h = ens.site_of_member

    h = capi.Handle(N, True, ens.t0_day, nd, device=local_rank, record_daily=True)
    h.set_sites(ens.forcing)
    h.set_members(site_h, scal_h, lay_h, u0_h)
    fp64_peak = capi.fp64_peak(local_rank)

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    def gather(ptr):
        if dist is None:
            return None
        t = torch.as_tensor(DeviceArray(ptr, (M, lo.NSCORES)), device=f"cuda:{local_rank}")
        return allgather_scores(t, M * world, rank, world)

    # our kernels launched per step: n integrator launches, one placement_keys kernel after each of them (regrouping)
    # and one before the first (cost sort of a warm handle), one score kernel; cub's radix-sort kernels are not counted
    launches = []

    def step_resident():
        h.run()
        ptr = h.scores_device(obs_day, obs_layer, obs_th, obs_d18)
        gather(ptr)
        n, ms = h.last_run_stats()
        launches.append(2 * n + 2 if n > 1 else n + 1)
        return ms

    def step_e2e():
        h.set_members(site_h, scal_h, lay_h, u0_h)            # H2D from pinned host memory
        h.run()
        sc = h.scores(obs_day, obs_layer, obs_th, obs_d18)    # D2H of the result
        st, cnt = h.get_status()
        if dist is not None:
            t = torch.from_numpy(sc).to(f"cuda:{local_rank}")
            allgather_scores(t, M * world, rank, world)
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

    barrier()
    t0 = time.perf_counter()
    for _ in range(max(1, args.steps)):
        sc, st2, cnt2 = step_e2e()
    barrier()
    wall_e2e = (time.perf_counter() - t0) / max(1, args.steps)

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
    flops = flop_eq(N, True, nd, cnt)
    n_ok = allsum(float((st == 0).sum()))

    if rank == 0:
        site_days = float(M) * world * nd
        ms_per_step = 1e3 * wall / args.steps
        value = site_days / (ms_per_step * 1e-3)
        e2e_val = site_days / wall_e2e
        h2d = M * (lo.NPS + lo.NPL * N + L) * 8 + M * 4
        d2h = M * lo.NSCORES * 8 + M * 4 + M * lo.NCOUNTERS * 8
        ach_tf = flops / (kms * 1e-3) / 1e12
        alg_bytes = float(M) * nd * 8 * (L + lo.NF)
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
        line = {
            "metric": METRIC, "value": value, "unit": "site-days/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": f"DAV2020-full isotope ensemble: {M} members per GPU x {world} GPU(s) "
                                   f"(= {8 * M:,} at 8 GPUs), N=5 layers, {nd} days, state length {L}, Philox seed 1234; "
                                   "per step: integrate + score kernel + score all-gather; adaptive placement "
                                   "(work quanta of 128 block passes, members regrouped by progress between the "
                                   "launches, lwfb_set_placement)",
                       "members_per_gpu": M, "n_days": nd, "n_layers": N,
                       "l2": "member blocks (%.0f MB) and daily record (%.1f GB) per launch exceed the 126 MB L2"
                             % (h2d / 1e6, M * (nd + 1) * L * 8 / 1e9),
                       "members_ok": n_ok, "mean_accepted_steps_per_day": float(cnt[:, 0].mean() / nd),
                       "rejected_fraction": float(cnt[:, 1].sum() / max(1, cnt[:, 0].sum() + cnt[:, 1].sum()))},
            "e2e": {"value": e2e_val, "unit": "site-days/s", "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h)},
            "gpu_launches": n_launched,
            "clocks": clocks,
            "roofline": {"bound": "fp64", "achieved": ach_tf, "peak": fp64_peak, "unit": "TFLOP/s",
                         "frac": ach_tf / fp64_peak if fp64_peak else None, "traffic": None,
                         "note": "integrate_lanes; achieved = counted FLOP-equivalents of a step (SURVEY 8d) / CUDA-event "
                                 "time from the first to the last integrator launch of the step (includes the host "
                                 "regrouping between the launches); "
                                 "peak = DFMA microbenchmark measured in this run (MEASURED_PEAKS.json has no FP64 entry)",
                         "kernel_ms": kms},
            "roofline_hbm": {"bound": "hbm", "achieved": alg_bytes / (kms * 1e-3) / 1e9, "peak": hbm_peak, "unit": "GB/s",
                             "frac": alg_bytes / (kms * 1e-3) / 1e9 / hbm_peak, "traffic": None,
                             "peak_source": "measured" if peaks else "fallback"},
        }
        if not args.no_cpu_baseline and world == 1:
            cores = os.cpu_count() or 1
            rate, n, dt = cpu_oracle_rate(ens, cores)
            line["cpu_baseline"] = {"value": rate, "unit": "site-days/s", "cores": cores, "kind": "port",
                                    "sample": f"first {n} members x {nd} days in {dt:.1f} s; CPU restatement of the "
                                              "reference algorithm (oracle, host libm), not Julia"}
        print(json.dumps(line))
    h.close()
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
