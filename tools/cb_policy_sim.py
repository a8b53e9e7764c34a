"""Offline model of the lane kernel's lock-step block: when should the block-uniform daily-callback phase run?

Input: measured attempted steps per member and day (steps_per_day*.npy, written by tools/cost_dist.py on the GPU).
A block of B members advances in passes; in a pass every member that is not waiting at its day boundary makes one step.
A callback phase costs C_CB passes for the whole block and releases the members that wait.  Policies:
  wait W     (round 1/2 kernel)  a phase runs when nobody can step or some member has waited W passes
  waste T    a phase runs when nobody can step or the member-passes idled since the last phase reach T x (members stepping)
usage: cb_policy_sim.py steps.npy B [days] [blocks]"""
import sys
import numpy as np

a = np.load(sys.argv[1]).astype(np.int64)
B = int(sys.argv[2])
D = int(sys.argv[3]) if len(sys.argv) > 3 else 60
NB = int(sys.argv[4]) if len(sys.argv) > 4 else 100
C_CB = 1.5
rng = np.random.default_rng(0)


def run_block(steps, policy, par):
    """steps [B, D]; returns (time in passes incl. callback phases, callback phases)"""
    Bn, Dn = steps.shape
    day = np.zeros(Bn, dtype=int)
    left = steps[:, 0].copy()             # steps left in the current day
    waiting = np.zeros(Bn, dtype=bool)
    waited = np.zeros(Bn, dtype=int)
    done = np.zeros(Bn, dtype=bool)
    t = 0.0; phases = 0; idle_acc = 0.0; occ = 0.0     # occ: member-passes of members that have not finished (regrouping takes the finished ones out of a block)
    while not done.all():
        can = ~waiting & ~done
        n_can = can.sum(); n_wait = waiting.sum()
        fire = n_can == 0
        if not fire and n_wait > 0:
            if policy == "wait": fire = waited[waiting].max() >= par
            else: fire = idle_acc >= par * n_can
        if fire:
            t += C_CB; phases += 1; idle_acc = 0.0; occ += C_CB * (~done).sum()
            idx = np.where(waiting)[0]
            day[idx] += 1
            fin = day[idx] >= Dn
            done[idx[fin]] = True
            go = idx[~fin]
            left[go] = steps[go, day[go]]
            waiting[idx] = False; waited[idx] = 0
            continue
        t += 1.0; occ += (~done).sum()
        left[can] -= 1
        waited[waiting] += 1
        idle_acc += n_wait
        arr = can & (left <= 0)
        waiting |= arr
    return occ, phases


for policy, pars in (("wait", [8, 16, 32, 64]), ("waste", [0.25, 0.5, 1.0, 1.5, 2.0, 3.0])):
    for par in pars:
        T = 0.0; P = 0; S = 0
        r = np.random.default_rng(1)
        for b in range(NB):
            mem = r.choice(a.shape[0], B, replace=False)
            d0 = r.integers(0, a.shape[1] - D + 1)
            st = a[mem, d0:d0 + D]
            t, p = run_block(st, policy, par)
            T += t; P += p; S += st.sum()
        print(f"{policy:6s} {par:5}: executed / needed = {T / S:.3f}   callback phases per block-day = {P / NB / D:.2f}")
