"""Developer tool (CPU), round 2: replays exact per-path traversal event traces from the oracle (tools/sched_sim.py:get_traces)
through models of three scheduling designs, with per-body costs taken from the ncu source page of the shipped kernel (59 warp
instructions per interior visit, 80 per rect test, 8 per vote, 330 per shaded segment).  The model of the shipped kernel
(segment-synchronous warps, interior body while nI >= 6 nL, four visits per vote) reproduces the measured 17.4 lanes per
interior-body instruction, which is what makes the other rows worth reading.

    python tools/sched_sim_r2.py regroup | pool | threeway | soft | pair      (maze 32, every 797th chunk group of the 1080p frame)

regroup   upper bound of re-forming warps inside a block at every segment boundary: rays sorted by their TRUE visit count
threeway  one warp, lanes desynchronised across segments: a vote picks interior / leaf / shade+set-up
soft      soft segment boundaries: shade and restart the finished lanes as soon as at most k stragglers are still traversing
pair      cooperative interior body: a finished lane takes one of the two slab tests of a still-traversing lane (42 instead of 59
          instructions per visit) whenever every traversing lane can get such a helper
pool      persistent warp over a pool of M paths in shared memory with per-body ready queues (built: pool_kernel.cu)
"""
import os, pickle, random, sys
from collections import deque
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
C_I, C_L, C_V, C_S = 59.0, 80.0, 8.0, 330.0


def traces():
    cache = "/tmp/mm_traces32.pkl"
    if os.path.exists(cache):
        return pickle.load(open(cache, "rb"))
    import sched_sim
    paths, _ = sched_sim.get_traces(maze=32, every=797)
    pickle.dump(paths, open(cache, "wb"))
    return paths


def seg_events(seg):
    e = []
    for run, leaf in seg:
        e.extend([0] * run)
        if leaf:
            e.append(leaf)
    return e


def path_events(p):
    e = []
    for seg in p:
        e.extend(seg_events(seg))
        e.append(-1)
    return e


def seg_cost(lanes, reps=4, w=6):
    """The shipped kernel's votes inside one segment for <= 32 rays.  Returns (warp instr, interior lane-visits, interior execs)."""
    ev = [seg_events(l) for l in lanes]
    pos = [0] * len(ev)
    tot, li, ni = 0.0, 0, 0
    while True:
        I = [i for i in range(len(ev)) if pos[i] < len(ev[i]) and ev[i][pos[i]] == 0]
        L = [i for i in range(len(ev)) if pos[i] < len(ev[i]) and ev[i][pos[i]] != 0]
        if not I and not L:
            break
        tot += C_V
        if I and len(I) >= w * len(L):
            for _ in range(reps):
                act = [i for i in I if pos[i] < len(ev[i]) and ev[i][pos[i]] == 0]
                if not act:
                    break
                tot += C_I; li += len(act); ni += 1
                for i in act:
                    pos[i] += 1
        else:
            tot += C_L * max(ev[i][pos[i]] for i in L)
            for i in L:
                pos[i] += 1
    return tot, li, ni


def run_regroup(paths):
    rays = sum(len(p) for p in paths)
    for B in (32, 64, 256, 1024):
        for mode in ("none", "oracle", "random"):
            tot, li, ni = 0.0, 0, 0
            for b0 in range(0, len(paths) - B + 1, B):
                blk = paths[b0:b0 + B]
                for s in range(max(len(p) for p in blk)):
                    if mode == "none":
                        groups = [[p[s] for p in blk[w0:w0 + 32] if len(p) > s] for w0 in range(0, B, 32)]
                    else:
                        alive = [p[s] for p in blk if len(p) > s]
                        if mode == "oracle":
                            alive.sort(key=lambda sg: sum(r for r, _ in sg) + 2 * sum(1 for _, l in sg if l))
                        else:
                            random.shuffle(alive)
                        groups = [alive[w0:w0 + 32] for w0 in range(0, len(alive), 32)]
                    for g in groups:
                        if g:
                            t, a, b = seg_cost(g); tot += t; li += a; ni += b
            print(f"block {B:5d} {mode:7s}: {tot / rays:7.1f} traversal warp-instr/ray, {li / max(ni, 1):.1f} lanes per interior exec")


def run_threeway(paths):
    warps = [paths[i:i + 32] for i in range(0, len(paths), 32)][:600]
    rays = sum(len(p) for w in warps for p in w)
    base = 0.0
    for w in warps:
        for s in range(max(len(p) for p in w)):
            base += seg_cost([p[s] for p in w if len(p) > s])[0] + C_S + 40
    print(f"shipped (segment-synchronous, voted): {base / rays:.1f} warp-instr/ray")
    for wL, wS in ((6, 6), (6, 3), (6, 2), (4, 2), (6, 1)):
        tot, li, ni = 0.0, 0, 0
        for w in warps:
            ev = [path_events(p) for p in w]; pos = [0] * len(ev)
            while True:
                I = [i for i in range(len(ev)) if pos[i] < len(ev[i]) and ev[i][pos[i]] == 0]
                L = [i for i in range(len(ev)) if pos[i] < len(ev[i]) and ev[i][pos[i]] > 0]
                S = [i for i in range(len(ev)) if pos[i] < len(ev[i]) and ev[i][pos[i]] == -1]
                if not (I or L or S):
                    break
                tot += 12.0
                if I and len(I) >= wL * len(L) and len(I) >= wS * len(S):
                    for _ in range(4):
                        act = [i for i in I if pos[i] < len(ev[i]) and ev[i][pos[i]] == 0]
                        if not act:
                            break
                        tot += C_I; li += len(act); ni += 1
                        for i in act:
                            pos[i] += 1
                elif L and len(L) >= len(S):
                    tot += C_L * max(ev[i][pos[i]] for i in L)
                    for i in L:
                        pos[i] += 1
                else:
                    tot += C_S + 30
                    for i in S:
                        pos[i] += 1
        print(f"three-way vote, leaf weight {wL}, shade weight {wS}: {tot / rays:.1f} warp-instr/ray, {li / ni:.1f} lanes per interior exec")


def run_pool(paths, gI=55.0, gL=30.0, gS=40.0, cS=300.0, cV=12.0, gen=60.0):
    W = 4096
    for M, N, th in ((64, 4, 24), (96, 4, 24), (128, 4, 24), (128, 2, 24), (128, 6, 24)):
        tot, rays = 0.0, 0
        lanes = [0, 0]
        for w0 in range(0, len(paths) - W + 1, W):
            evs = [path_events(p) for p in paths[w0:w0 + W]]
            rays += sum(len(p) for p in paths[w0:w0 + W])
            nxt, slots, pos = 0, [None] * M, [0] * M
            qI, qL, qS, empty = deque(), deque(), deque(), list(range(M))

            def classify(s):
                e = evs[slots[s]][pos[s]]
                (qI if e == 0 else qS if e == -1 else qL).append(s)
            while True:
                if len(empty) >= 32 and nxt < len(evs):
                    tot += gen
                    for _ in range(min(32, len(evs) - nxt)):
                        s = empty.pop(); slots[s] = nxt; pos[s] = 0; nxt += 1; classify(s)
                    continue
                if not (qI or qL or qS):
                    break
                tot += cV
                nI, nL, nS = len(qI), len(qL), len(qS)
                act = "I" if nI >= 32 else "L" if nL >= th else "S" if nS >= th else max((("I", nI), ("L", nL), ("S", nS)), key=lambda t: t[1])[0]
                if act == "I":
                    sel = [qI.popleft() for _ in range(min(32, nI))]
                    tot += gI
                    live = list(sel)
                    for _ in range(N):
                        live = [s for s in live if evs[slots[s]][pos[s]] == 0]
                        if not live:
                            break
                        tot += C_I; lanes[0] += len(live); lanes[1] += 1
                        for s in live:
                            pos[s] += 1
                    for s in sel:
                        classify(s)
                elif act == "L":
                    sel = [qL.popleft() for _ in range(min(32, nL))]
                    tot += gL + C_L * max(evs[slots[s]][pos[s]] for s in sel)
                    for s in sel:
                        pos[s] += 1; classify(s)
                else:
                    sel = [qS.popleft() for _ in range(min(32, nS))]
                    tot += gS + cS
                    for s in sel:
                        pos[s] += 1
                        if pos[s] >= len(evs[slots[s]]):
                            slots[s] = None; empty.append(s)
                        else:
                            classify(s)
        print(f"pool M={M:3d}, {N} visits per batch, threshold {th}: {tot / rays:6.1f} warp-instr/ray, {lanes[0] / lanes[1]:.1f} lanes per interior exec")


def run_soft(paths):
    warps = [paths[i:i + 32] for i in range(0, len(paths), 32)][:800]
    rays = sum(len(p) for w in warps for p in w)
    for k in (0, 1, 2, 4, 8):
        tot, li, ni, ns, ls = 0.0, 0, 0, 0, 0
        for w in warps:
            ev = [path_events(p) for p in w]; pos = [0] * len(ev)
            while True:
                I = [i for i in range(len(ev)) if pos[i] < len(ev[i]) and ev[i][pos[i]] == 0]
                L = [i for i in range(len(ev)) if pos[i] < len(ev[i]) and ev[i][pos[i]] > 0]
                S = [i for i in range(len(ev)) if pos[i] < len(ev[i]) and ev[i][pos[i]] == -1]
                if not (I or L or S):
                    break
                if S and len(I) + len(L) <= k:                       # k = 0 is the shipped kernel
                    tot += C_S + 40; ns += 1; ls += len(S)
                    for i in S:
                        pos[i] += 1
                    continue
                tot += C_V
                if I and len(I) >= 6 * len(L):
                    for _ in range(4):
                        act = [i for i in I if pos[i] < len(ev[i]) and ev[i][pos[i]] == 0]
                        if not act:
                            break
                        tot += C_I; li += len(act); ni += 1
                        for i in act:
                            pos[i] += 1
                else:
                    tot += C_L * max(ev[i][pos[i]] for i in L)
                    for i in L:
                        pos[i] += 1
        print(f"soft boundary, {k} stragglers allowed: {tot / rays:6.1f} warp-instr/ray, {li / ni:.1f} lanes per interior exec, "
              f"{ns / len(warps):.1f} shade execs per warp at {ls / ns:.1f} lanes")


def run_pair(paths, c_pair=42.0, c_setup=30.0):
    warps = [paths[i:i + 32] for i in range(0, len(paths), 32)][:800]
    rays = sum(len(p) for w in warps for p in w)
    for pair in (False, True):
        tot, execs, paired = 0.0, 0, 0
        for w in warps:
            for s in range(max(len(p) for p in w)):
                ev = [seg_events(p[s]) for p in w if len(p) > s]
                n = len(ev); pos = [0] * n; helped = set(); free = 32 - n
                while True:
                    I = [i for i in range(n) if pos[i] < len(ev[i]) and ev[i][pos[i]] == 0]
                    L = [i for i in range(n) if pos[i] < len(ev[i]) and ev[i][pos[i]] != 0]
                    if not I and not L:
                        break
                    tot += C_V
                    if I and len(I) >= 6 * len(L):
                        active = [i for i in range(n) if pos[i] < len(ev[i])]
                        avail = free + (n - len(active)) - sum(1 for o in active if o in helped)
                        need = [o for o in I if o not in helped]
                        use = pair and len(need) <= avail
                        if use and need:
                            tot += c_setup; helped.update(need)
                        for _ in range(4):
                            act = [i for i in I if pos[i] < len(ev[i]) and ev[i][pos[i]] == 0]
                            if not act:
                                break
                            tot += c_pair if use else C_I; execs += 1; paired += 1 if use else 0
                            for i in act:
                                pos[i] += 1
                    else:
                        tot += C_L * max(ev[i][pos[i]] for i in L)
                        for i in L:
                            pos[i] += 1
        print(f"{'pair mode' if pair else 'shipped  '}: {tot / rays:6.1f} traversal warp-instr/ray, {paired / max(execs, 1):.3f} of the interior execs in pair mode")


if __name__ == "__main__":
    mode = sys.argv[1] if len(sys.argv) > 1 else "threeway"
    {"regroup": run_regroup, "threeway": run_threeway, "pool": run_pool, "soft": run_soft, "pair": run_pair}[mode](traces())
