"""Developer tool (CPU): replays exact per-path traversal event traces from the oracle through models of the warp
scheduling policies the kernel could use, and reports estimated warp-instruction counts.  Lets scheduling ideas be
compared without GPU time.  Costs per body (warp instructions) are taken from the ncu source page of the kernel."""
import ctypes as C
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import mirror_maze_b200 as mm
from oracle import oracle

octants = []   # per path, per segment: direction octant (filled by get_traces)
C_I, C_L, C_S, C_O = 103.0, 95.0, 330.0, 10.0     # interior body, leaf body per rect, shade+next-segment, per-iteration vote overhead


def get_traces(maze=32, W=1920, H=1080, spp=16, bounces=8, every=97):
    L = oracle.lib()
    L.mmo_trace.restype = C.c_uint64
    noise = mm.load_noise()
    sc = mm.MazeScene(maze, 0)
    u = mm.default_uniform(maze, W, H, 4)
    ch = mm.gen_chunks(W, H, 4)
    p = mm.full_frame_params(u, spp=spp, bounce_limit=bounces)
    n_groups = p.grid_x * p.grid_y
    p.group_first, p.group_step, p.group_count = 13, every, (n_groups - 13 + every - 1) // every
    args = [sc.planes.ctypes.data, len(sc.planes), sc.nodes.ctypes.data, len(sc.nodes), sc.indices.ctypes.data, sc.materials.ctypes.data,
            sc.emissions.ctypes.data, noise.ctypes.data, 512, 512, C.addressof(u), C.addressof(p), ch.ctypes.data, len(ch)]
    argt = [C.c_void_p, C.c_uint32, C.c_void_p, C.c_uint32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32, C.c_uint32,
            C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32, C.c_void_p, C.c_uint64]
    L.mmo_trace.argtypes = argt
    n = L.mmo_trace(*args, None, 0)
    buf = np.zeros(n, dtype=np.uint16)
    L.mmo_trace(*args, buf.ctypes.data, n)
    # parse into paths: list of segments; segment = list of (run, leafcount) with final (run, None)
    paths, segs, cur = [], [], []
    octants.clear()
    octs = []
    it = iter(buf.tolist())
    for tok in it:
        if tok == 0xFFFE:
            paths.append(segs); segs = []
            octants.append(octs); octs = []
            continue
        if (tok & 0xFFF8) == 0xFFF0:
            octs.append((tok & 7, next(it)))      # (octant, |dir.y|/|dir| in 0..255)
            continue
        nxt = next(it)
        if nxt == 0xFFFF:
            cur.append((tok, 0)); segs.append(cur); cur = []
        else:
            cur.append((tok, nxt))
    return paths, p.group_count * 16 * spp


def lane_events(path):
    """Flatten a path into a list of events: 'I' per interior visit, ('L', count) per leaf, 'S' per shade (one per segment)."""
    ev = []
    for seg in path:
        for run, leaf in seg:
            ev.extend(["I"] * run)
            if leaf:
                ev.append(("L", leaf))
        ev.append("S")
    return ev


def sim_v1(warp):
    """Warp-synchronous segments, while-while inside a segment (kernel v1)."""
    total = 0.0
    nseg = max(len(p) for p in warp)
    for s in range(nseg):
        lanes = [list(p[s]) for p in warp if len(p) > s]
        pos = [0] * len(lanes)
        while True:
            act = [i for i in range(len(lanes)) if pos[i] < len(lanes[i])]
            if not act:
                break
            total += C_I * max(lanes[i][pos[i]][0] for i in act)
            leafc = max(lanes[i][pos[i]][1] for i in act)
            total += C_L * leafc
            for i in act:
                pos[i] += 1
        total += C_S + 60          # shade + per-segment set-up, all lanes together
    return total


def sim_sm(warp, th=16, wI=1, wL=1, wS=1, c_o=C_O, shade_scale=1.0):
    """Lane state machines with warp votes (kernel v2)."""
    evs = [lane_events(p) for p in warp]
    pos = [0] * len(evs)
    total = 0.0
    lanesI = lanesL = lanesS = 0.0
    nI_ = nL_ = nS_ = 0
    while True:
        I = [i for i in range(len(evs)) if pos[i] < len(evs[i]) and evs[i][pos[i]] == "I"]
        Lq = [i for i in range(len(evs)) if pos[i] < len(evs[i]) and isinstance(evs[i][pos[i]], tuple)]
        S = [i for i in range(len(evs)) if pos[i] < len(evs[i]) and evs[i][pos[i]] == "S"]
        if not (I or Lq or S):
            break
        total += c_o
        if len(I) >= th:
            act = 0
        else:
            sI, sL, sS = len(I) * wI, len(Lq) * wL, len(S) * wS
            act = 0 if (sI >= sL and sI >= sS and I) else (1 if (sL >= sS and Lq) else 2)
        if act == 0:
            total += C_I; lanesI += len(I); nI_ += 1
            for i in I: pos[i] += 1
        elif act == 1:
            total += C_L * max(evs[i][pos[i]][1] for i in Lq); lanesL += len(Lq); nL_ += 1
            for i in Lq: pos[i] += 1
        else:
            total += C_S * shade_scale; lanesS += len(S); nS_ += 1
            for i in S: pos[i] += 1
    return total, (lanesI / max(nI_, 1), lanesL / max(nL_, 1), lanesS / max(nS_, 1), nI_, nL_, nS_)


if __name__ == "__main__" and not (len(sys.argv) > 1 and sys.argv[1] in ("multi", "pool", "block")):
    maze = int(sys.argv[1]) if len(sys.argv) > 1 else 32
    paths, n_paths = get_traces(maze=maze, every=int(sys.argv[2]) if len(sys.argv) > 2 else 397)
    assert len(paths) == n_paths, (len(paths), n_paths)
    warps = [paths[i:i + 32] for i in range(0, len(paths), 32)]
    rays = sum(len(p) for p in paths)
    inner = sum(r for p in paths for s in p for r, _ in s)
    print(f"maze {maze}: {len(paths)} paths, {len(warps)} warps, {rays} rays, {inner / rays:.2f} interior visits/ray")
    ideal = sum(C_I * r + C_L * l for p in paths for s in p for r, l in s) / 32 + rays * (C_S + 60) / 32
    print(f"ideal (all lanes always busy): {ideal / rays:8.1f} warp-instr/ray")
    v1 = sum(sim_v1(w) for w in warps)
    print(f"v1 warp-synchronous           : {v1 / rays:8.1f} warp-instr/ray  (efficiency {ideal / v1:.2f})")
    for cfg in [(16, 1, 1, 1), (16, 1, 2, 2), (16, 2, 1, 1), (33, 1, 1, 1), (16, 1, 3, 3), (16, 1, 4, 4), (24, 1, 2, 2), (16, 1, 2, 3), (8, 1, 2, 2)]:
        tot, stats = 0.0, np.zeros(6)
        for w in warps:
            t, st = sim_sm(w, *cfg)
            tot += t; stats += np.array(st)
        stats[:3] /= len(warps)
        print(f"v2 th={cfg[0]:2d} w={cfg[1:]}: {tot / rays:8.1f} warp-instr/ray (eff {ideal / tot:.2f})  lanes/exec I={stats[0]:.1f} L={stats[1]:.1f} S={stats[2]:.1f} "
              f"execs I={stats[3]:.0f} L={stats[4]:.0f} S={stats[5]:.0f}")


def sim_multi(warpK, K, th=28, wI=1, wL=1, wS=1, c_swap=70.0):
    """K paths per lane: a lane offers whichever of its paths is ready for the body being executed (kernel v3 idea).
    warpK = list of 32*K paths; lane i owns paths i, i+32, ...  A swap (path state to/from shared memory) is charged
    per lane-switch at 1/32 of c_swap per lane (it runs warp-wide when any lane swaps: charged fully once per iteration with swaps)."""
    evs = [lane_events(p) for p in warpK]
    pos = [0] * len(evs)
    cur = [0] * 32                      # which of its K paths each lane currently holds in registers
    total = 0.0
    nI_ = 0; lanesI = 0.0; swaps = 0
    def nxt(j):
        return evs[j][pos[j]] if pos[j] < len(evs[j]) else None
    while True:
        readyI, readyL, readyS = [], [], []
        for lane in range(32):
            mine = [lane + 32 * k for k in range(K) if lane + 32 * k < len(evs)]
            order = sorted(mine, key=lambda j: 0 if (j - lane) // 32 == cur[lane] else 1)
            i = next((j for j in order if nxt(j) == "I"), None)
            l = next((j for j in order if isinstance(nxt(j), tuple)), None)
            s = next((j for j in order if nxt(j) == "S"), None)
            if i is not None: readyI.append((lane, i))
            if l is not None: readyL.append((lane, l))
            if s is not None: readyS.append((lane, s))
        if not (readyI or readyL or readyS):
            break
        total += C_O
        if len(readyI) >= th:
            act = 0
        else:
            sI, sL, sS = len(readyI) * wI, len(readyL) * wL, len(readyS) * wS
            act = 0 if (sI >= sL and sI >= sS and readyI) else (1 if (sL >= sS and readyL) else 2)
        sel = (readyI, readyL, readyS)[act]
        sw = 0
        for lane, j in sel:
            if (j - lane) // 32 != cur[lane]:
                cur[lane] = (j - lane) // 32; sw += 1
        if sw:
            total += c_swap; swaps += sw
        if act == 0:
            total += C_I; lanesI += len(sel); nI_ += 1
        elif act == 1:
            total += C_L * max(nxt(j)[1] for _, j in sel)
        else:
            total += C_S
        for _, j in sel:
            pos[j] += 1
    return total, lanesI / max(nI_, 1), swaps


def run_multi():
    maze = int(sys.argv[2]) if len(sys.argv) > 2 else 32
    paths, n_paths = get_traces(maze=maze, every=int(sys.argv[3]) if len(sys.argv) > 3 else 1597)
    rays = sum(len(p) for p in paths)
    ideal = sum(C_I * r + C_L * l for p in paths for s in p for r, l in s) / 32 + rays * (C_S + 60) / 32
    print(f"{len(paths)} paths {rays} rays; ideal {ideal / rays:.1f} warp-instr/ray")
    for K in (1, 2, 3, 4):
        for th, w in ((16, (1, 1, 1)), (24, (1, 1, 1)), (28, (1, 1, 1)), (28, (1, 2, 2))):
            tot, lanes, sw, nw = 0.0, 0.0, 0, 0
            for i in range(0, len(paths) - 32 * K + 1, 32 * K):
                t, l, s = sim_multi(paths[i:i + 32 * K], K, th, *w)
                tot += t; lanes += l; sw += s; nw += 1
            r = sum(len(p) for p in paths[: nw * 32 * K])
            print(f"K={K} th={th} w={w}: {tot / r:7.1f} warp-instr/ray (eff {ideal / rays / (tot / r):.2f}) lanes/I-exec {lanes / nw:.1f} swaps/ray {sw / r:.2f}")


if __name__ == "__main__" and len(sys.argv) > 1 and sys.argv[1] == "multi":
    run_multi()


def sim_pool(paths, M, th=28, c_park=25.0, c_vote=12.0):
    """A warp owns a pool of M paths whose state lives in shared memory; each body runs on up to 32 paths that are
    ready for it (kernel v3 idea).  Lanes keep interior-state paths in registers; a lane whose path leaves the interior
    state parks it and refills (c_park thread-instructions per transition, charged as warp instructions / 32 * lanes)."""
    evs = [lane_events(p) for p in paths]
    pos = [0] * len(evs)
    total, execs, lanes = 0.0, [0, 0, 0], [0.0, 0.0, 0.0]
    alive = list(range(len(evs)))
    while alive:
        I = [j for j in alive if evs[j][pos[j]] == "I"]
        Lq = [j for j in alive if isinstance(evs[j][pos[j]], tuple)]
        S = [j for j in alive if evs[j][pos[j]] == "S"]
        total += c_vote
        if len(I) >= th or (not Lq and not S):
            act, sel = 0, I[:32]
        elif len(Lq) >= len(S):
            act, sel = 1, Lq[:32]
        else:
            act, sel = 2, S[:32]
        if act == 0:
            total += C_I
        elif act == 1:
            total += C_L * max(evs[j][pos[j]][1] for j in sel) + c_park
        else:
            total += C_S + 60 + c_park
        execs[act] += 1; lanes[act] += len(sel)
        for j in sel:
            pos[j] += 1
        alive = [j for j in alive if pos[j] < len(evs[j])]
    return total, execs, lanes


def run_pool():
    maze = int(sys.argv[2]) if len(sys.argv) > 2 else 32
    paths, n_paths = get_traces(maze=maze, every=int(sys.argv[3]) if len(sys.argv) > 3 else 1597)
    rays = sum(len(p) for p in paths)
    ideal = sum(C_I * r + C_L * l for p in paths for s in p for r, l in s) / 32 + rays * (C_S + 60) / 32
    print(f"{len(paths)} paths {rays} rays; ideal {ideal / rays:.1f} warp-instr/ray")
    for M in (32, 64, 96, 128, 256):
        for th in (24, 32):
            tot, ex, ln = 0.0, np.zeros(3), np.zeros(3)
            n = 0
            for i in range(0, len(paths) - M + 1, M):
                t, e, l = sim_pool(paths[i:i + M], M, th)
                tot += t; ex += e; ln += l; n += M
            r = sum(len(p) for p in paths[:n])
            print(f"M={M:3d} th={th}: {tot / r:7.1f} warp-instr/ray (eff {ideal / rays / (tot / r):.2f}) lanes/exec I={ln[0]/ex[0]:.1f} L={ln[1]/ex[1]:.1f} S={ln[2]/ex[2]:.1f}")


if __name__ == "__main__" and len(sys.argv) > 1 and sys.argv[1] == "pool":
    run_pool()


def sim_block(paths, NP=512, n_i=6, n_ls=2, th_refill=32, regen=True, c_refill=30.0, c_wb=12.0):
    """Block-shared pool: n_i interior warps keep interior-state paths resident and refill emptied lanes from the
    pool (a lane may only claim paths of its residue class mod 32); n_ls warps batch leaf / shade work.  Event-driven:
    every warp's bodies take time proportional to their warp-instruction count (equal issue share).  Returns total
    warp-instructions issued (busy, incl. idle polling at 20/poll) and the makespan in instruction-times."""
    import heapq
    evs_all = [lane_events(p) for p in paths]
    nxt_new = [NP]                                # next unstarted path (regeneration from the global counter)
    pool = list(range(min(NP, len(evs_all))))     # pool slot -> path id
    pos = {j: 0 for j in pool}
    state = {}                                    # slot -> 'I','L','S','O' (owned), 'D'
    def st_of(slot):
        j = pool[slot]
        if j is None: return 'D'
        e = evs_all[j][pos[j]]
        return 'I' if e == 'I' else ('L' if isinstance(e, tuple) else 'S')
    for s in range(len(pool)): state[s] = st_of(s)
    resident = [[None] * 32 for _ in range(n_i)]  # per I-warp lane -> slot
    issued = 0.0
    heap = [(0.0, w) for w in range(n_i + n_ls)]
    heapq.heapify(heap)
    live = len(pool)
    tend = 0.0
    lanesI = execI = 0
    while heap and live > 0:
        tnow, w = heapq.heappop(heap)
        cost = 0.0
        if w < n_i:
            res = resident[w]
            empties = [l for l in range(32) if res[l] is None]
            if len(empties) > 32 - th_refill or len(empties) == 32:
                got = 0
                for l in empties:
                    cand = next((s for s in range(l, len(pool), 32) if state[s] == 'I'), None)
                    if cand is not None:
                        state[cand] = 'O'; res[l] = cand; got += 1
                if got: cost += c_refill
            act = [l for l in range(32) if res[l] is not None]
            if act:
                cost += C_I + C_O
                lanesI += len(act); execI += 1
                wb = False
                for l in act:
                    s = res[l]; j = pool[s]; pos[j] += 1
                    ns = st_of(s)
                    if ns != 'I':
                        state[s] = ns; res[l] = None; wb = True
                if wb: cost += c_wb
            else:
                cost += 20.0                      # poll
        else:
            nL = [s for s in state if state[s] == 'L']; nS = [s for s in state if state[s] == 'S']
            def claim(kind):
                sel = []
                for l in range(32):
                    cand = next((s for s in range(l, len(pool), 32) if state[s] == kind), None)
                    if cand is not None:
                        state[cand] = 'O'; sel.append(cand)
                return sel
            if nL or nS:
                kind = 'L' if len(nL) >= len(nS) else 'S'
                sel = claim(kind)
                if kind == 'L':
                    cost += C_L * max(evs_all[pool[s]][pos[pool[s]]][1] for s in sel) + c_refill
                else:
                    cost += C_S + 60 + c_refill
                for s in sel:
                    j = pool[s]; pos[j] += 1
                    if pos[j] >= len(evs_all[j]):
                        if regen and nxt_new[0] < len(evs_all):
                            j2 = nxt_new[0]; nxt_new[0] += 1; pool[s] = j2; pos[j2] = 0; state[s] = st_of(s)
                        else:
                            pool[s] = None; state[s] = 'D'; live -= 1
                    else:
                        state[s] = st_of(s)
            else:
                cost += 20.0
        issued += cost
        tend = tnow + cost
        heapq.heappush(heap, (tnow + cost, w))
    return issued, tend, lanesI / max(execI, 1)


def run_block():
    maze = int(sys.argv[2]) if len(sys.argv) > 2 else 32
    paths, n_paths = get_traces(maze=maze, every=int(sys.argv[3]) if len(sys.argv) > 3 else 1597)
    rays = sum(len(p) for p in paths)
    ideal = sum(C_I * r + C_L * l for p in paths for s in p for r, l in s) / 32 + rays * (C_S + 60) / 32
    print(f"{len(paths)} paths {rays} rays; ideal {ideal / rays:.1f} warp-instr/ray")
    for NP, n_i, n_ls, thr in ((512, 6, 2, 32), (512, 6, 2, 24), (512, 7, 1, 32), (384, 6, 2, 32), (256, 6, 2, 32), (512, 5, 3, 32), (1024, 6, 2, 32)):
        issued, tend, li = sim_block(paths, NP, n_i, n_ls, thr)
        nw = n_i + n_ls
        print(f"NP={NP} I-warps={n_i} LS-warps={n_ls} refill_th={thr}: issued {issued / rays:6.1f} warp-instr/ray, makespan*warps {tend * nw / rays:6.1f} "
              f"(eff vs ideal {ideal / rays / (tend * nw / rays):.2f}), lanes/I-exec {li:.1f}")


if __name__ == "__main__" and len(sys.argv) > 1 and sys.argv[1] == "block":
    run_block()
