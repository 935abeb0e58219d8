"""host_ref.py — independent Python restatement of the reference's host surface (TEST INFRASTRUCTURE).

Follows reference src/main.rs:328-352 (TreeBuilder), :357-396 (Kruskal), :397-438 (walls), :443-586 (scene),
:91-263 (BVH) and the published algorithms of rand 0.8.5 / rand_chacha 0.3.1 / rand_core 0.6.4 (Cargo.lock:342-371;
NOT vendored under /root/reference, so the RNG stream is PARITY UNPINNED against a real `cargo run`).
Used by tests/test_host_surface.py to cross-check the C++ restatement in mirror_maze_b200/csrc array for array at small n.
All scene arithmetic is numpy float32 (one rounding per operation).
"""
import numpy as np

F = np.float32
M32 = 0xFFFFFFFF
M64 = 0xFFFFFFFFFFFFFFFF


def _rotl(v, n):
    return ((v << n) | (v >> (32 - n))) & M32


def chacha_block(key_words, counter, stream, rounds):
    s = [0x61707865, 0x3320646E, 0x79622D32, 0x6B206574] + list(key_words) + [
        counter & M32, (counter >> 32) & M32, stream & M32, (stream >> 32) & M32]
    x = list(s)

    def qr(a, b, c, d):
        x[a] = (x[a] + x[b]) & M32; x[d] = _rotl(x[d] ^ x[a], 16)
        x[c] = (x[c] + x[d]) & M32; x[b] = _rotl(x[b] ^ x[c], 12)
        x[a] = (x[a] + x[b]) & M32; x[d] = _rotl(x[d] ^ x[a], 8)
        x[c] = (x[c] + x[d]) & M32; x[b] = _rotl(x[b] ^ x[c], 7)

    for _ in range(rounds // 2):
        qr(0, 4, 8, 12); qr(1, 5, 9, 13); qr(2, 6, 10, 14); qr(3, 7, 11, 15)
        qr(0, 5, 10, 15); qr(1, 6, 11, 12); qr(2, 7, 8, 13); qr(3, 4, 9, 14)
    return [(a + b) & M32 for a, b in zip(x, s)]


class StdRng:
    """StdRng::seed_from_u64 = PCG32-expanded seed -> ChaCha12, 64-bit counter, stream 0, words in order."""

    def __init__(self, seed):
        state = seed & M64
        key = []
        for _ in range(8):
            state = (state * 6364136223846793005 + 11634580027462260723) & M64
            xorshifted = (((state >> 18) ^ state) >> 27) & M32
            rot = state >> 59
            key.append(((xorshifted >> rot) | (xorshifted << ((32 - rot) & 31))) & M32)
        self.key, self.counter, self.buf = key, 0, []

    def next_u32(self):
        if not self.buf:
            self.buf = chacha_block(self.key, self.counter, 0, 12)
            self.counter += 1
        return self.buf.pop(0)

    def gen_f32(self):
        return F(self.next_u32() >> 8) * F(1.0 / 16777216.0)

    def gen_range(self, low, high):
        rng = (high - low) & M32
        lz = 32 - rng.bit_length()
        zone = (((rng << lz) & M32) - 1) & M32
        while True:
            m = self.next_u32() * rng
            if (m & M32) <= zone:
                return low + (m >> 32)


def build_maze(n, rng):
    parent = []
    edges = []
    grid = [[0] * n for _ in range(n)]
    sets = [[0] * n for _ in range(n)]
    for y in range(n):
        for x in range(n):
            if y != 0:
                edges.append((x, y, True))
            if x != 0:
                edges.append((x, y, False))
            sets[y][x] = len(parent)
            parent.append(None)
    for i in range(len(edges) - 1, 0, -1):
        j = rng.gen_range(0, i + 1)
        edges[i], edges[j] = edges[j], edges[i]

    def root(i):
        while parent[i] is not None:
            i = parent[i]
        return i

    for (x, y, up) in edges:
        nx, ny = (x, y - 1) if up else (x - 1, y)
        a, b = sets[y][x], sets[ny][nx]
        if root(a) != root(b):
            parent[root(b)] = a
            if up:
                grid[y][x] |= 1; grid[ny][nx] |= 2
            else:
                grid[y][x] |= 4; grid[ny][nx] |= 8
    return np.array(grid, dtype=np.uint8)


def extract_walls(n, grid):
    vert, hori = [], []
    for x in range(n):
        start, length = 0, 0
        for y in range(n):
            if x == 0:
                length += 1
                continue
            elif grid[y][x] & 4 == 0 and grid[y][x - 1] & 8 == 0:
                length += 1
            else:
                if length > 0:
                    vert.append((F(x), F(start), F(length)))
                length = 0
                start = y + 1
        vert.append((F(x), F(start), F(length)))
    for y in range(n):
        start, length = 0, 0
        for x in range(n):
            if y == 0:
                length += 1
                continue
            elif grid[y][x] & 1 == 0 and grid[y - 1][x] & 2 == 0:
                length += 1
            else:
                if length > 0:
                    hori.append((F(y), F(start), F(length)))
                length = 0
                start = x + 1
        hori.append((F(y), F(start), F(length)))
    return vert, hori


def assemble_scene(n, vert, hori, rng):
    planes, mats, emis = [], [], []
    wc = (F(0.3), F(0.35), F(0.4))
    base = F(-10.0) * (F(n) / F(2.0))
    half = F(10.0) * (F(n) / F(2.0))
    side = F(10.0) * F(n)
    z, ten = F(0.0), F(10.0)

    def push(o, v, u, c, m, e):
        planes.append((o, v, u, c)); mats.append(m); emis.append(e)

    for (wx, ws, wl) in vert:
        push((base + wx * ten, F(2.0), base + ws * ten), (z, z, wl * ten), (z, F(-10.0), z), wc,
             0 if rng.gen_f32() < F(0.85) else 1, (F(1.0), z, z, z))
        if wl <= F(2.0) and rng.gen_f32() < F(0.3):
            push((base + wx * ten + F(0.1), F(2.0), base + ws * ten), (z, z, F(9.9)), (z, F(-6.0), z), wc, 0,
                 (F(1.0), F(0.8), F(0.3), F(2.0)))
    for (wy, ws, wl) in hori:
        push((base + ws * ten, F(2.0), base + wy * ten), (wl * ten, z, z), (z, F(-10.0), z), wc,
             0 if rng.gen_f32() < F(0.90) else 1, (F(1.0), z, z, z))
        if wl <= F(2.0) and rng.gen_f32() < F(0.3):
            push((base + ws * ten, F(2.0), base + wy * ten + F(0.1)), (F(9.9), z, z), (z, F(-6.0), z), wc, 0,
                 (F(1.0), F(0.8), F(0.3), F(2.0)))
    one4 = (F(1.0), F(1.0), F(1.0), z)
    push((-half, F(2.0), -half), (z, F(-20.0), z), (side, z, z), wc, 0, one4)
    push((-half, F(2.0), half), (side, z, z), (z, F(-20.0), z), wc, 0, one4)
    push((-half, F(2.0), -half), (z, z, side), (z, F(-20.0), z), wc, 0, one4)
    push((half, F(2.0), -half), (z, F(-20.0), z), (z, z, side), wc, 0, one4)
    push((-half, F(2.0), half), (z, z, -side), (side, z, z), (F(0.4), F(0.45), F(0.3)), 0, one4)
    push((F(-5.0), F(2.0), -half + F(0.1)), (ten, z, z), (z, F(-6.0), z), (z, z, z), 0, (F(1.0), F(0.8), F(0.3), F(2.0)))
    push((-half, F(-8.0), half), (z, z, -side), (side, z, z), (z, z, z), 0, (F(1.0), F(0.8), F(0.3), F(0.02)))
    P = np.array([[c for vec in p for c in vec] for p in planes], dtype=F).reshape(-1, 4, 3)
    return P, np.array(mats, dtype=np.uint8), np.array(emis, dtype=F)


def build_bvh(P):
    """P: [n,4,3] float32 (origin, v, u, color).  Literal exhaustive-SAH builder, vectorised over primitives."""
    n = len(P)
    origin, v, u = P[:, 0], P[:, 1], P[:, 2]
    corners = np.stack([origin, origin + u, origin + v], axis=1)          # [n,3,3]
    pmin, pmax = corners.min(axis=1), corners.max(axis=1)
    centers = origin + (u + v) * F(0.5)
    idx = list(range(n))
    nodes = []                                                            # [min3, max3, left_first, tri_count]

    def bounds(first, count):
        ii = idx[first:first + count]
        mn = np.minimum(np.full(3, F(1e30)), pmin[ii].min(axis=0))
        mx = np.maximum(np.full(3, F(-1e30)), pmax[ii].max(axis=0))
        return mn.astype(F), mx.astype(F)

    def area(mn, mx):
        e = (mx - mn).astype(F)
        return F(F(F(e[0] * e[1]) + F(e[1] * e[2])) + F(e[2] * e[0]))

    def subdivide(self_i):
        mn, mx, first, count = nodes[self_i]
        if count == 1:
            return
        ii = np.array(idx[first:first + count])
        best_cost, best_pos, best_axis = F(1e30), F(0.0), 6
        with np.errstate(all="ignore"):
            for axis in range(3):
                c = centers[ii, axis]
                for cand in c:
                    left = c < cand
                    lc, rc = int(left.sum()), int((~left).sum())
                    lmn = np.minimum(F(1e30), pmin[ii[left]].min(axis=0)) if lc else np.full(3, F(1e30))
                    lmx = np.maximum(F(-1e30), pmax[ii[left]].max(axis=0)) if lc else np.full(3, F(-1e30))
                    rmn = np.minimum(F(1e30), pmin[ii[~left]].min(axis=0)) if rc else np.full(3, F(1e30))
                    rmx = np.maximum(F(-1e30), pmax[ii[~left]].max(axis=0)) if rc else np.full(3, F(-1e30))
                    cost = F(F(F(lc) * area(lmn.astype(F), lmx.astype(F))) + F(F(rc) * area(rmn.astype(F), rmx.astype(F))))
                    cost = cost if cost > 0 else F(1e30)
                    if cost <= best_cost:
                        best_cost, best_pos, best_axis = cost, cand, axis
            parent_cost = F(F(count) * area(mn, mx))
        if best_cost > parent_cost:
            return
        i, j = first, first + count - 1
        while i <= j:
            if centers[idx[i], best_axis] < best_pos:
                i += 1
            else:
                idx[i], idx[j] = idx[j], idx[i]
                j -= 1
        left_count = i - first
        if left_count == 0 or left_count == count:
            return
        li = len(nodes)
        nodes.append([*bounds(first, left_count), first, left_count])
        nodes.append([*bounds(i, count - left_count), i, count - left_count])
        subdivide(li)
        subdivide(li + 1)
        nodes[self_i][2] = li
        nodes[self_i][3] = 0

    nodes.append([*bounds(0, n), 0, n])
    subdivide(0)
    out = np.zeros(len(nodes), dtype=[("aabb_min", "<f4", 3), ("aabb_max", "<f4", 3), ("left_first", "<u4"), ("tri_count", "<u4")])
    for k, (mn, mx, lf, tc) in enumerate(nodes):
        out[k] = (mn, mx, lf, tc)
    return out, np.array(idx, dtype=np.uint32)


def build_scene(n, seed=0):
    rng = StdRng(seed)
    grid = build_maze(n, rng)
    vert, hori = extract_walls(n, grid)
    P, mats, emis = assemble_scene(n, vert, hori, rng)
    nodes, indices = build_bvh(P)
    return {"grid": grid, "vert": np.array(vert, dtype=F).reshape(-1, 3), "hori": np.array(hori, dtype=F).reshape(-1, 3),
            "planes": P, "materials": mats, "emissions": emis, "nodes": nodes, "indices": indices}


# ---- frame inputs without the product library (bench.py --impl reference) -------------------------------------------------
# The reference's start-of-run camera / uniform (src/main.rs:732-755, maths.rs:139-156) and chunk order (:293-302), generalised
# to n x n like the product's mm_default_uniform / mm_gen_chunks, in plain Python + libm (the same sinf / cosf / asinf the
# C++ host surface calls), so the reference arm of bench.py builds its whole workload with no product .so mapped.
# tests/test_host_surface.py checks both against the product's, byte for byte.

def _libm():
    import ctypes
    import ctypes.util
    m = ctypes.CDLL(ctypes.util.find_library("m") or "libm.so.6")
    for name in ("sinf", "cosf", "asinf"):
        f = getattr(m, name)
        f.restype, f.argtypes = ctypes.c_float, [ctypes.c_float]
    return m


def calculate_quaternion(direction):
    """maths.rs:139-156 in float32."""
    m = _libm()
    d = np.asarray(direction, dtype=F)
    mag = lambda v: F(np.sqrt(F(F(F(v[0] * v[0]) + F(v[1] * v[1])) + F(v[2] * v[2]))))
    cam = np.array([F(d[0] / mag(d)), F(d[1] / mag(d)), F(d[2] / mag(d))], dtype=F)
    z = np.array([0.0, 0.0, 1.0], dtype=F)
    axis = np.array([F(F(z[1] * cam[2]) - F(z[2] * cam[1])), F(F(z[2] * cam[0]) - F(z[0] * cam[2])), F(F(z[0] * cam[1]) - F(z[1] * cam[0]))], dtype=F)
    la = mag(axis)
    axis_n = np.array([F(axis[0] / la), F(axis[1] / la), F(axis[2] / la)], dtype=F)
    half_theta = F(F(m.asinf(float(la))) / F(2.0))
    s, c = F(m.sinf(float(half_theta))), F(m.cosf(float(half_theta)))
    return np.array([F(axis_n[0] * s), F(axis_n[1] * s), F(axis_n[2] * s), c], dtype=F)


def default_uniform_bytes(maze_n, view_width, view_height, chunk_width=4, time=0):
    """The 56-byte `Uniform` (main.rs:41-49) of the start pose, as bytes: camera (-5, 0, -10*(n/2)+5), focal 1,
    rotation = calculate_quaternion((0.1, 0, 1)), viewport (2*W/H, 2)."""
    import struct
    vw, vh = F(view_width), F(view_height)
    viewport_h = F(2.0)
    viewport_w = F(viewport_h * F(vw / vh))
    q = calculate_quaternion((0.1, 0.0, 1.0))
    center = (F(-5.0), F(0.0), F(F(F(-10.0) * F(F(maze_n) / F(2.0))) + F(5.0)))
    return struct.pack("<3f f 4f 2f 2f 2I", *[float(v) for v in center], 1.0, *[float(v) for v in q], float(viewport_w), float(viewport_h),
                       float(vw), float(vh), int(chunk_width), int(time))


def gen_chunks(view_width, view_height, chunk_width):
    """gen_pixels without the shuffle (main.rs:293-302): x outer, y inner; (u32, u32) pairs."""
    w, h = int(view_width) // chunk_width, int(view_height) // chunk_width
    out = np.zeros((w * h, 2), dtype=np.uint32)
    out[:, 0] = np.repeat(np.arange(w, dtype=np.uint32) * chunk_width, h)
    out[:, 1] = np.tile(np.arange(h, dtype=np.uint32) * chunk_width, w)
    return out
