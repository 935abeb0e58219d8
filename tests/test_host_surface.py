"""Kept host surface (maze -> walls -> scene -> BVH -> camera/uniform/chunks), C++ restatement of reference
src/main.rs:91-263, 293-302, 328-352, 357-588, 732-755 and src/maths.rs:139-178.  The reference has no tests, so
these are: an independent Python restatement (oracle/host_ref.py) compared array for array, structural invariants
(SURVEY §4), the reference's literals at its own size n = 10, and literal-vs-fast BVH builder equality."""
import math

import numpy as np
import pytest

F = np.float32


def planes_as_array(sc):
    return np.stack([sc.planes["origin"], sc.planes["v"], sc.planes["u"], sc.planes["color"]], axis=1)


@pytest.mark.parametrize("n", [1, 2, 3, 4, 7, 10, 16])
def test_cpp_matches_python_restatement(mm, n):
    from oracle import host_ref

    ref = host_ref.build_scene(n, 0)
    sc = mm.MazeScene(n, 0, fast_bvh=False)
    assert np.array_equal(ref["grid"], sc.grid)
    assert np.array_equal(ref["vert"], sc.vert_walls) and np.array_equal(ref["hori"], sc.hori_walls)
    assert planes_as_array(sc).tobytes() == ref["planes"].tobytes()
    assert np.array_equal(ref["materials"], sc.materials)
    assert ref["emissions"].tobytes() == sc.emissions.tobytes()
    assert ref["nodes"].tobytes() == sc.nodes.tobytes()
    assert np.array_equal(ref["indices"], sc.indices)


@pytest.mark.parametrize("n,seed", [(5, 0), (10, 0), (16, 3), (32, 0), (48, 9)])
def test_fast_bvh_builder_emits_the_literal_arrays(mm, n, seed):
    a = mm.MazeScene(n, seed, fast_bvh=False)
    b = mm.MazeScene(n, seed, fast_bvh=True)
    assert a.nodes.tobytes() == b.nodes.tobytes()
    assert np.array_equal(a.indices, b.indices)


def test_fast_bvh_on_adversarial_planes(mm):
    """Coincident centres, zero-area planes and equal costs exercise the `<=` tie rule (main.rs:123)."""
    from mirror_maze_b200.host import PLANE_DTYPE, build_bvh

    rng = np.random.default_rng(5)
    P = np.zeros(300, dtype=PLANE_DTYPE)
    P["origin"] = rng.integers(-4, 5, size=(300, 3)).astype(F) * F(10)
    P["v"] = rng.integers(0, 3, size=(300, 3)).astype(F) * F(10) * (rng.integers(0, 2, size=(300, 3)))
    P["u"] = rng.integers(-2, 1, size=(300, 3)).astype(F) * F(10) * (rng.integers(0, 2, size=(300, 3)))
    na, ia = build_bvh(P, fast=False)
    nb, ib = build_bvh(P, fast=True)
    assert na.tobytes() == nb.tobytes() and np.array_equal(ia, ib)


@pytest.mark.parametrize("n", [10, 16, 32, 64])
def test_bvh_structural_invariants(scenes, n):
    sc = scenes(n)
    nodes, idx = sc.nodes, sc.indices
    assert sorted(idx.tolist()) == list(range(sc.n_planes))               # a permutation
    seen_planes = []
    origin, u, v = sc.planes["origin"], sc.planes["u"], sc.planes["v"]
    corners = np.stack([origin, origin + u, origin + v], axis=1)

    def walk(i, depth):
        nd = nodes[i]
        if nd["tri_count"] > 0:
            ids = idx[nd["left_first"]: nd["left_first"] + nd["tri_count"]]
            seen_planes.extend(ids.tolist())
            c = corners[ids].reshape(-1, 3)
            assert np.array_equal(c.min(axis=0), nd["aabb_min"]) and np.array_equal(c.max(axis=0), nd["aabb_max"])
            return depth
        l, r = int(nd["left_first"]), int(nd["left_first"]) + 1           # children adjacent (main.rs:162-168)
        assert l % 2 == 1 and r < len(nodes)
        for c in (l, r):
            assert (nodes[c]["aabb_min"] >= nd["aabb_min"]).all() and (nodes[c]["aabb_max"] <= nd["aabb_max"]).all()
        return max(walk(l, depth + 1), walk(r, depth + 1))

    depth = walk(0, 1)
    assert sorted(seen_planes) == list(range(sc.n_planes))                # every plane in exactly one leaf
    assert depth <= 48
    assert len(nodes) % 2 == 1 and len(nodes) <= 2 * sc.n_planes - 1


def test_reference_size_scene_literals(mm):
    """At n = 10 the generalised scene must reproduce the reference's hard-wired numbers (main.rs:517-586, 735)."""
    sc = mm.MazeScene(10, 0)
    P = sc.planes
    tail = P[-7:]
    assert tail[0]["origin"].tolist() == [-50.0, 2.0, -50.0] and tail[0]["v"].tolist() == [0.0, -20.0, 0.0] and tail[0]["u"].tolist() == [100.0, 0.0, 0.0]
    assert tail[1]["origin"].tolist() == [-50.0, 2.0, 50.0] and tail[1]["v"].tolist() == [100.0, 0.0, 0.0]
    assert tail[2]["v"].tolist() == [0.0, 0.0, 100.0] and tail[3]["origin"].tolist() == [50.0, 2.0, -50.0]
    assert tail[4]["v"].tolist() == [0.0, 0.0, -100.0] and np.allclose(tail[4]["color"], [0.4, 0.45, 0.3])
    assert tail[5]["origin"][2] == F(-49.9) and tail[5]["origin"][0] == F(-5.0)        # entry light
    assert tail[6]["origin"].tolist() == [-50.0, -8.0, 50.0]                            # roof
    assert sc.emissions[-1].tolist() == [1.0, F(0.8), F(0.3), F(0.02)]
    assert sc.emissions[-2].tolist() == [1.0, F(0.8), F(0.3), 2.0]
    u = mm.default_uniform(10, 1024, 768, 4)
    assert [u.cam.camera_center.x, u.cam.camera_center.y, u.cam.camera_center.z] == [-5.0, 0.0, -45.0]
    assert u.cam.focal_length == 1.0 and u.cam.viewport.y == 2.0 and u.cam.viewport.x == F(2.0) * (F(1024) / F(768))


def test_maze_is_a_spanning_tree(scenes):
    for n in (10, 32):
        g = scenes(n).grid
        opened = sum(bin(int(c)).count("1") for c in g.ravel()) // 2
        assert opened == n * n - 1                                         # Kruskal: exactly n^2 - 1 passages
        for y in range(n):
            for x in range(n):
                if g[y, x] & 1: assert y > 0 and g[y - 1, x] & 2
                if g[y, x] & 4: assert x > 0 and g[y, x - 1] & 8
        # connectivity
        seen, todo = {(0, 0)}, [(0, 0)]
        while todo:
            x, y = todo.pop()
            for bit, dx, dy in ((1, 0, -1), (2, 0, 1), (4, -1, 0), (8, 1, 0)):
                if g[y, x] & bit and (x + dx, y + dy) not in seen:
                    seen.add((x + dx, y + dy)); todo.append((x + dx, y + dy))
        assert len(seen) == n * n


def test_walls_cover_closed_edges_and_keep_degenerate_runs(scenes):
    sc = scenes(16)
    n, g = 16, sc.grid
    closed_v = sum(1 for x in range(1, n) for y in range(n) if not (g[y, x] & 4))
    assert int(sc.vert_walls[sc.vert_walls[:, 0] > 0][:, 2].sum()) == closed_v
    assert sc.vert_walls[0].tolist() == [0.0, 0.0, float(n)]                # x == 0: one full-height wall
    assert (sc.vert_walls[:, 2] == 0).any() or (sc.hori_walls[:, 2] == 0).any()   # trailing zero-length runs are kept (:416,437)


def test_scene_statistics_are_plausible(scenes):
    sc = scenes(64)
    walls = sc.emissions[:, 3] == 0.0
    frac_mirror = sc.materials[walls].mean()
    assert 0.05 < frac_mirror < 0.2                                        # 15 % / 10 % mirror odds (main.rs:460,494)
    assert (sc.emissions[:, 3] == 2.0).sum() > 100                         # light panels on short walls


def test_quaternion_and_camera(mm):
    q = mm.calculate_quaternion([0.1, 0.0, 1.0])                          # main.rs:740
    assert abs(float(np.linalg.norm(q)) - 1.0) < 1e-6 and q[0] == 0.0 and q[2] == 0.0 and q[1] > 0
    ht = math.acos(float(q[3]))
    assert abs(2 * ht - math.atan2(0.1, 1.0)) < 1e-5        # acos of an f32 near 1 loses bits
    fwd = mm.quat_mult([0.0, 0.0, 1.0], q)
    assert abs(float(np.linalg.norm(fwd)) - 1.0) < 1e-6
    q2 = mm.update_quat_angle(q, 0.7)
    assert abs(float(q2[3]) - math.cos(0.7)) < 1e-6 and abs(float(q2[1]) - math.sin(0.7)) < 1e-5


def test_chunk_order_is_gen_pixels_without_shuffle(mm):
    ch = mm.gen_chunks(16, 12, 4)
    assert [(int(c["x"]), int(c["y"])) for c in ch] == [(4 * i, 4 * j) for i in range(4) for j in range(3)]   # main.rs:298-302
    assert len(mm.gen_chunks(1920, 1080, 4)) == 480 * 270


def test_check_collision(mm, scenes):
    sc = scenes(10)
    u = mm.default_uniform(10, 64, 64)
    c = np.array([u.cam.camera_center.x, u.cam.camera_center.y, u.cam.camera_center.z], dtype=F)
    d = np.array([0.5, 0.2, 0.5], dtype=F)                                 # player_diag main.rs:738
    assert mm.check_collision(sc.nodes, c - d, c + d) == -1                # the start cell is free
    wall_x = np.array([-50.0, 0.0, -45.0], dtype=F)                        # on the outer x = -50 wall
    assert mm.check_collision(sc.nodes, wall_x - d, wall_x + d) >= 0


def _literal_edge_test(x, L):
    """shaders.metal:60-63 for one edge: d = x / L (IEEE fp32), 0 <= d && d <= L."""
    with np.errstate(divide="ignore", invalid="ignore", over="ignore", under="ignore"):
        d = (x.astype(np.float32) / np.float32(L)).astype(np.float32)
    return (np.float32(0.0) <= d) & (d <= np.float32(L))


def test_rect_edge_thresholds_equal_divide_then_compare(mm):
    """The kernel's divide-free edge test: lo <= x <= up must decide exactly what 0 <= RN(x/L) <= L decides, for x at and
    around both boundaries (several ulps each side), zeros of both signs, denormals, infinities, NaN and random x."""
    rng = np.random.default_rng(1234)
    lengths = [10.0, 5.0, 4.9, 0.1, 0.2, 1.0, 2.0, 3.0, 1.9999999, 2.0000002, 2.0 ** -20, 2.0 ** 23, 7.5, 1e-3, 12345.678,
               1.0000001, 0.99999994, 16777215.0 / 4, 1e6]
    lengths += list(np.exp(rng.uniform(np.log(2.0 ** -20), np.log(2.0 ** 23), 4000)))
    checked = 0
    for L in lengths:
        L = np.float32(L)
        th = mm.rect_edge_thresholds(L)
        assert th is not None, L
        lo, up = th
        xs = [np.float32(0.0), np.float32(-0.0), np.float32(np.inf), np.float32(-np.inf), np.float32(np.nan), lo, up, L, L * L]
        for base in (up, lo, np.float32(L * L)):
            v = np.float32(base)
            a = b = v
            for _ in range(6):
                a = np.nextafter(a, np.float32(np.inf)); b = np.nextafter(b, np.float32(-np.inf))
                xs += [a, b]
        den = np.float32(1.401298464324817e-45)
        xs += [np.float32(-k) * den for k in range(0, 40)] + [np.float32(k) * den for k in range(1, 4)]
        xs += list((rng.standard_normal(24) * float(L) * float(L)).astype(np.float32))
        x = np.array(xs, dtype=np.float32)
        want = _literal_edge_test(x, L)
        got = (lo <= x) & (x <= up)
        assert np.array_equal(want, got), (L, lo, up, x[want != got][:5])
        checked += len(x)
    assert checked > 300000
    # degenerate and unguarded lengths
    lo, up = mm.rect_edge_thresholds(0.0)
    assert np.isnan(lo) and np.isnan(up)                       # x / 0 never passes the literal test either
    x = np.array([0.0, -0.0, 1.0, -1.0, np.inf, np.nan], dtype=np.float32)
    assert not _literal_edge_test(x, np.float32(0.0)).any()
    for bad in (1e-30, 2.0 ** -21, 2.0 ** 24, np.inf, np.nan, -1.0, 1e-45):
        assert mm.rect_edge_thresholds(bad) is None


def test_reference_arm_inputs_equal_the_products(mm):
    """bench.py --impl reference builds its workload with oracle/host_ref.py only (no product library mapped): the uniform
    bytes and the chunk list must be the product's (mm_default_uniform / mm_gen_chunks), and so must the scene arrays."""
    from oracle import host_ref

    for n, W, H, t in ((32, 1920, 1080, 0), (64, 3840, 2160, 5), (10, 1024, 768, 3), (256, 1920, 1080, 1)):
        assert bytes(mm.default_uniform(n, W, H, 4, t)) == host_ref.default_uniform_bytes(n, W, H, 4, t)
        assert mm.gen_chunks(W, H, 4).tobytes() == host_ref.gen_chunks(W, H, 4).tobytes()
    ref, sc = host_ref.build_scene(32, 0), mm.MazeScene(32, 0)
    assert np.ascontiguousarray(ref["planes"]).tobytes() == np.ascontiguousarray(sc.planes).tobytes()
    assert np.ascontiguousarray(ref["nodes"]).tobytes() == np.ascontiguousarray(sc.nodes).tobytes()
    assert np.array_equal(ref["indices"], sc.indices) and np.array_equal(ref["materials"], sc.materials)
    assert np.ascontiguousarray(ref["emissions"], dtype=np.float32).tobytes() == np.ascontiguousarray(sc.emissions, dtype=np.float32).tobytes()


def test_axis_aligned_rect_test_decides_like_the_literal_one(mm, oracle):
    """mm_axis_rect: for axis-aligned rects the kernel evaluates a = RN(RN(c - o_k) / d_k) and interval tests on the intersection
    point's in-plane coordinates instead of ray_rect_intersect (shaders.metal:51-67).  Checked against the oracle's literal
    function: random rays, rays aimed within a few ulps of every rect border, rays from points on the rect's plane, zero
    direction components; every wall / floor / roof / panel shape of the mazes plus odd sizes, both windings, all three axes."""
    from mirror_maze_b200.host import PLANE_DTYPE

    F = np.float32
    rng = np.random.default_rng(7)

    def collapsed(rec, o, d, t):
        c, lo_a, hi_a, lo_b, hi_b, k = rec
        if k == 3:
            return False, t
        a_ax, b_ax = (1 if k == 0 else 0), (1 if k == 2 else 2)
        with np.errstate(all="ignore"):
            a = F(F(c - o[k]) / d[k])
            pa = F(o[a_ax] + F(d[a_ax] * a)); pb = F(o[b_ax] + F(d[b_ax] * a))
        hit = bool(lo_a <= pa <= hi_a and lo_b <= pb <= hi_b and d[k] != 0 and a > F(0.1) and a < t)
        return hit, (float(a) if hit else t)

    shapes = [((-50.0, 2.0, -30.0), (0, 0, 20.0), (0, -10.0, 0)), ((-50.0, 2.0, -30.0), (30.0, 0, 0), (0, -10.0, 0)),
              ((-160.0, 2.0, -160.0), (320.0, 0, 0), (0, 0, 320.0)), ((-160.0, -8.0, -160.0), (0, 0, 320.0), (320.0, 0, 0)),
              ((-4.95, 1.0, -49.9), (9.9, 0, 0), (0, -6.0, 0)), ((12.5, 0.37, 3.1), (0, 0, -7.3), (0, 1.9, 0)),
              ((0.0, 2.0, 0.0), (0, 0, 10.0), (0, -10.0, 0)), ((1280.0, 2.0, -1270.0), (0, -10.0, 0), (-1e-3, 0, 0))]
    checked = hits = 0
    for origin, v, u in shapes:
        P = np.zeros(1, dtype=PLANE_DTYPE)
        P[0]["origin"], P[0]["v"], P[0]["u"], P[0]["color"] = origin, v, u, (0.5, 0.5, 0.5)
        rec = mm.axis_rect(P[0])
        assert rec is not None and rec[5] in (0, 1, 2)
        k = rec[5]
        org, vv, uu = np.array(origin, F), np.array(v, F), np.array(u, F)
        targets = []
        for sv in (0.0, 1.0, 0.5, 0.25):                       # points on / around the borders and inside
            for su in (0.0, 1.0, 0.5, 0.75):
                base = org + F(sv) * vv + F(su) * uu
                for _ in range(6):
                    jit = base.copy()
                    for j in range(3):
                        jit[j] = np.nextafter(jit[j], F(np.inf if rng.random() < 0.5 else -np.inf)) if rng.random() < 0.7 else jit[j]
                        if rng.random() < 0.3:
                            jit[j] = F(jit[j] + F(rng.normal() * 1e-5 * max(1.0, abs(float(jit[j])))))
                    targets.append(jit)
        for tgt in targets:
            for _ in range(5):
                o = (tgt + rng.normal(size=3).astype(F) * F(rng.choice([0.3, 5.0, 80.0]))).astype(F)
                if rng.random() < 0.15:
                    o[k] = org[k]                               # ray origin on the rect's plane
                d = (tgt - o).astype(F)
                n = F(np.sqrt(float(d @ d))) or F(1.0)
                d = (d / n).astype(F) if rng.random() < 0.8 else d
                if rng.random() < 0.1:
                    d[int(rng.integers(0, 3))] = F(0.0)
                for t in (F(1e30), F(rng.uniform(0.05, 200.0))):
                    lit = oracle.ray_rect(o, d, float(t), P[0])
                    col = collapsed(rec, o, d, t)
                    assert lit[0] == col[0] and (not lit[0] or np.float32(lit[1]) == np.float32(col[1])), (origin, v, u, o, d, t, lit, col)
                    checked += 1
                    hits += int(lit[0])
    assert checked > 7000 and checked // 10 < hits < checked - checked // 10          # both outcomes, plentifully
    # degenerate rects (zero-length walls, main.rs:416,437) never hit in either form; tilted rects are refused
    Z = np.zeros(1, dtype=PLANE_DTYPE); Z[0]["origin"], Z[0]["v"], Z[0]["u"] = (1, 2, 3), (0, 0, 0), (0, -10, 0)
    assert mm.axis_rect(Z[0])[5] == 3
    T = np.zeros(1, dtype=PLANE_DTYPE); T[0]["origin"], T[0]["v"], T[0]["u"] = (1, 2, 3), (1, 0, 1), (0, -10, 0)
    assert mm.axis_rect(T[0]) is None
    # every rect of the mazes has the axis-aligned form
    for n_maze in (10, 32):
        sc = mm.MazeScene(n_maze, 0)
        assert all(mm.axis_rect(pl) is not None for pl in sc.planes)
