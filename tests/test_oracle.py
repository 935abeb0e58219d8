"""Pins the CPU oracle (oracle/mm_oracle.cpp).  The reference has no tests or golden vectors for this path and cannot
run here (SURVEY §4, §8 c: PARITY UNPINNED by the reference), so the pins are: the hand-derived known-answer vectors
of SURVEY Appendix E for the seed and PCG hash (reference src/shaders.metal:181-186,298), an independent numpy
transcription of the shader compared path by path, unit edge cases of the slab and rect tests (:51-67, :87-95), and the
committed fixtures under tests/golden/ (regenerate with tests/golden/make_golden.py)."""
import hashlib
import json
import os

import numpy as np
import pytest

from cases import CASES, build_case

F = np.float32
HERE = os.path.dirname(os.path.abspath(__file__))
GOLDEN = json.load(open(os.path.join(HERE, "golden", "golden.json")))

# (texid_x, texid_y, time) -> seed, state after first call, r1, r2, r3
KAT = [
    ((0, 0, 0), 1, 1039132858, 3160415554, 1313351202, 1589324906),
    ((1, 0, 0), 15824, 786748693, 1731853957, 1895800264, 4033427381),
    ((0, 1, 0), 9737334, 2126070387, 2322506250, 1236482286, 3051696906),
    ((100, 0, 7), 1582308, 2510127673, 3879997200, 3231775265, 3820095257),
    ((31, 31, 0), 302347840, 3383872581, 3691425956, 512067647, 533278338),
    ((1023, 767, 0), 3189754112, 3863318021, 750421190, 4285080687, 2814865408),
    ((1023, 767, 5), 3189754112, 3863318021, 750421190, 4285080687, 2814865408),     # `time` lost to fp32 rounding
    ((1919, 1079, 0), 1947011968, 3186561669, 2133415391, 4222260886, 2817372876),
    ((400, 441, 0), 4294967295, 3838507344, 1880471250, 2083244681, 3757147803),     # float sum > 2^32: saturates
]


def digest(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


@pytest.mark.parametrize("texid,seed,state1,r1,r2,r3", KAT)
def test_seed_and_pcg_known_answers(oracle, texid, seed, state1, r1, r2, r3):
    n = F(128) / F(255)                                                    # noise texel (0,0) = 128 in every channel
    assert oracle.seed(n, n, texid[0], texid[1], texid[2]) == seed
    words, _ = oracle.random_words(seed, 3)
    assert words == [r1, r2, r3]
    _, st = oracle.random_words(seed, 1)
    assert st == state1


def test_random_float_conversion(oracle):
    assert oracle.random_floats(1, 1)[0] == float(F(3160415554) * F(2.0 ** -32))   # 0.7358416 (Appendix E)
    assert abs(oracle.random_floats(1, 1)[0] - 0.7358416) < 1e-7
    vals = oracle.random_floats(12345, 4000)
    assert 0.0 <= min(vals) and max(vals) <= 1.0 and 0.47 < np.mean(vals) < 0.53


def test_noise_texel_zero_is_the_only_one_sampled(mm, noise):
    assert noise[0, 0].tolist() == [128, 128, 128, 255]
    # integer thread coordinates with normalised coords + repeat + nearest wrap to texel (0,0): changing every other
    # texel must not change a single output bit
    from oracle import oracle as o

    sc, u, p, ch = build_case(mm, "chunk2_spp4")
    a = o.render(sc, noise, u, p, ch)[0]
    other = noise.copy()
    other[1:, :, :] = 7
    other[0, 1:, :] = 9
    b = o.render(sc, other, u, p, ch)[0]
    assert a.tobytes() == b.tobytes()
    other[0, 0, 0] = 3
    assert o.render(sc, other, u, p, ch)[0].tobytes() != a.tobytes()


def test_slab_edge_cases(oracle):
    from oracle import np_oracle

    inf = float("inf")
    box = ([-1.0, -1.0, 2.0], [1.0, 1.0, 4.0])
    cases = [
        ([0, 0, 0], [0, 0, 1], 1e30, 2.0),          # zero dir components: +-inf slabs, inside in x and y
        ([2, 0, 0], [0, 0, 1], 1e30, 1e30),         # outside in x with dir.x == 0: -inf..-inf -> miss
        ([1, 0, 0], [0, 0, 1], 1e30, 1e30),         # on the x = 1 plane with dir.x == 0: 0/0 = NaN is dropped, leaving -inf..-inf
        ([0, 0, 3], [0, 0, 1], 1e30, -1.0),         # origin inside: negative tmin is returned
        ([0, 0, 5], [0, 0, 1], 1e30, 1e30),         # behind the ray: tmax <= 0
        ([0, 0, 0], [0, 0, 1], 1.5, 1e30),          # farther than beam.t
        ([0, 0, 0], [1e-30, 0, 1], 1e30, 2.0),      # denormal-ish direction
        ([0, 0, 0], [float("nan"), 0, 1], 1e30, 2.0),   # NaN direction component is ignored by fmin/fmax
    ]
    for o, d, t, want in cases:
        got = oracle.intersect_aabb(o, d, t, *box)
        assert got == F(want), (o, d, t, got)
        with np.errstate(all="ignore"):
            ref = np_oracle._aabb([np.array([F(v)]) for v in o], [np.array([F(v)]) for v in d], np.array([F(t)]),
                                  [np.array([F(v)]) for v in box[0]], [np.array([F(v)]) for v in box[1]])[0]
        assert got == ref
    # zero-thickness box (a wall): hit exactly on the plane
    assert oracle.intersect_aabb([0, 0, 0], [0.5, 0.1, 1], 1e30, [-5, -5, 3], [5, 5, 3]) == F(3.0)
    assert inf > 0


def test_rect_edge_cases(mm, oracle):
    from mirror_maze_b200.host import PLANE_DTYPE

    r = np.zeros(1, dtype=PLANE_DTYPE)
    r["origin"], r["v"], r["u"] = [0, 2, 5], [0, 0, 10], [0, -10, 0]       # a vertical wall as the scene builds them
    hit, t = oracle.ray_rect([-3, 0, 7], [1, 0, 0], 1e30, r[0])
    assert hit and t == 3.0
    assert not oracle.ray_rect([-3, 0, 7], [1, 0, 0], 2.5, r[0])[0]        # farther than beam.t (strict <)
    assert not oracle.ray_rect([-3, 0, 7], [1, 0, 0], 3.0, r[0])[0]        # tie: first visited wins
    assert not oracle.ray_rect([-0.05, 0, 7], [1, 0, 0], 1e30, r[0])[0]    # a > 0.1 self-hit epsilon
    assert not oracle.ray_rect([-3, 0, 7], [0, 0, 1], 1e30, r[0])[0]       # parallel: norm_check == 0
    assert oracle.ray_rect([-3, 2, 5], [1, 0, 0], 1e30, r[0])[0]           # corner is inside (0 <= d <= len)
    assert not oracle.ray_rect([-3, 2.001, 5], [1, 0, 0], 1e30, r[0])[0]
    z = np.zeros(1, dtype=PLANE_DTYPE)
    z["origin"], z["v"], z["u"] = [0, 2, 5], [0, 0, 0], [0, -10, 0]        # zero-length wall: NaN normal, never hit
    assert not oracle.ray_rect([-3, 0, 5], [1, 0, 0], 1e30, z[0])[0]


def test_quat_mult_device_grouping(mm, oracle):
    q = mm.calculate_quaternion([0.1, 0.0, 1.0])
    v = oracle.quat_mult([0.0, 0.0, 1.0], q)
    assert abs(float(np.linalg.norm(v)) - 1.0) < 1e-6
    assert np.allclose(v, mm.quat_mult([0.0, 0.0, 1.0], q), atol=1e-6)     # host twin groups differently (maths.rs:171)


@pytest.mark.parametrize("name", sorted(CASES))
def test_oracle_matches_golden_fixture(mm, oracle, noise, scenes, name):
    sc, u, p, ch = build_case(mm, name, scenes)
    img, cnt, dbg = oracle.render(sc, noise, u, p, ch, debug=True)
    g = GOLDEN[name]
    assert sc.n_planes == g["planes"] and sc.n_nodes == g["nodes"]
    for k, v in g["counters"].items():
        assert cnt[k] == v, k
    for k in ("first_hit", "segments", "mirror_hits", "radiance"):
        assert digest(dbg[k]) == g[k], k
    assert digest(img) == g["image"]


@pytest.mark.parametrize("name", ["cfg1", "yaw", "chunk3_spp32", "mirror_limit2", "ragged"])
def test_numpy_transcription_agrees_bit_for_bit(mm, oracle, noise, scenes, name):
    from oracle import np_oracle

    sc, u, p, ch = build_case(mm, name, scenes)
    img, cnt, dbg = oracle.render(sc, noise, u, p, ch, debug=True)
    img2, cnt2, dbg2 = np_oracle.render(sc, noise, u, p, ch)
    for k in dbg:
        assert dbg[k].tobytes() == dbg2[k].tobytes(), k
    assert img.tobytes() == img2.tobytes()
    for k in cnt2:
        assert cnt[k] == cnt2[k], k


def test_thread_count_does_not_change_results(mm, oracle, noise, scenes):
    sc, u, p, ch = build_case(mm, "cfg2_small", scenes)
    a = oracle.render(sc, noise, u, p, ch, threads=1)
    b = oracle.render(sc, noise, u, p, ch, threads=4)
    assert a[0].tobytes() == b[0].tobytes() and a[1] == b[1]


def test_group_partition_is_bit_identical_to_full_grid(mm, oracle, noise, scenes):
    """Seeds depend on the virtual group index, not on which call renders it (multi-GPU tile contract)."""
    sc, u, p, ch = build_case(mm, "yaw", scenes)
    full, cnt, _ = oracle.render(sc, noise, u, p, ch)
    n_groups = p.grid_x * p.grid_y
    acc = np.zeros_like(full)
    rays = 0
    for rank in range(3):
        first, step, count = mm.tile_partition(n_groups, rank, 3)
        q = mm.Params.from_buffer_copy(bytes(p))
        q.group_first, q.group_step, q.group_count = first, step, count
        part, c, _ = oracle.render(sc, noise, u, q, ch, out=acc)
        rays += c["rays"]
    assert acc.tobytes() == full.tobytes() and rays == cnt["rays"]


def test_reduction_generalisation_below_8_spp(mm, oracle, noise, scenes):
    """spp = 1: the pixel is its single tone-mapped sample (sqrt(max(L,0)) / 1)."""
    sc, u, p, ch = build_case(mm, "cfg1", scenes)
    img, _, dbg = oracle.render(sc, noise, u, p, ch, debug=True)
    L = dbg["radiance"].reshape(-1, 16, 3)                                 # [group, pixel_number, rgb]
    g0 = np.sqrt(np.maximum(L[0], 0))
    x0, y0 = int(ch[0]["x"]), int(ch[0]["y"])
    for pn in range(16):
        assert img[y0 + pn % 4, x0 + pn // 4, :3].tobytes() == g0[pn].tobytes()   # x uses /, y uses % (shaders.metal:273-274)
    assert (img[..., 3] == 1.0).all()


def test_ragged_frame_leaves_border_unwritten(mm, oracle, noise, scenes):
    sc, u, p, ch = build_case(mm, "ragged", scenes)
    img = oracle.render(sc, noise, u, p, ch)[0]
    assert (img[:, 48:, :] == 0).all() and (img[28:, :, :] == 0).all() and (img[:28, :48, 3] == 1).all()


def test_oracle_rejects_ill_defined_dispatch_shapes(mm, oracle, noise, scenes):
    """T = chunk^2 * spp must be <= 32 or a multiple of 32: the virtual threadgroup is (32, T/32) (main.rs:641-644)."""
    sc = scenes(10)
    u5 = mm.default_uniform(10, 40, 30, 5)
    with pytest.raises(RuntimeError):
        oracle.render(sc, noise, u5, mm.full_frame_params(u5, spp=2, bounce_limit=2), mm.gen_chunks(40, 30, 5))


@pytest.mark.parametrize("name", ["cfg1", "yaw", "on_plane", "tiny_origin"])
def test_rcp_slab_variant_numpy_agrees(mm, oracle, noise, scenes, name):
    """MM_FLAG_RCP_SLAB (opt-in t = (b - o) * (1/d)): the two oracle transcriptions agree bit for bit in this mode too, and
    the mode is a different arithmetic (some radiance bits differ from the literal mode on a large enough case)."""
    from oracle import np_oracle

    sc, u, p, ch = build_case(mm, name, scenes)
    lit = oracle.render(sc, noise, u, p, ch, debug=True)
    p.flags = mm.FLAG_RCP_SLAB
    img, cnt, dbg = oracle.render(sc, noise, u, p, ch, debug=True)
    img2, cnt2, dbg2 = np_oracle.render(sc, noise, u, p, ch)
    for k in dbg:
        assert dbg[k].tobytes() == dbg2[k].tobytes(), k
    assert img.tobytes() == img2.tobytes()
    for k in cnt2:
        assert cnt[k] == cnt2[k], k
    same_first = (dbg["first_hit"] == lit[2]["first_hit"]).mean()
    assert same_first > 0.999          # primary hits essentially never depend on the last ulp of a slab quotient


def test_rejection_threshold_equals_length_test():
    """The kernel's rejection loop compares the squared length with 1 + 2^-23 instead of taking the square root
    (shaders.metal:315-318 `while (length(rd) > 1)`).  Equivalence for every fp32 in [0.5, 2); outside, monotonicity of
    the correctly rounded square root decides (sqrt(s) < 1 for s < 0.5, > 1 for s >= 2)."""
    bits = np.arange(np.float32(0.5).view(np.uint32), np.float32(2.0).view(np.uint32), dtype=np.uint32)
    s = bits.view(np.float32)
    assert np.array_equal(np.sqrt(s) > np.float32(1.0), s > np.float32(1.00000011920928955))
    assert np.float32(1.00000011920928955) == np.float32(1.0) + np.float32(2.0 ** -23)


def test_fused_random_direction_component_is_bit_identical():
    """The kernel evaluates (random(state) - 0.5) * 2.0 as fma(float(r), 2^-31, -1): both scalings are exact, so the
    single rounding of the FMA equals the rounding of the subtraction."""
    rng = np.random.default_rng(0)
    r = np.concatenate([rng.integers(0, 2 ** 32, 1 << 24, dtype=np.uint64),
                        np.array([0, 1, 2, 2 ** 31 - 1, 2 ** 31, 2 ** 31 + 1, 2 ** 32 - 1, 2 ** 32 - 128, 2 ** 24, 2 ** 24 + 1, 2 ** 25 + 1], dtype=np.uint64)]).astype(np.uint32)
    f = r.astype(np.float32)
    literal = ((f * np.float32(2.0 ** -32)) - np.float32(0.5)) * np.float32(2.0)
    fused = (f.astype(np.float64) * 2.0 ** -31 - 1.0).astype(np.float32)        # the exact value, rounded once
    assert np.array_equal(literal.view(np.uint32), fused.view(np.uint32))
