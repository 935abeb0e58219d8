"""Parity of the CUDA path against the CPU oracle, through the C-ABI (mm_render etc.), on a real B200.

Bar (BASELINE.json north_star): first-hit primitive ids and bounce (segment) counts bit-exact; radiance within
max-abs 1e-3 per channel in fp32.  The kernel obeys the oracle's canonical arithmetic, so these tests assert the
stronger property — every observable, every pixel and every counter bit-identical — and keep the 1e-3 bound as the
documented fallback tolerance (TOL) in the one place a looser comparison is meaningful (full-size property checks)."""
import hashlib
import json
import os

import numpy as np
import pytest

from cases import CASES, REF_SHADER_CASES, build_case

pytestmark = pytest.mark.gpu
TOL = 1e-3
HERE = os.path.dirname(os.path.abspath(__file__))
GOLDEN = json.load(open(os.path.join(HERE, "golden", "golden.json")))
COUNTER_KEYS = ("paths", "rays", "inner_visits", "leaf_visits", "rect_tests", "hits", "max_stack")


def digest(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def assert_same(got, ref):
    img, cnt, dbg = got
    rimg, rcnt, rdbg = ref
    assert np.array_equal(dbg["first_hit"], rdbg["first_hit"]), "first-hit primitive ids differ"
    assert np.array_equal(dbg["segments"], rdbg["segments"]), "bounce (segment) counts differ"
    assert np.array_equal(dbg["mirror_hits"], rdbg["mirror_hits"])
    assert np.nanmax(np.abs(dbg["radiance"] - rdbg["radiance"]), initial=0.0) <= TOL
    assert dbg["radiance"].tobytes() == rdbg["radiance"].tobytes(), "radiance not bit-identical"
    assert img.tobytes() == rimg.tobytes(), "image not bit-identical"
    for k in COUNTER_KEYS:
        assert cnt[k] == rcnt[k], k


@pytest.mark.parametrize("name", sorted(CASES))
def test_cuda_matches_oracle(mm, oracle, noise, scenes, renderer, name):
    sc, u, p, ch = build_case(mm, name, scenes)
    renderer.upload_scene(sc, noise)
    ref = oracle.render(sc, noise, u, p, ch, debug=True)
    got = renderer.render(u, p, ch, debug=True)
    assert_same(got, ref)


@pytest.mark.parametrize("name", sorted(CASES))
def test_cuda_matches_committed_golden(mm, noise, scenes, renderer, name):
    """Same check without the oracle's .so: against the fixture committed under tests/golden/."""
    sc, u, p, ch = build_case(mm, name, scenes)
    renderer.upload_scene(sc, noise)
    img, cnt, dbg = renderer.render(u, p, ch, debug=True)
    g = GOLDEN[name]
    for k in ("first_hit", "segments", "mirror_hits", "radiance"):
        assert digest(dbg[k]) == g[k], k
    assert digest(img) == g["image"]
    for k, v in g["counters"].items():
        assert cnt[k] == v, k


@pytest.mark.parametrize("flags_name", ["literal", "counters_only", "literal_counters", "regroup", "regroup_literal", "regroup_counters", "general_rects",
                                        "general_rects_counters"])
@pytest.mark.parametrize("name", ["cfg1", "cfg2_small", "maze64", "ref_dispatch", "chunk5_spp32", "chunk1_spp256", "mirror_limit2", "bounce0", "ragged"])
def test_every_kernel_variant_is_bit_identical(mm, oracle, noise, scenes, renderer, name, flags_name):
    # "regroup" = MM_FLAG_REGROUP: trace_kernel_rg, the default kernel plus warp-level ray compaction at segment boundaries
    flags = {"literal": mm.FLAG_FORCE_LITERAL, "counters_only": mm.FLAG_COUNTERS,
             "literal_counters": mm.FLAG_FORCE_LITERAL | mm.FLAG_COUNTERS, "regroup": mm.FLAG_REGROUP,
             "regroup_literal": mm.FLAG_REGROUP | mm.FLAG_FORCE_LITERAL, "regroup_counters": mm.FLAG_REGROUP | mm.FLAG_COUNTERS,
             # the default leaf test is the collapsed axis-aligned one (every maze rect is axis-aligned); this is the general one
             "general_rects": mm.FLAG_GENERAL_RECTS, "general_rects_counters": mm.FLAG_GENERAL_RECTS | mm.FLAG_COUNTERS}[flags_name]
    sc, u, p, ch = build_case(mm, name, scenes)
    renderer.upload_scene(sc, noise)
    ref = oracle.render(sc, noise, u, p, ch, debug=True)
    p.flags = flags
    got = renderer.render(u, p, ch, debug=True)
    assert_same(got, ref)
    if flags & mm.FLAG_FORCE_LITERAL:
        assert got[1]["literal_rays"] == got[1]["rays"]
    # non-debug kernel variant: image and counters only
    img, cnt, _ = renderer.render(u, p, ch)
    assert img.tobytes() == ref[0].tobytes() and cnt["rays"] == ref[1]["rays"] and cnt["hits"] == ref[1]["hits"]


def test_shared_reciprocal_quotient_is_exact(renderer):
    """The fast slab quotient equals __fdiv_rn on 2^31 guarded operand pairs, half of them next to rounding midpoints."""
    assert renderer.selftest_quotient(1 << 31, seed=20261018) == 0
    assert renderer.selftest_quotient(1 << 28, seed=7) == 0


def test_tile_partition_equals_full_frame(mm, oracle, noise, scenes, renderer):
    """The multi-GPU split (interleaved groups) run as separate calls on one GPU rebuilds the one-call frame."""
    sc, u, p, ch = build_case(mm, "cfg2_small", scenes)
    renderer.upload_scene(sc, noise)
    full, cnt, _ = renderer.render(u, p, ch)
    n_groups = p.grid_x * p.grid_y
    for world in (2, 3, 8):
        r2 = mm.Renderer(0)                      # fresh zero-filled screen
        r2.upload_scene(sc, noise)
        rays = 0
        for rank in range(world):
            q = mm.Params.from_buffer_copy(bytes(p))
            q.group_first, q.group_step, q.group_count = mm.tile_partition(n_groups, rank, world)
            img, c, _ = r2.render(u, q, ch)
            rays += c["rays"]
        assert img.tobytes() == full.tobytes() and rays == cnt["rays"]
        r2.close()


def test_fused_exchange_stores_every_pixel_into_every_frame(mm, noise, scenes, renderer):
    """mm_render_peers_device, as the multi-GPU peer/multicast exchange uses it (ranks emulated on one GPU, the 'peer'
    frames are plain local buffers): after every rank's launch, each frame holds the whole single-call frame."""
    import torch

    sc, u, p, ch = build_case(mm, "yaw", scenes)
    renderer.upload_scene(sc, noise)
    full = renderer.render(u, p, ch)[0]
    dev = torch.device("cuda", 0)
    r2 = mm.Renderer(0)
    r2.upload_scene(sc, noise)
    r2.set_chunks(ch)
    world = 3
    frames = [torch.zeros((int(u.view_height), int(u.view_width), 4), dtype=torch.float32, device=dev) for _ in range(world)]
    for rank in range(world):
        q = mm.Params.from_buffer_copy(bytes(p))
        q.group_first, q.group_step, q.group_count = mm.tile_partition(p.grid_x * p.grid_y, rank, world)
        r2.render_peers_device(u, q, [f.data_ptr() for f in frames])
    r2.sync()
    torch.cuda.synchronize()
    for f in frames:
        assert f.cpu().numpy().tobytes() == full.tobytes()
    with pytest.raises(mm.MMError):
        r2.render_peers_device(u, p, [])
    with pytest.raises(mm.MMError):
        r2.render_peers_device(u, p, [frames[0].data_ptr()] * 9)
    r2.close()


def test_device_tiles_gather_and_scatter(mm, noise, scenes, renderer):
    """mm_render_device -> tiles, then mm_scatter_tiles_device, as the NCCL path uses them (world emulated on one GPU)."""
    import torch

    sc, u, p, ch = build_case(mm, "yaw", scenes)
    renderer.upload_scene(sc, noise)
    full = renderer.render(u, p, ch)[0]
    dev = torch.device("cuda", 0)
    r2 = mm.Renderer(0)
    r2.upload_scene(sc, noise)
    r2.set_chunks(ch)
    world = 4
    parts = [mm.tile_partition(p.grid_x * p.grid_y, r, world) for r in range(world)]
    ppc = u.chunk_width ** 2
    max_count = max(pt[2] for pt in parts)
    gathered = torch.zeros((world, max_count, ppc, 4), dtype=torch.float32, device=dev)
    for rank in range(world):
        q = mm.Params.from_buffer_copy(bytes(p))
        q.group_first, q.group_step, q.group_count = parts[rank]
        r2.render_device(u, q, tiles_ptr=gathered[rank].data_ptr())
    image = torch.zeros((int(u.view_height), int(u.view_width), 4), dtype=torch.float32, device=dev)
    for rank in range(world):
        q = mm.Params.from_buffer_copy(bytes(p))
        q.group_first, q.group_step, q.group_count = parts[rank]
        r2.scatter_tiles_device(u, q, gathered[rank].data_ptr(), image.data_ptr())
    r2.sync()
    torch.cuda.synchronize()
    assert image.cpu().numpy().tobytes() == full.tobytes()
    r2.close()


def test_persistent_screen_texture_semantics(mm, noise, scenes):
    """Like the reference's private screen texture (main.rs:702-709), pixels of chunks that a dispatch does not render
    keep the value of the previous dispatch (progressive refresh, main.rs:778-784)."""
    sc, u, p, ch = build_case(mm, "yaw", scenes)
    r = mm.Renderer(0)
    r.upload_scene(sc, noise)
    q = mm.Params.from_buffer_copy(bytes(p))
    q.group_first, q.group_step, q.group_count = 0, 2, (p.grid_x * p.grid_y + 1) // 2
    half = r.render(u, q, ch)[0].copy()
    assert (half[..., 3] == 0).any() and (half[..., 3] == 1).any()
    u2 = mm.default_uniform(CASES["yaw"]["maze"], u.view_width, u.view_height, 4, time=5)
    q.group_first = 1
    q.group_count = (p.grid_x * p.grid_y) // 2
    both = r.render(u2, q, ch)[0]
    kept = half[..., 3] == 1
    assert (both[..., 3] == 1).all() and both[kept].tobytes() == half[kept].tobytes()
    r.close()


def test_full_size_properties(mm, noise, scenes, renderer):
    """BASELINE configs[1] at full size (32x32 maze, 1080p, 16 spp, 8 bounces): size-independent properties —
    determinism, fast == literal traversal, 2-way tile split == full frame, counters consistent, a crop equal to the
    oracle's render of that crop."""
    from oracle import oracle as o

    sc = scenes(32)
    renderer.upload_scene(sc, noise)
    u = mm.default_uniform(32, 1920, 1080, 4)
    ch = mm.gen_chunks(1920, 1080, 4)
    p = mm.full_frame_params(u, spp=16, bounce_limit=8, flags=mm.FLAG_COUNTERS)
    a, ca, _ = renderer.render(u, p, ch)
    b, cb, _ = renderer.render(u, p, ch)
    assert a.tobytes() == b.tobytes() and ca == cb                               # deterministic
    assert ca["paths"] == 1920 * 1080 * 16 and ca["rays"] >= ca["hits"] and ca["rays"] <= ca["paths"] * (8 + 15)
    assert ca["leaf_visits"] <= ca["rect_tests"] <= 2 * ca["leaf_visits"] and ca["max_stack"] < 14
    assert np.isfinite(a).all() and (a[..., 3] == 1).all() and a[..., :3].min() >= 0
    p.flags = mm.FLAG_COUNTERS | mm.FLAG_FORCE_LITERAL
    c, cc, _ = renderer.render(u, p, ch)
    assert c.tobytes() == a.tobytes()                                             # fast slab == literal divides
    assert all(cc[k] == ca[k] for k in COUNTER_KEYS)
    # crop: 16 chunk columns from the middle of the chunk list, rendered by the oracle with the same group indices
    p.flags = 0
    q = mm.Params.from_buffer_copy(bytes(p))
    q.group_first, q.group_step, q.group_count = 270 * 200, 1, 270 * 4
    crop = np.zeros_like(a)
    o.render(sc, noise, u, q, ch, out=crop)
    m = crop[..., 3] == 1
    assert m.sum() == 270 * 4 * 16 and crop[m].tobytes() == a[m].tobytes()


def test_error_codes(mm, noise, scenes):
    r = mm.Renderer(0)
    sc, u, p, ch = build_case(mm, "bounce1", scenes)
    with pytest.raises(mm.MMError) as e:
        r.render(u, p, ch)
    assert e.value.code == -3                                                    # MM_ERR_NO_SCENE
    r.upload_scene(sc, noise)
    bad = mm.Params.from_buffer_copy(bytes(p)); bad.spp = 3
    with pytest.raises(mm.MMError) as e:
        r.render(u, bad, ch)
    assert e.value.code == -5                                                    # MM_ERR_UNSUPPORTED
    u5 = mm.default_uniform(10, 40, 30, 5)                                       # chunk 5, spp 2: T = 50 is not a (32, h) threadgroup
    with pytest.raises(mm.MMError) as e:
        r.render(u5, mm.full_frame_params(u5, spp=2, bounce_limit=2), mm.gen_chunks(40, 30, 5))
    assert e.value.code == -5
    bad = mm.Params.from_buffer_copy(bytes(p)); bad.grid_x += 1
    with pytest.raises(mm.MMError) as e:
        r.render(u, bad, ch)
    assert e.value.code == -1
    bad = mm.Params.from_buffer_copy(bytes(p)); bad.group_first, bad.group_step, bad.group_count = 5, 7, 10 ** 6
    with pytest.raises(mm.MMError) as e:
        r.render(u, bad, ch)
    assert e.value.code == -1

    class Broken:
        pass

    br = Broken()
    br.planes, br.indices, br.materials, br.emissions = sc.planes, sc.indices, sc.materials, sc.emissions
    br.nodes = sc.nodes.copy()
    br.nodes[0]["left_first"] = len(sc.nodes) + 5                                # child out of range
    with pytest.raises(mm.MMError) as e:
        r.upload_scene(br, noise)
    assert e.value.code == -4                                                    # MM_ERR_BVH
    br.nodes = sc.nodes.copy()
    inner = np.nonzero(br.nodes["tri_count"] == 0)[0]
    br.nodes[inner[-1]]["left_first"] = 0                                        # cycle back to the root
    with pytest.raises(mm.MMError) as e:
        r.upload_scene(br, noise)
    assert e.value.code == -4
    # the context survives errors
    r.upload_scene(sc, noise)
    assert r.render(u, p, ch)[1]["paths"] == 32 * 32 * 8
    r.close()


def test_single_leaf_and_tiny_scenes(mm, oracle, noise, renderer):
    """Root is a leaf (one plane) / two planes: the traversal starts at node 0 without testing its box (shaders.metal:121)."""
    from mirror_maze_b200.host import PLANE_DTYPE, build_bvh

    class S:
        pass

    for n_planes in (1, 2, 3):
        P = np.zeros(n_planes, dtype=PLANE_DTYPE)
        for i in range(n_planes):
            P[i]["origin"], P[i]["v"], P[i]["u"], P[i]["color"] = [-20 + 15 * i, 10, 12 + 4 * i], [12, 0, 0], [0, -20, 0], [0.5, 0.6, 0.7]
        s = S()
        s.planes = P
        s.nodes, s.indices = build_bvh(P)
        s.materials = np.array([i % 2 for i in range(n_planes)], dtype=np.uint8)
        s.emissions = np.tile(np.array([[1.0, 0.5, 0.25, 1.5]], dtype=np.float32), (n_planes, 1))
        u = mm.default_uniform(10, 64, 32, 4, camera_center=(0, 0, 0))
        ch = mm.gen_chunks(64, 32, 4)
        p = mm.full_frame_params(u, spp=8, bounce_limit=3)
        renderer.upload_scene(s, noise)
        assert_same(renderer.render(u, p, ch, debug=True), oracle.render(s, noise, u, p, ch, debug=True))


def test_headless_cpp_driver_matches_python_binding(mm, noise, scenes, tmp_path):
    """mm_headless (C++ consumer of the C-ABI, the stand-in for the Rust driver) renders the same bits as the ctypes path."""
    import os
    import subprocess

    exe = os.path.join(os.path.dirname(mm.library_path()), "mm_headless")
    raw, nz = tmp_path / "frame.f32", tmp_path / "noise.rgba8"
    nz.write_bytes(noise.tobytes())
    out = subprocess.run([exe, "--maze", "16", "--width", "128", "--height", "96", "--spp", "8", "--bounces", "6", "--noise", str(nz),
                          "--raw", str(raw)], capture_output=True, text=True, timeout=120)
    assert out.returncode == 0, out.stderr
    sc = scenes(16)
    r = mm.Renderer(0)
    r.upload_scene(sc, noise)
    u = mm.default_uniform(16, 128, 96, 4)
    img = r.render(u, mm.full_frame_params(u, spp=8, bounce_limit=6), mm.gen_chunks(128, 96, 4))[0]
    assert raw.read_bytes() == img.tobytes()
    r.close()


def test_random_poses_match_oracle(mm, oracle, noise, scenes, renderer):
    """24 seeded camera poses (cell centres, arbitrary yaw, odd `time`) in the 32x32 maze: every observable bit-identical."""
    sc = scenes(32)
    renderer.upload_scene(sc, noise)
    rng = np.random.default_rng(1)
    ch = mm.gen_chunks(48, 32, 4)
    literal_total = 0
    for i in range(24):
        cell = rng.integers(0, 32, size=2)
        center = (-160.0 + 10.0 * cell[0] + 5.0, float(rng.uniform(-6.0, 1.5)), -160.0 + 10.0 * cell[1] + 5.0)
        u = mm.default_uniform(32, 48, 32, 4, time=int(rng.integers(0, 1000)), camera_center=center,
                               half_theta=float(rng.uniform(0.0, np.pi)))
        p = mm.full_frame_params(u, spp=4, bounce_limit=8)
        got = renderer.render(u, p, ch, debug=True)
        assert_same(got, oracle.render(sc, noise, u, p, ch, debug=True))
        literal_total += got[1]["literal_rays"]
    assert literal_total >= 0


def test_literal_lanes_inside_fast_warps(mm, oracle, noise, scenes, renderer):
    """tiny_origin: the camera's |x| < 2^-40 pushes every primary ray onto the literal-divide traversal."""
    sc, u, p, ch = build_case(mm, "tiny_origin", scenes)
    renderer.upload_scene(sc, noise)
    img, cnt, dbg = renderer.render(u, p, ch, debug=True)
    assert cnt["literal_rays"] >= cnt["paths"]            # at least the primary segment of every path
    assert cnt["literal_rays"] < cnt["rays"]              # later segments are back on the fast path
    assert_same((img, cnt, dbg), oracle.render(sc, noise, u, p, ch, debug=True))


@pytest.mark.parametrize("name", REF_SHADER_CASES)
def test_cuda_image_equals_reference_shader_image(mm, noise, scenes, renderer, name):
    """The CUDA kernel against the reference's OWN shader source (compiled as C++ in the authoring container,
    oracle/_ref/libref_shader.so, see oracle/ref_shader/): the fp32 image of every dispatch shape the unmodified shader
    can address, bit for bit — directly when the library travelled to this box, and through its committed digest."""
    from oracle import ref_shader

    sc, u, p, ch = build_case(mm, name, scenes)
    renderer.upload_scene(sc, noise)
    img, cnt, _ = renderer.render(u, p, ch)
    assert hashlib.sha256(img.tobytes()).hexdigest() == GOLDEN[name]["ref_shader_image"]
    if ref_shader.available():
        assert img.tobytes() == ref_shader.render(sc, noise, u, p, ch).tobytes()


def test_unguarded_edge_lengths_take_the_literal_rect_test(mm, oracle, noise, scenes):
    """A scene with an edge longer than 2^23 and one shorter than 2^-20 is outside the guarded range of the divide-free
    edge test: the upload reports fast_rect_ok = 0 and the kernel runs the literal divides — still bit-identical."""
    import types
    from mirror_maze_b200.host import PLANE_DTYPE, build_bvh

    base = scenes(10)
    extra = np.zeros(2, dtype=PLANE_DTYPE)
    extra[0] = ((-8.5e6, 1.9, -8.5e6), (1.7e7, 0.0, 0.0), (0.0, 0.0, 1.7e7), (0.5, 0.5, 0.5))       # giant floor just above the floor
    extra[1] = ((-5.0, 0.0, -40.0), (5e-7, 0.0, 0.0), (0.0, -3.0, 0.0), (0.9, 0.1, 0.1))            # sliver in front of the camera
    planes = np.concatenate([base.planes, extra])
    nodes, indices = build_bvh(planes)
    sc = types.SimpleNamespace(planes=planes, nodes=nodes, indices=indices,
                               materials=np.concatenate([base.materials, np.zeros(2, np.uint8)]),
                               emissions=np.concatenate([base.emissions, np.zeros((2, 4), np.float32)]))
    u = mm.default_uniform(10, 64, 32, 4)
    ch = mm.gen_chunks(64, 32, 4)
    p = mm.full_frame_params(u, spp=8, bounce_limit=5)
    r = mm.Renderer(0)
    r.upload_scene(sc, noise)
    assert r.scene_info()["fast_rect_ok"] == 0 and r.scene_info()["fast_slab_ok"] == 1
    got = r.render(u, p, ch, debug=True)
    assert (got[2]["first_hit"] == len(planes) - 2).any()     # the giant floor is what primary rays see below the horizon
    assert_same(got, oracle.render(sc, noise, u, p, ch, debug=True))
    r.close()
    r = mm.Renderer(0)
    r.upload_scene(base, noise)
    assert r.scene_info()["fast_rect_ok"] == 1
    r.close()


@pytest.mark.parametrize("maze", [32, 64, 256])
def test_full_size_frame_equals_oracle(mm, noise, scenes, maze):
    """BASELINE configs[1] (32x32 maze), the north-star headline (64x64) and configs[3] (256x256) at FULL size — 1920x1080, 16 spp,
    8 bounces, 33.2 M paths, about 265 M rays each — the whole frame and every counter against the CPU oracle (10-20 s of host
    time per maze on the GPU box's cores)."""
    from oracle import oracle as o

    sc = scenes(maze)
    r = mm.Renderer(0)
    r.upload_scene(sc, noise)
    u = mm.default_uniform(maze, 1920, 1080, 4)
    ch = mm.gen_chunks(1920, 1080, 4)
    p = mm.full_frame_params(u, spp=16, bounce_limit=8, flags=mm.FLAG_COUNTERS)
    img, cnt, _ = r.render(u, p, ch)
    ref, rcnt, _ = o.render(sc, noise, u, p, ch)
    assert img.tobytes() == ref.tobytes()
    for k in COUNTER_KEYS:
        assert cnt[k] == rcnt[k], k
    p.flags = 0                                              # the timed (non-counting) kernel variant, zero-copy into pinned memory
    hf = mm.HostFrame(1080, 1920)
    chunks = np.ascontiguousarray(ch)
    c2 = r.render_into(u, p, chunks.ctypes.data, len(chunks), hf.ptr)
    assert hf.array.tobytes() == ref.tobytes() and c2["rays"] == rcnt["rays"]
    hf.close(); r.close()


def test_scatter_gathered_single_launch(mm, noise, scenes, renderer):
    """mm_scatter_gathered_device: the whole all-gather layout (world x max_count tiles, padded) in one launch."""
    import torch

    sc, u, p, ch = build_case(mm, "ragged", scenes)          # 84 groups: not a multiple of 8 -> padded rows
    renderer.upload_scene(sc, noise)
    full = renderer.render(u, p, ch)[0]
    r2 = mm.Renderer(0)
    r2.upload_scene(sc, noise)
    r2.set_chunks(ch)
    for world in (3, 8):
        n_groups = p.grid_x * p.grid_y
        parts = [mm.tile_partition(n_groups, r, world) for r in range(world)]
        max_count = max(pt[2] for pt in parts)
        gathered = torch.full((world * max_count, u.chunk_width ** 2, 4), 7.0, dtype=torch.float32, device="cuda:0")
        torch.cuda.synchronize()
        for rank, (first, step, count) in enumerate(parts):
            q = mm.Params.from_buffer_copy(bytes(p))
            q.group_first, q.group_step, q.group_count = first, step, count
            r2.render_device(u, q, tiles_ptr=gathered[rank * max_count].data_ptr())
        image = torch.zeros((int(u.view_height), int(u.view_width), 4), dtype=torch.float32, device="cuda:0")
        torch.cuda.synchronize()
        r2.scatter_gathered_device(u, p, world, max_count, gathered.data_ptr(), image.data_ptr())
        r2.sync()
        assert image.cpu().numpy().tobytes() == full.tobytes()
    r2.close()


@pytest.mark.parametrize("name", ["cfg1", "cfg2_small", "maze64", "yaw", "on_plane", "tiny_origin", "mirror_limit2", "maze256"])
def test_rcp_slab_variant_matches_oracle(mm, oracle, noise, scenes, renderer, name):
    """MM_FLAG_RCP_SLAB: the opt-in reciprocal-multiply slab arithmetic, bit for bit against the oracle in the same mode."""
    sc, u, p, ch = build_case(mm, name, scenes)
    renderer.upload_scene(sc, noise)
    p.flags = mm.FLAG_RCP_SLAB
    ref = oracle.render(sc, noise, u, p, ch, debug=True)
    assert_same(renderer.render(u, p, ch, debug=True), ref)
    p.flags = mm.FLAG_RCP_SLAB | mm.FLAG_FORCE_LITERAL          # general min/max form for every ray
    assert_same(renderer.render(u, p, ch, debug=True), ref)


def test_roofline_microbenchmarks(renderer):
    """mm_microbench: measured peaks for the node-fetch and FP32-issue rooflines (plausibility only)."""
    assert 500.0 < renderer.microbench(0, 94 * 1024) < 40000.0         # GB/s of useful bytes, L1-resident table
    assert 500.0 < renderer.microbench(0, 6 * 1024 * 1024) < 40000.0   # L2-resident table
    assert 10.0 < renderer.microbench(1) < 40.0                        # T lane-instr/s (148 SMs x 128 lanes x <= 1.965 GHz = 37.2)


def test_multi_gpu_frame_equals_single_gpu_frame_for_every_exchange():
    """tools/mgpu_check.py under torchrun on two GPUs, when the box has them: the N-GPU frame through the NCCL gather,
    the fused NVLink peer stores and the fused NVSwitch multicast stores, each bit-identical to the one-GPU frame."""
    import socket
    import subprocess
    import sys
    import torch

    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    root = os.path.dirname(HERE)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", str(port), os.path.join(root, "tools", "mgpu_check.py")]
    out = subprocess.run(cmd, cwd=root, capture_output=True, text=True, timeout=600)
    assert out.returncode == 0 and "MULTI-GPU PARITY OK" in out.stdout, out.stdout[-2000:] + out.stderr[-2000:]


# ---- full-size frame coordinates of the other BASELINE configurations (crops checked against the oracle) ----------------

FULL_SIZE = {
    # north-star headline: 64x64 maze, 1080p x 16 spp x 8 bounces
    "hl64": dict(maze=64, W=1920, H=1080, spp=16, bounce=8),
    # BASELINE configs[2]: 64x64 maze, 3840x2160, 64 spp (T = 1024 threads per virtual group), 16 bounces
    "cfg3": dict(maze=64, W=3840, H=2160, spp=64, bounce=16),
    # BASELINE configs[3]: 256x256 maze (43.6 k planes, BVH depth 21), 1080p x 16 spp x 8 bounces
    "cfg4": dict(maze=256, W=1920, H=1080, spp=16, bounce=8),
}


@pytest.mark.parametrize("name", sorted(FULL_SIZE))
def test_full_size_crops_match_oracle(mm, oracle, noise, scenes, name):
    """The real frame coordinates of the 64x64 headline, config 3 and config 4: crops of chunk groups at the first, a middle
    and the last chunk columns of the full-size virtual grid (so the largest texid values and the u32 wraps of
    texid * 15823 / texid * 9737333 in the seed, shaders.metal:298, are the real ones), every observable bit-identical."""
    c = FULL_SIZE[name]
    sc = scenes(c["maze"])
    u = mm.default_uniform(c["maze"], c["W"], c["H"], 4)
    ch = mm.gen_chunks(c["W"], c["H"], 4)
    p = mm.full_frame_params(u, spp=c["spp"], bounce_limit=c["bounce"])
    n_groups = p.grid_x * p.grid_y
    col = c["H"] // 4                                            # groups per chunk column (gen_pixels order: x outer, y inner)
    count = 96 if c["spp"] >= 64 else 270
    r = mm.Renderer(0)                                           # fresh zero-filled screen
    r.upload_scene(sc, noise)
    for first in (0, (n_groups // 2 // col) * col + col // 3, n_groups - count):
        q = mm.Params.from_buffer_copy(bytes(p))
        q.group_first, q.group_step, q.group_count = first, 1, count
        ref_img = np.zeros((c["H"], c["W"], 4), dtype=np.float32)
        ref = oracle.render(sc, noise, u, q, ch, debug=True, out=ref_img)
        got = r.render(u, q, ch, debug=True)
        m = ref[0][..., 3] == 1
        assert m.sum() == count * 16
        assert got[0][m].tobytes() == ref[0][m].tobytes(), f"{name}: crop at group {first}"
        assert_same((ref[0], got[1], got[2]), ref)
    # a strided sample over the whole grid as well (every 997th group: all chunk rows and columns)
    q = mm.Params.from_buffer_copy(bytes(p))
    q.group_first, q.group_step = 5, 997
    q.group_count = (n_groups - 5 + 996) // 997
    if c["spp"] >= 64:
        q.group_step = 4999; q.group_count = (n_groups - 5 + 4998) // 4999
    ref = oracle.render(sc, noise, u, q, ch, debug=True)
    got = r.render(u, q, ch, debug=True)
    m = ref[0][..., 3] == 1
    assert got[0][m].tobytes() == ref[0][m].tobytes()
    assert_same((ref[0], got[1], got[2]), ref)
    r.close()


def test_zero_copy_async_and_staged_host_frames_are_identical(mm, noise, scenes):
    """mm_render into (a) a pageable numpy array (staged copy), (b) a mapped pinned HostFrame (the kernel's zero-copy stores),
    (c) the same with MM_FLAG_NO_ZERO_COPY (DMA copy), (d) mm_render_async + mm_wait with a kept chunk list (chunks = NULL),
    (e) a caller-registered buffer (mm_host_register): all the same bits."""
    import ctypes as C

    sc, u, p, ch = build_case(mm, "cfg2_small", scenes)
    r = mm.Renderer(0)
    r.upload_scene(sc, noise)
    H, W = int(u.view_height), int(u.view_width)
    a, ca, _ = r.render(u, p, ch)                                       # (a)
    hf = mm.HostFrame(H, W)
    chunks = np.ascontiguousarray(ch)
    cb = r.render_into(u, p, chunks.ctypes.data, len(chunks), hf.ptr)   # (b)
    assert hf.array.tobytes() == a.tobytes() and cb == ca
    hf.array[...] = -1.0
    q = mm.Params.from_buffer_copy(bytes(p)); q.flags = mm.FLAG_NO_ZERO_COPY
    r.render_into(u, q, chunks.ctypes.data, len(chunks), hf.ptr)        # (c)
    assert hf.array.tobytes() == a.tobytes()
    hf.array[...] = -1.0
    r.render_async(u, p, None, 0, hf.ptr)                               # (d) chunk list kept from the previous call
    cd = r.wait()
    assert hf.array.tobytes() == a.tobytes() and cd == ca
    mine = np.full((H, W, 4), -1.0, dtype=np.float32)                   # (e)
    lib = mm.load_library()
    assert lib.mm_host_register(mine.ctypes.data, mine.nbytes) == 0
    r.render_into(u, p, None, 0, mine.ctypes.data)
    assert lib.mm_host_unregister(mine.ctypes.data) == 0
    assert mine.tobytes() == a.tobytes()
    # zero-copy writes only the chunks a dispatch renders: the buffer is the caller's persistent copy of the screen
    hf.array[...] = 0.0
    half = mm.Params.from_buffer_copy(bytes(p))
    half.group_first, half.group_step, half.group_count = 0, 2, (p.grid_x * p.grid_y + 1) // 2
    r2 = mm.Renderer(0)
    r2.upload_scene(sc, noise)
    r2.render_into(u, half, chunks.ctypes.data, len(chunks), hf.ptr)
    half.group_first, half.group_count = 1, (p.grid_x * p.grid_y) // 2
    r2.render_into(u, half, None, 0, hf.ptr)
    assert hf.array.tobytes() == a.tobytes()
    r2.close(); r.close(); hf.close()


def test_deep_and_padded_bvh(mm, oracle, noise):
    """A skewed BVH of depth 51 (stack occupancy 50: all the reference's stack holds, shaders.metal:123) renders like the
    oracle; depth 52 is refused with MM_ERR_BVH; a node array passed at its 2n-1 capacity with an unreachable garbage tail
    uploads and renders the same frame."""
    import types
    from mirror_maze_b200.host import PLANE_DTYPE, NODE_DTYPE

    def chain(n_planes):
        # planes side by side along x; node 2k+1 = leaf k, node 2k+2 = the rest: depth n_planes
        P = np.zeros(n_planes, dtype=PLANE_DTYPE)
        for i in range(n_planes):
            P[i]["origin"], P[i]["v"], P[i]["u"], P[i]["color"] = [-60.0 + 2.5 * i, 6.0, 30.0], [2.0, 0.0, 0.0], [0.0, -12.0, 0.0], [0.6, 0.5, 0.4]
        lo = np.array([[P[i]["origin"][0], -6.0, 30.0] for i in range(n_planes)], dtype=np.float32)
        hi = np.array([[P[i]["origin"][0] + 2.0, 6.0, 30.0] for i in range(n_planes)], dtype=np.float32)
        nodes = np.zeros(2 * n_planes - 1, dtype=NODE_DTYPE)
        def box(i, a, b):
            nodes[i]["aabb_min"], nodes[i]["aabb_max"] = lo[a:b].min(axis=0), hi[a:b].max(axis=0)
        at, k = 0, 0
        while True:
            box(at, k, n_planes)
            if n_planes - k == 1:
                nodes[at]["left_first"], nodes[at]["tri_count"] = k, 1
                break
            nodes[at]["left_first"], nodes[at]["tri_count"] = 2 * k + 1, 0
            box(2 * k + 1, k, k + 1)
            nodes[2 * k + 1]["left_first"], nodes[2 * k + 1]["tri_count"] = k, 1
            at, k = 2 * k + 2, k + 1
        return types.SimpleNamespace(planes=P, nodes=nodes, indices=np.arange(n_planes, dtype=np.uint32),
                                     materials=(np.arange(n_planes) % 3 == 0).astype(np.uint8),
                                     emissions=np.tile(np.array([[1.0, 0.8, 0.3, 1.0]], dtype=np.float32), (n_planes, 1)))

    u = mm.default_uniform(10, 64, 32, 4, camera_center=(0.0, 0.0, 0.0))
    ch = mm.gen_chunks(64, 32, 4)
    p = mm.full_frame_params(u, spp=8, bounce_limit=4)
    r = mm.Renderer(0)
    s51 = chain(51)
    r.upload_scene(s51, noise)
    assert r.scene_info()["bvh_depth"] == 51 == mm.MAX_BVH_DEPTH
    ref = oracle.render(s51, noise, u, p, ch, debug=True)
    assert_same(r.render(u, p, ch, debug=True), ref)
    with pytest.raises(mm.MMError) as e:
        r.upload_scene(chain(52), noise)
    assert e.value.code == -4
    padded = chain(51)
    tail = np.zeros(40, dtype=NODE_DTYPE)
    tail["left_first"] = 0xFFFFFF00                                  # garbage interior nodes nobody reaches
    padded.nodes = np.concatenate([padded.nodes, tail])
    r.upload_scene(padded, noise)
    assert r.render(u, p, ch)[0].tobytes() == ref[0].tobytes()
    r.close()


def _multi_frames(mm, noise, sc, u, p, ch, devices, exchange):
    H, W = int(u.view_height), int(u.view_width)
    m = mm.MultiRenderer(devices, exchange)
    m.upload_scene(sc, noise)
    hf = mm.HostFrame(H, W)
    cnt = m.render(u, p, ch, hf)                                  # zero-copy assembly in pinned host memory
    pinned = hf.array.copy()
    paged = np.zeros((H, W, 4), dtype=np.float32)
    cnt2 = None
    if exchange != "none" or len(devices) == 1:
        cnt2 = m.render(u, p, None, paged)                        # staged copy of device 0's assembled frame
    ms = m.last_ms()
    m.close(); hf.close()
    return pinned, paged, cnt, cnt2, ms


def test_multi_context_on_one_device_equals_single_context(mm, noise, scenes, renderer):
    """mm_multi over a one-device list: the same frame and counters as mm_render."""
    sc, u, p, ch = build_case(mm, "ragged", scenes)
    renderer.upload_scene(sc, noise)
    q = mm.Params.from_buffer_copy(bytes(p)); q.flags = mm.FLAG_COUNTERS
    r = mm.Renderer(0); r.upload_scene(sc, noise)
    full, cnt, _ = r.render(u, q, ch)
    r.close()
    pinned, paged, c1, c2, ms = _multi_frames(mm, noise, sc, u, q, ch, [0], "peer")
    assert pinned.tobytes() == full.tobytes() and paged.tobytes() == full.tobytes()
    assert c1 == cnt and c2 == cnt and ms > 0


@pytest.mark.parametrize("exchange", ["peer", "nccl", "none"])
def test_multi_gpu_one_process_frame_equals_single_gpu_frame(mm, noise, scenes, exchange):
    """mm_multi (one process, no torch in the data path): the frame split over every visible GPU (2..8) is bit-identical to
    the one-GPU frame, for the fused peer-store exchange, the NCCL tile gather and host-only assembly; counters add up."""
    import torch

    n = min(torch.cuda.device_count(), 8)
    if n < 2:
        pytest.skip("needs two GPUs")
    sc, u, p, ch = build_case(mm, "cfg2_small", scenes)
    q = mm.Params.from_buffer_copy(bytes(p)); q.flags = mm.FLAG_COUNTERS
    r = mm.Renderer(0); r.upload_scene(sc, noise)
    full, cnt, _ = r.render(u, q, ch)
    r.close()
    for devs in ([0, 1], list(range(n))):
        pinned, paged, c1, c2, ms = _multi_frames(mm, noise, sc, u, q, ch, devs, exchange)
        assert pinned.tobytes() == full.tobytes(), (exchange, devs)
        assert c1 == cnt
        if c2 is not None:
            assert paged.tobytes() == full.tobytes() and c2 == cnt
        # device-resident frames: every device holds the whole frame after a peer / nccl exchange
        if exchange != "none":
            m = mm.MultiRenderer(devs, exchange); m.upload_scene(sc, noise)
            hf = mm.HostFrame(int(u.view_height), int(u.view_width))
            m.render(u, q, ch, hf)
            import ctypes as C
            cudart = C.CDLL("libcudart.so.12")                          # the runtime torch already loaded
            cudart.cudaMemcpy.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_int]
            for i, d in enumerate(devs):
                t = torch.empty((int(u.view_height), int(u.view_width), 4), dtype=torch.float32, device=f"cuda:{d}")
                torch.cuda.synchronize(d)
                assert cudart.cudaMemcpy(t.data_ptr(), m.frame_device_ptr(i), t.numel() * 4, 4) == 0          # cudaMemcpyDefault
                assert t.cpu().numpy().tobytes() == full.tobytes(), (exchange, d)
            m.close(); hf.close()


@pytest.mark.parametrize("name", ["cfg1", "cfg2_small", "cfg3_small", "maze64", "maze256", "ref_dispatch", "yaw", "tiny_origin", "on_plane",
                                  "mirror_limit2", "bounce0", "bounce1", "all_miss", "ragged", "chunk2_spp4", "chunk5_spp32", "chunk16_spp1",
                                  "chunk8_spp2", "chunk3_spp32"])
def test_pool_kernel_matches_oracle(mm, oracle, noise, scenes, renderer, name):
    """MM_FLAG_POOL_KERNEL: the persistent ray-pool kernel (pool_kernel.cu) — warps that own a pool of paths in shared memory
    and run generate / interior / leaf / shade bodies from work queues — gives every observable and counter of the oracle,
    including rays that take the literal-divide traversal (tiny_origin), paths that end at once (bounce0), 1 to 64 samples per
    pixel, ragged frames and thread groups that straddle warps (chunk5_spp32)."""
    sc, u, p, ch = build_case(mm, name, scenes)
    renderer.upload_scene(sc, noise)
    ref = oracle.render(sc, noise, u, p, ch, debug=True)
    p.flags = mm.FLAG_POOL_KERNEL
    assert_same(renderer.render(u, p, ch, debug=True), ref)
    p.flags = mm.FLAG_POOL_KERNEL | mm.FLAG_COUNTERS
    img, cnt, _ = renderer.render(u, p, ch)
    assert img.tobytes() == ref[0].tobytes()
    for k in COUNTER_KEYS:
        assert cnt[k] == ref[1][k], k
    p.flags = mm.FLAG_POOL_KERNEL
    img, cnt, _ = renderer.render(u, p, ch)
    assert img.tobytes() == ref[0].tobytes() and cnt["rays"] == ref[1]["rays"] and cnt["hits"] == ref[1]["hits"]
