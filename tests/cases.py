"""Parity cases shared by the oracle tests (CPU) and the CUDA parity tests (GPU).  Each case is a full set of inputs
for one dispatch of the hot path; sizes are chosen so that the CPU oracle finishes in seconds."""
import numpy as np

CASES = {
    # BASELINE.json configs[0]: 16x16 maze, 256x256, 1 spp, 4 bounces
    "cfg1": dict(maze=16, W=256, H=256, chunk=4, spp=1, bounce=4, mirror=15),
    # the reference's own dispatch shape: 32x24 groups of 1024 threads = 768 chunks x 16 px x 64 spp, limits 5/15
    # (main.rs:599-602,641-650; shaders.metal:294-295), chunk list = a bag of 768 shuffled chunk origins, time = 3
    "ref_dispatch": dict(maze=10, W=1024, H=768, chunk=4, spp=64, bounce=5, mirror=15, time=3, grid=(32, 24), bag_seed=11),
    # BASELINE.json configs[1] at 1/10 linear scale: 32x32 maze, 16 spp, 8 bounces
    "cfg2_small": dict(maze=32, W=192, H=108, chunk=4, spp=16, bounce=8, mirror=15),
    # north-star 64x64 maze at small frame
    "maze64": dict(maze=64, W=160, H=88, chunk=4, spp=16, bounce=8, mirror=15),
    # BASELINE.json configs[2] scaled down: 64x64 maze, 64 spp (T = 1024 threads per virtual group), 16 bounces
    "cfg3_small": dict(maze=64, W=128, H=72, chunk=4, spp=64, bounce=16, mirror=15),
    # BASELINE.json configs[3] scaled down: 256x256 maze (BVH depth 21, 43.6 k planes), 8 spp here
    "maze256": dict(maze=256, W=96, H=64, chunk=4, spp=8, bounce=8, mirror=15),
    # moved and turned camera (mouse yaw, main.rs:923-929), odd time
    "yaw": dict(maze=16, W=128, H=96, chunk=4, spp=8, bounce=8, mirror=15, time=77, center=(25.0, 0.0, -15.0), half_theta=1.1),
    # other chunk widths / spp below 8 (SURVEY §8 D12 generalisation)
    "chunk2_spp4": dict(maze=16, W=64, H=48, chunk=2, spp=4, bounce=6, mirror=15),
    "chunk8_spp2": dict(maze=10, W=64, H=64, chunk=8, spp=2, bounce=5, mirror=15),
    "chunk3_spp32": dict(maze=10, W=48, H=48, chunk=3, spp=32, bounce=5, mirror=15),
    # odd virtual-dispatch shapes: 1 pixel x 256 samples per group; 25 pixels x 32 samples (T = 800, groups straddle
    # blocks); 256 pixels x 1 sample
    "chunk1_spp256": dict(maze=10, W=12, H=8, chunk=1, spp=256, bounce=5, mirror=15),
    "chunk5_spp32": dict(maze=10, W=20, H=15, chunk=5, spp=32, bounce=5, mirror=15),
    "chunk16_spp1": dict(maze=16, W=64, H=48, chunk=16, spp=1, bounce=6, mirror=15),
    # mirror_limit reached: break inside the mirror branch (shaders.metal:331-334)
    "mirror_limit2": dict(maze=32, W=96, H=64, chunk=4, spp=8, bounce=8, mirror=2, center=(-5.0, 0.0, 35.0), half_theta=2.4),
    # no bounces at all / one bounce
    "bounce0": dict(maze=10, W=32, H=32, chunk=4, spp=8, bounce=0, mirror=15),
    "bounce1": dict(maze=10, W=32, H=32, chunk=4, spp=8, bounce=1, mirror=15),
    # camera outside the box looking away: every path misses on the first segment
    "all_miss": dict(maze=10, W=32, H=32, chunk=4, spp=8, bounce=4, mirror=15, center=(0.0, 0.0, -500.0), half_theta=3.14159),
    # camera x within 2^-40 of 0 but not 0: outside the guarded range of the shared-reciprocal slab test, so every
    # primary ray takes the literal-divide traversal while later segments take the fast one (mixed warps)
    "tiny_origin": dict(maze=10, W=64, H=32, chunk=4, spp=8, bounce=4, mirror=15, center=(1e-20, 0.0, -45.0)),
    # camera exactly on a wall plane coordinate (x = 0): quotients that are exactly 0 and 0/0-free ties
    "on_plane": dict(maze=16, W=64, H=32, chunk=4, spp=8, bounce=5, mirror=15, center=(0.0, 0.0, -75.0), half_theta=0.3),
    # ragged: frame not a multiple of the chunk => gen_chunks floors (main.rs:294-295) and border pixels stay unwritten
    "ragged": dict(maze=10, W=50, H=30, chunk=4, spp=8, bounce=3, mirror=15),
    # Dispatches the reference's UNMODIFIED shader can address (oracle/ref_shader.py): its literal limits 5 / 15, at least
    # 8 samples, at most 1024 threads per group and grid_x == (W / 2) / chunk^2 (shaders.metal:266,294-295,343-358).
    # ref_dispatch above is the reference's own 1024 x 768 dispatch; these vary maze, pose, chunk, spp and time.
    "refsh_maze16_yaw": dict(maze=16, W=256, H=128, chunk=4, spp=16, bounce=5, mirror=15, time=9, half_theta=0.7,
                             grid=(8, 6), bag_seed=5),
    "refsh_chunk2_spp64": dict(maze=10, W=128, H=64, chunk=2, spp=64, bounce=5, mirror=15, time=1, grid=(16, 10), bag_seed=6),
    "refsh_chunk8_spp16": dict(maze=32, W=512, H=256, chunk=8, spp=16, bounce=5, mirror=15, time=2, half_theta=2.2,
                               center=(15.0, 0.0, 25.0), grid=(4, 5), bag_seed=7),
    "refsh_chunk1_spp256": dict(maze=10, W=64, H=64, chunk=1, spp=256, bounce=5, mirror=15, time=77, grid=(32, 8), bag_seed=8),
    "refsh_spp8": dict(maze=10, W=256, H=64, chunk=4, spp=8, bounce=5, mirror=15, time=4, half_theta=1.3, grid=(8, 4), bag_seed=9),
}

# cases the compiled reference shader can run (the fixtures hold its image digests, see tests/golden/make_golden.py)
REF_SHADER_CASES = ["ref_dispatch", "refsh_maze16_yaw", "refsh_chunk2_spp64", "refsh_chunk8_spp16", "refsh_chunk1_spp256", "refsh_spp8"]

# cases small enough for the numpy transcription
NP_CASES = ["cfg1", "yaw", "tiny_origin", "on_plane", "chunk1_spp256", "chunk5_spp32", "chunk16_spp1", "chunk2_spp4", "chunk8_spp2", "chunk3_spp32", "mirror_limit2", "bounce0", "bounce1", "all_miss", "ragged"]


def build_case(mm, name, scenes=None):
    c = CASES[name]
    sc = scenes(c["maze"]) if scenes else mm.MazeScene(c["maze"], 0)
    u = mm.default_uniform(c["maze"], c["W"], c["H"], c["chunk"], c.get("time", 0), camera_center=c.get("center"),
                           half_theta=c.get("half_theta"))
    chunks = mm.gen_chunks(c["W"], c["H"], c["chunk"])
    p = mm.full_frame_params(u, spp=c["spp"], bounce_limit=c["bounce"], mirror_limit=c["mirror"])
    if "grid" in c:
        gx, gy = c["grid"]
        rng = np.random.default_rng(c["bag_seed"])          # stands in for thread_rng (main.rs:303-305)
        chunks = chunks[rng.permutation(len(chunks))[: gx * gy]].copy()
        p.grid_x, p.grid_y = gx, gy
    return sc, u, p, chunks
