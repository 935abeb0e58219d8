"""Developer tool: times the ray-pool kernel on the metric's frame over a grid of its tuning overrides (MM_POOL_* environment
variables, read at every launch).  Run under gpurun.  usage: pool_tune.py [maze] [key=v1,v2 ...]"""
import itertools, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import mirror_maze_b200 as mm

maze = int(sys.argv[1]) if len(sys.argv) > 1 and sys.argv[1].isdigit() else 32
grid = {}
for a in sys.argv[1:]:
    if "=" in a:
        k, v = a.split("=")
        grid[k] = v.split(",")
noise = mm.load_noise()
sc = mm.MazeScene(maze, 0)
r = mm.Renderer(0)
r.upload_scene(sc, noise)
u = mm.default_uniform(maze, 1920, 1080, 4)
ch = mm.gen_chunks(1920, 1080, 4)
hf = mm.HostFrame(1080, 1920)
import numpy as np
chunks = np.ascontiguousarray(ch)
keys = sorted(grid)
ref = None
for combo in itertools.product(*[grid[k] for k in keys]):
    for k, v in zip(keys, combo):
        os.environ[k] = v
    flags = int(os.environ.get("FLAGS", str(mm.FLAG_POOL_KERNEL)))
    p = mm.full_frame_params(u, spp=16, bounce_limit=8, flags=flags)
    best = 1e9
    for it in range(3):
        cnt = r.render_into(u, p, chunks.ctypes.data, len(chunks), hf.ptr)
        best = min(best, r.last_ms())
    import hashlib
    h = hashlib.sha256(hf.array.tobytes()).hexdigest()[:12]
    if ref is None:
        ref = h
    print(dict(zip(keys, combo)), f"{best:.2f} ms", r.scene_info()["blocks_per_sm"], "blocks/SM", "OK" if h == ref else "IMAGE DIFFERS", flush=True)
