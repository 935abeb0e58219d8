"""Developer tool: times the north-star frame with the library named by MM_LIBRARY (A/B of build variants)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import mirror_maze_b200 as mm
noise = mm.load_noise()
r = mm.Renderer(0)
for n in [int(x) for x in (sys.argv[1] if len(sys.argv) > 1 else "32,64").split(",")]:
    sc = mm.MazeScene(n, 0)
    r.upload_scene(sc, noise)
    u = mm.default_uniform(n, 1920, 1080, 4)
    ch = mm.gen_chunks(1920, 1080, 4)
    for flags in (0, mm.FLAG_FORCE_SHARED):
        if n >= 64 and flags: continue
        p = mm.full_frame_params(u, spp=16, bounce_limit=8, flags=flags)
        best = 1e9
        for it in range(4):
            img, cnt, _ = r.render(u, p, ch)
            best = min(best, r.last_ms())
        info = r.scene_info()
        print(f"{os.environ.get('MM_LIBRARY','default'):28s} N={n} flags={flags} {best:7.2f} ms {cnt['rays']/best/1e3:8.1f} Mrays/s smem={info['nodes_in_shared']} blk/SM={info['blocks_per_sm']} thr={info['block_threads']}", flush=True)
