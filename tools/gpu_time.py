"""Developer tool: times the north-star frame (optionally with the library named by MM_LIBRARY, for A/B of build variants)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import mirror_maze_b200 as mm
noise = mm.load_noise()
r = mm.Renderer(0)
modes = [("exact", 0), ("rcp", mm.FLAG_RCP_SLAB)] if "--rcp" in sys.argv else [("exact", 0)]
mazes = [a for a in sys.argv[1:] if not a.startswith("--")]
for n in [int(x) for x in (mazes[0] if mazes else "32,64").split(",")]:
    sc = mm.MazeScene(n, 0)
    r.upload_scene(sc, noise)
    u = mm.default_uniform(n, 1920, 1080, 4)
    ch = mm.gen_chunks(1920, 1080, 4)
    for name, flags in modes:
        p = mm.full_frame_params(u, spp=16, bounce_limit=8, flags=flags)
        best = 1e9
        for it in range(4):
            img, cnt, _ = r.render(u, p, ch)
            best = min(best, r.last_ms())
        info = r.scene_info()
        print(f"{os.environ.get('MM_LIBRARY','default'):24s} N={n} {name:6s} {best:7.2f} ms {cnt['rays']/best/1e3:8.1f} Mrays/s rays={cnt['rays']} literal={cnt['literal_rays']} blk/SM={info['blocks_per_sm']}", flush=True)
