"""Developer tool: sweeps the warp-scheduler parameters (MM_SCHED=th,wI,wL,wS) on the north-star frame."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import mirror_maze_b200 as mm

def main():
    noise = mm.load_noise()
    r = mm.Renderer(0)
    mazes = [int(x) for x in (sys.argv[1] if len(sys.argv) > 1 else "32,64").split(",")]
    scheds = sys.argv[2].split(";") if len(sys.argv) > 2 else ["16,1,1,1", "33,1,1,1", "12,1,1,1", "20,1,1,1", "24,1,1,1", "16,1,2,2", "16,2,1,1", "16,1,2,4", "16,1,1,2", "0,1,1,1"]
    for n in mazes:
        sc = mm.MazeScene(n, 0)
        r.upload_scene(sc, noise)
        u = mm.default_uniform(n, 1920, 1080, 4)
        ch = mm.gen_chunks(1920, 1080, 4)
        for flags in (0, mm.FLAG_FORCE_GLOBAL):
            if n >= 64 and flags == 0 and not os.environ.get("MM_SMEM_NODE_LIMIT"):
                continue
            for sched in scheds:
                os.environ["MM_SCHED"] = sched
                p = mm.full_frame_params(u, spp=16, bounce_limit=8, flags=flags)
                best = 1e9
                for it in range(3):
                    img, cnt, _ = r.render(u, p, ch)
                    best = min(best, r.last_ms())
                info = r.scene_info()
                print(f"N={n} flags={flags} sched={sched:12s} {best:7.2f} ms {cnt['rays']/best/1e3:8.1f} Mrays/s smem={info['nodes_in_shared']} blk/SM={info['blocks_per_sm']}", flush=True)

if __name__ == "__main__":
    main()
