"""BASELINE.json configs[4]: camera fly-throughs with temporal sample accumulation, frames batched across GPUs.

One fly-through = the reference's frame loop (reference src/main.rs:767-895) run headless for N frames: scripted WASD +
yaw input -> mm_move_camera (collision) / mm_update_quat_angle -> pop a bag of chunk origins (progressive refresh,
main.rs:778-784) -> mm_render into the persistent screen -> mm_present (5-tap blur).  Under torchrun every rank runs its
own fly-through (no exchange until the end).  Prints one JSON line per rank-0 with frames/s and Mrays/s over all ranks.
"""
import argparse, json, math, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import mirror_maze_b200 as mm


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--maze", type=int, default=64)
    ap.add_argument("--frames", type=int, default=120)
    ap.add_argument("--width", type=int, default=1920)
    ap.add_argument("--height", type=int, default=1080)
    ap.add_argument("--spp", type=int, default=16)
    ap.add_argument("--bounces", type=int, default=8)
    ap.add_argument("--refresh", type=int, default=64, help="1/refresh of the screen is re-rendered per frame (reference: 64)")
    a = ap.parse_args()
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    dist = None
    if world > 1:
        import torch, torch.distributed as dist
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    noise = mm.load_noise()
    sc = mm.MazeScene(a.maze, 0)
    r = mm.Renderer(local)
    r.upload_scene(sc, noise)
    u = mm.default_uniform(a.maze, a.width, a.height, 4)
    n_chunks = (a.width // 4) * (a.height // 4)
    per_frame = max(1, n_chunks // a.refresh)
    gx = max(1, int(math.sqrt(per_frame * a.width / a.height)))
    while per_frame % gx:
        gx -= 1
    p = mm.Params(spp=a.spp, bounce_limit=a.bounces, mirror_limit=15, grid_x=gx, grid_y=per_frame // gx)
    bag = mm.ChunkBag(a.width, a.height, 4, seed=1000 + rank)
    q = np.array([u.cam.rotation.x, u.cam.rotation.y, u.cam.rotation.z, u.cam.rotation.w], dtype=np.float32)
    half_theta = math.acos(float(q[3]))
    center = np.array([u.cam.camera_center.x, u.cam.camera_center.y, u.cam.camera_center.z], dtype=np.float32)
    rng = np.random.default_rng(rank)
    import torch
    host = torch.zeros((a.height, a.width, 4), dtype=torch.float32).pin_memory()     # pinned frame buffer for the read-back
    out = host.numpy()
    rays = 0
    blocked = 0
    t0 = time.perf_counter()
    for frame in range(a.frames):
        center, b = mm.move_camera(sc.nodes, center, q, [13], fps=60.0)          # hold W
        blocked += b
        if b or frame % 30 == 29:                                                   # turn when blocked, and now and then
            half_theta = (half_theta - float(rng.uniform(-0.6, 0.6))) % math.pi     # main.rs:923-924 rem_euclid(PI)
            nq = mm.update_quat_angle(q, half_theta)
            if not np.isnan(nq).any():                                              # main.rs:830-841
                q = nq
                bag.reshuffle()
        u.cam.camera_center = mm.Float3(*[float(v) for v in center])
        u.cam.rotation = mm.Float4(*[float(v) for v in q])
        u.time = frame
        ch = bag.next(per_frame)
        cnt = r.render_into(u, p, ch.ctypes.data, len(ch), None)      # compute pass into the persistent screen, no read-back
        r.present(out)                                                # present pass (blur) + one read-back of the frame
        rays += cnt["rays"]
    dt = time.perf_counter() - t0
    tot = np.array([rays, a.frames, dt], dtype=np.float64)
    if dist is not None:
        import torch
        t = torch.tensor(tot, device=f"cuda:{local}")
        mx = t.clone()
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        dist.all_reduce(mx, op=dist.ReduceOp.MAX)
        tot = np.array([t[0].item(), t[1].item(), mx[2].item()])
        dist.destroy_process_group()
    if rank == 0:
        print(json.dumps({"workload": f"{world} fly-through(s) x {a.frames} frames, {a.maze}x{a.maze} maze, {a.width}x{a.height}, {a.spp} spp, "
                                      f"{a.bounces} bounces, 1/{a.refresh} of the screen per frame + 5-tap blur, frame read back to pinned host memory every frame",
                          "frames_per_s": round(tot[1] / tot[2], 2), "Mrays_per_s": round(tot[0] / tot[2] / 1e6, 1), "seconds": round(tot[2], 3),
                          "n_gpus": world, "chunks_per_frame": per_frame, "blocked_moves_rank0": int(blocked), "frame_mean": float(out[..., :3].mean())}))


if __name__ == "__main__":
    main()
