"""Developer tool (GPU box): many seeded random camera poses, GPU against the CPU oracle, every observable bit for bit."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import mirror_maze_b200 as mm
from oracle import oracle

def main():
    n_poses = int(sys.argv[1]) if len(sys.argv) > 1 else 200
    noise = mm.load_noise()
    r = mm.Renderer(0)
    rng = np.random.default_rng(2026)
    bad = rays = lit = 0
    t0 = time.time()
    for maze in (10, 32, 64, 128):
        sc = mm.MazeScene(maze, 0)
        r.upload_scene(sc, noise)
        ch = mm.gen_chunks(96, 64, 4)
        for i in range(n_poses // 4):
            cell = rng.integers(0, maze, size=2)
            half = 5.0 * maze
            center = (-half + 10.0 * cell[0] + float(rng.uniform(0.5, 9.5)), float(rng.uniform(-7.5, 1.9)), -half + 10.0 * cell[1] + float(rng.uniform(0.5, 9.5)))
            if i % 7 == 0:
                center = (-half + 10.0 * cell[0], center[1], center[2])            # exactly on a wall-plane coordinate
            u = mm.default_uniform(maze, 96, 64, 4, time=int(rng.integers(0, 100000)), camera_center=center, half_theta=float(rng.uniform(0.0, np.pi)))
            kernel = (0, mm.FLAG_POOL_KERNEL, mm.FLAG_REGROUP)[i % 3]               # the shipped kernel and the two opt-in scheduling kernels
            p = mm.full_frame_params(u, spp=int(rng.choice([1, 4, 8, 16])), bounce_limit=int(rng.integers(1, 12)), mirror_limit=int(rng.choice([2, 15])), flags=mm.FLAG_COUNTERS | kernel)
            img, cnt, dbg = r.render(u, p, ch, debug=True)
            rimg, rcnt, rdbg = oracle.render(sc, noise, u, p, ch, debug=True)
            ok = img.tobytes() == rimg.tobytes() and all(dbg[k].tobytes() == rdbg[k].tobytes() for k in dbg) and \
                all(cnt[k] == rcnt[k] for k in ("paths", "rays", "inner_visits", "leaf_visits", "rect_tests", "hits", "max_stack"))
            bad += 0 if ok else 1
            rays += cnt["rays"]; lit += cnt["literal_rays"]
            if not ok:
                print("MISMATCH maze", maze, "pose", i, center)
    print(f"stress parity: {n_poses} poses over mazes 10/32/64/128 (kernels: shipped / pool / regroup in turn), {rays} rays ({lit} on the literal-divide path), mismatches: {bad}, {time.time() - t0:.1f} s")
    r.close()
    bad += against_reference_shader(noise, max(8, n_poses // 5))
    return 1 if bad else 0


def against_reference_shader(noise, n_poses):
    """The same idea against the reference's own shader compiled as C++ (oracle/_ref, when the library travelled here):
    random poses, times and chunk subsets in dispatch shapes the unmodified shader can address (limits 5 / 15,
    grid_x = (W / 2) / chunk^2, spp 8..64) — GPU image vs reference-shader image, bit for bit."""
    from oracle import ref_shader
    if not ref_shader.available():
        print("reference-shader stress: oracle/_ref/libref_shader.so not present, skipped")
        return 0
    rng = np.random.default_rng(77)
    bad, t0, px = 0, time.time(), 0
    for i in range(n_poses):
        maze = int(rng.choice([10, 16, 32]))
        chunk = int(rng.choice([2, 4, 8]))
        spp = int(rng.choice([s for s in (8, 16, 32, 64) if chunk * chunk * s <= 1024]))
        gx = int(rng.choice([2, 4, 8])); gy = int(rng.integers(1, 7))
        W = 2 * chunk * chunk * gx; H = int(rng.choice([32, 64, 96]))
        sc = mm.MazeScene(maze, 0)
        r = mm.Renderer(0)                  # fresh zero-filled screen: the dispatch renders a subset of the chunks and the
        r.upload_scene(sc, noise)           # library's screen is persistent across calls (like the reference's texture)
        half = 5.0 * maze
        cell = rng.integers(0, maze, size=2)
        center = (-half + 10.0 * cell[0] + float(rng.uniform(0.5, 9.5)), float(rng.uniform(-7.5, 1.9)), -half + 10.0 * cell[1] + float(rng.uniform(0.5, 9.5)))
        u = mm.default_uniform(maze, W, H, chunk, time=int(rng.integers(0, 100000)), camera_center=center, half_theta=float(rng.uniform(0.0, np.pi)))
        allc = mm.gen_chunks(W, H, chunk)
        ch = allc[rng.permutation(len(allc))[: gx * gy]].copy()
        if len(ch) < gx * gy:
            r.close()
            continue
        p = mm.full_frame_params(u, spp=spp, bounce_limit=5, mirror_limit=15)
        p.grid_x, p.grid_y = gx, gy
        img, cnt, _ = r.render(u, p, ch)
        r.close()
        ref = ref_shader.render(sc, noise, u, p, ch)
        ok = img.tobytes() == ref.tobytes()
        bad += 0 if ok else 1
        px += int((ref[..., 3] == 1).sum())
        if not ok:
            print("REFERENCE-SHADER MISMATCH pose", i, maze, chunk, spp, gx, gy, center)
    print(f"reference-shader stress: {n_poses} random dispatches (mazes 10/16/32, chunk 2/4/8, spp 8..64), {px} pixels, "
          f"GPU vs the reference's own shader: mismatches: {bad}, {time.time() - t0:.1f} s")
    return bad

if __name__ == "__main__":
    sys.exit(main())
