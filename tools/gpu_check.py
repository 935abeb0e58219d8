"""Quick GPU bring-up check (developer tool): parity of the CUDA path against the CPU oracle on small cases in every
kernel mode, then a timing of the north-star configuration.  Run under gpurun."""
import sys, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import mirror_maze_b200 as mm
from oracle import oracle


def compare(tag, a, b):
    img, cnt, dbg = a
    img2, cnt2, dbg2 = b
    ok = True
    for k in ("first_hit", "segments", "mirror_hits", "radiance"):
        eq = np.array_equal(dbg[k].view(np.uint32), dbg2[k].view(np.uint32))
        if not eq:
            bad = np.nonzero((dbg[k].reshape(len(dbg['segments']), -1) != dbg2[k].reshape(len(dbg['segments']), -1)).any(axis=1))[0]
            print(f"  {tag}: {k} differs on {len(bad)} paths, first {bad[:5]}")
        ok &= eq
    ieq = np.array_equal(img.view(np.uint32), img2.view(np.uint32))
    if not ieq:
        print(f"  {tag}: image max abs diff {np.abs(img - img2).max()}")
    ok &= ieq
    for k in ("paths", "rays", "inner_visits", "leaf_visits", "rect_tests", "hits", "max_stack"):
        if cnt[k] != cnt2[k]:
            print(f"  {tag}: counter {k}: gpu {cnt[k]} oracle {cnt2[k]}")
            ok = False
    print(f"{tag}: {'BIT-EXACT' if ok else 'MISMATCH'}  literal_rays={cnt['literal_rays']} rays={cnt['rays']}")
    return ok


def main():
    noise = mm.load_noise()
    r = mm.Renderer(0)
    allok = True
    for (n, W, H, spp, b) in [(16, 256, 256, 1, 4), (16, 128, 64, 16, 8), (10, 64, 64, 64, 5), (32, 96, 64, 8, 8)]:
        sc = mm.MazeScene(n, 0)
        r.upload_scene(sc, noise)
        u = mm.default_uniform(n, W, H, 4)
        ch = mm.gen_chunks(W, H, 4)
        p = mm.full_frame_params(u, spp=spp, bounce_limit=b)
        ref = oracle.render(sc, noise, u, p, ch, debug=True)
        for name, flags in (("fast", 0), ("general rects", mm.FLAG_GENERAL_RECTS), ("literal", mm.FLAG_FORCE_LITERAL)):
            p.flags = flags
            got = r.render(u, p, ch, debug=True)
            allok &= compare(f"N={n} {W}x{H} spp={spp} b={b} [{name}]", got, ref)
        print("   info", r.scene_info())
    # timing: north-star configs
    for n in (32, 64):
        sc = mm.MazeScene(n, 0)
        r.upload_scene(sc, noise)
        u = mm.default_uniform(n, 1920, 1080, 4)
        ch = mm.gen_chunks(1920, 1080, 4)
        for name, flags in (("fast", 0), ("general rects", mm.FLAG_GENERAL_RECTS), ("literal", mm.FLAG_FORCE_LITERAL), ("rcp", mm.FLAG_RCP_SLAB),
                            ("rcp regroup", mm.FLAG_RCP_SLAB | mm.FLAG_REGROUP), ("fast+counters", mm.FLAG_COUNTERS)):
            p = mm.full_frame_params(u, spp=16, bounce_limit=8, flags=flags)
            for it in range(3):
                img, cnt, _ = r.render(u, p, ch)
                ms = r.last_ms()
            print(f"N={n} 1080p x16spp b8 [{name}]: {ms:.2f} ms  rays={cnt['rays']}  {cnt['rays']/ms/1e3:.1f} Mrays/s  literal_rays={cnt['literal_rays']} info={r.scene_info()}")
            if flags & mm.FLAG_COUNTERS:
                print("   counters", cnt)
    print("ALL OK" if allok else "FAILURES")
    return 0 if allok else 1


if __name__ == "__main__":
    sys.exit(main())
