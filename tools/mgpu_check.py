"""Run under torchrun (one rank per GPU): the N-GPU frame must be bit-identical to the frame one GPU renders alone, for
every exchange: NCCL all-gather + scatter, fused NVLink peer stores, fused NVSwitch multicast stores."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.distributed as dist
import mirror_maze_b200 as mm

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
noise = mm.load_noise()
ok = True
for (n, W, H, spp, b) in [(16, 256, 128, 8, 6), (32, 1920, 1080, 16, 8)]:
    sc = mm.MazeScene(n, 0)
    u = mm.default_uniform(n, W, H, 4)
    ch = mm.gen_chunks(W, H, 4)
    p = mm.full_frame_params(u, spp=spp, bounce_limit=b)
    r = mm.Renderer(local)
    r.upload_scene(sc, noise)
    r2 = mm.Renderer(local)
    r2.upload_scene(sc, noise)
    full, cnt, _ = r2.render(u, p, ch)
    for exchange, multicast in (("gather", False), ("peer", False), ("peer", True)):
        try:
            fr = mm.TiledFrameRenderer(r, u, p, ch, rank=rank, world=world, dist=dist, exchange=exchange, multicast=multicast)
        except Exception as e:
            print(f"rank {rank}: exchange {exchange} multicast={multicast} unavailable: {type(e).__name__}: {e}", flush=True)
            ok = False
            continue
        same = True
        for it in range(3):                       # frames back to back: the barriers must keep them apart
            with torch.cuda.stream(fr.stream):
                fr.image.zero_()                  # a store that does not arrive must show
                img = fr.render_frame(u)
            fr.stream.synchronize()
            same &= img.cpu().numpy().tobytes() == full.tobytes()
        print(f"rank {rank}/{world} maze {n} {W}x{H} exchange={fr.exchange} ({fr.exchange_note}): == single-GPU frame: {same}", flush=True)
        ok &= same
        dist.barrier()
    r.close(); r2.close()
flag = torch.tensor([1 if ok else 0], device=f"cuda:{local}")
dist.all_reduce(flag, op=dist.ReduceOp.MIN)
if rank == 0:
    print("MULTI-GPU PARITY", "OK" if flag.item() == 1 else "FAILED", flush=True)
dist.destroy_process_group()
sys.exit(0 if flag.item() == 1 else 1)
