"""Host-side logic of the multi-GPU tile path on CPU: the interleaved group partition covers the grid exactly once,
and tiles gathered over torch.distributed (gloo, world_size 2) scatter back into the one-GPU frame.  The per-tile
content comes from the CPU oracle here (this is a test of the plumbing, not of the kernel)."""
import os
import socket

import numpy as np
import pytest

from cases import build_case


def test_partition_covers_every_group_once(mm):
    for n_groups in (1, 2, 7, 768, 129600):
        for world in (1, 2, 3, 4, 8):
            seen = []
            for rank in range(world):
                first, step, count = mm.tile_partition(n_groups, rank, world)
                seen += [first + k * step for k in range(count)]
            assert sorted(seen) == list(range(n_groups))
            counts = [mm.tile_partition(n_groups, r, world)[2] for r in range(world)]
            assert max(counts) - min(counts) <= 1                       # balanced


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, q):
    import sys

    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    import torch
    import torch.distributed as dist

    import mirror_maze_b200 as mm
    from mirror_maze_b200.renderer import scatter_tiles_host
    from oracle import oracle

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        noise = mm.load_noise()
        sc, u, p, ch = build_case(mm, "chunk2_spp4")
        n_groups = p.grid_x * p.grid_y
        ppc = u.chunk_width ** 2
        parts = [mm.tile_partition(n_groups, r, world) for r in range(world)]
        max_count = max(pt[2] for pt in parts)
        first, step, count = parts[rank]
        mine = mm.Params.from_buffer_copy(bytes(p))
        mine.group_first, mine.group_step, mine.group_count = first, step, count
        img = oracle.render(sc, noise, u, mine, ch)[0]
        # compact tiles in the kernel's layout: tile k, pixel pn -> (x + pn // chunk, y + pn % chunk)
        tiles = torch.zeros((max_count, ppc, 4), dtype=torch.float32)
        pn = np.arange(ppc)
        for k in range(count):
            c = ch[first + k * step]
            tiles[k] = torch.from_numpy(img[int(c["y"]) + pn % u.chunk_width, int(c["x"]) + pn // u.chunk_width])
        flat = torch.zeros((world * max_count, ppc, 4), dtype=torch.float32)     # concatenated layout (gloo and nccl)
        dist.all_gather_into_tensor(flat, tiles)
        gathered = flat.view(world, max_count, ppc, 4)
        frame = np.zeros_like(img)
        for r, (f, s, cnt) in enumerate(parts):
            scatter_tiles_host(gathered[r].numpy(), ch, f, s, cnt, u.chunk_width, frame)
        full = oracle.render(sc, noise, u, p, ch)[0]
        q.put((rank, frame.tobytes() == full.tobytes()))
    except Exception as e:                                              # report instead of letting the parent time out
        q.put((rank, repr(e)))
    finally:
        dist.destroy_process_group()


def test_gloo_world2_gather_and_scatter(mm, oracle):
    import torch.multiprocessing as mp

    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for pr in procs:
        pr.start()
    results = [q.get(timeout=120) for _ in procs]
    for pr in procs:
        pr.join(timeout=60)
    assert sorted(results) == [(0, True), (1, True)]


def _shared_frame_worker(rank, world, port, q):
    import sys

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, root)
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    import torch.distributed as dist

    import bench
    import mirror_maze_b200 as mm
    from oracle import oracle

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        noise = mm.load_noise()
        sc, u, p, ch = build_case(mm, "ragged")                       # 84 groups, frame not a multiple of the chunk
        H, W = int(u.view_height), int(u.view_width)
        name = f"mm_test_frame_{port}"
        shared = bench.SharedHostFrame(mm, name, H * W * 16, create=True, register=False) if rank == 0 else None
        dist.barrier()
        if rank != 0:
            shared = bench.SharedHostFrame(mm, name, H * W * 16, create=False, register=False)
        mine = mm.Params.from_buffer_copy(bytes(p))
        mine.group_first, mine.group_step, mine.group_count = mm.tile_partition(p.grid_x * p.grid_y, rank, world)
        # every rank writes only the pixels of its own groups into the ONE frame (what the kernels' zero-copy stores do)
        oracle.render(sc, noise, u, mine, ch, out=shared.array.reshape(H, W, 4))
        dist.barrier()
        ok = True
        if rank == 0:
            full = oracle.render(sc, noise, u, p, ch)[0]
            ok = shared.array.tobytes() == full.tobytes()
        dist.barrier()
        shared.close()
        q.put((rank, ok))
    except Exception as e:
        q.put((rank, repr(e)))
    finally:
        dist.destroy_process_group()


def test_gloo_world2_shared_host_frame_assembly(mm, oracle):
    """bench.py's N > 1 end-to-end leg: every rank stores the pixels of its interleaved groups into one frame in shared host
    memory (/dev/shm mapping; on a GPU box it is also pinned + mapped into each rank's CUDA context) — the union is the
    one-rank frame, no gather and no copy."""
    import torch.multiprocessing as mp

    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_shared_frame_worker, args=(r, 2, port, q)) for r in range(2)]
    for pr in procs:
        pr.start()
    results = [q.get(timeout=120) for _ in procs]
    for pr in procs:
        pr.join(timeout=60)
    assert sorted(results) == [(0, True), (1, True)]
