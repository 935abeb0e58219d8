"""f-1 progressive chunk bag + present blur (reference src/main.rs:293-326,778-784;
src/shaders.metal:214-225), f-3 scripted camera movement with collision (src/main.rs:786-826, 265-291)."""
import numpy as np
import pytest

from cases import build_case

F = np.float32


def test_chunk_bag_matches_random_pixels_semantics(mm):
    from oracle import host_ref

    W, H, chunk, seed = 64, 48, 4, 5
    bag = mm.ChunkBag(W, H, chunk, seed)
    # Python restatement: gen_pixels order, Fisher-Yates from the top with StdRng, pop from the end, refill with a clone
    rng = host_ref.StdRng(seed)
    original = [(chunk * i, chunk * j) for i in range(W // chunk) for j in range(H // chunk)]
    for i in range(len(original) - 1, 0, -1):
        j = rng.gen_range(0, i + 1)
        original[i], original[j] = original[j], original[i]
    pixels = list(original)
    for take in (7, 100, 85, 192, 3):                                     # crosses the refill boundary (192 chunks)
        want = []
        for _ in range(take):
            if not pixels:
                pixels = list(original)
            want.append(pixels.pop())
        got = bag.next(take)
        assert [(int(c["x"]), int(c["y"])) for c in got] == want
    assert len(bag) == len(pixels)
    seen = mm.ChunkBag(W, H, chunk, seed).next(192)
    assert sorted((int(c["x"]), int(c["y"])) for c in seen) == sorted(original)   # one full bag covers the screen once


def test_move_camera_and_collision(mm, scenes):
    sc = scenes(10)
    u = mm.default_uniform(10, 64, 64)
    q = [u.cam.rotation.x, u.cam.rotation.y, u.cam.rotation.z, u.cam.rotation.w]
    c0 = np.array([u.cam.camera_center.x, u.cam.camera_center.y, u.cam.camera_center.z], dtype=F)
    c1, blocked = mm.move_camera(sc.nodes, c0, q, [13], fps=60.0)          # W: forward along the rotated +z
    assert not blocked
    step = mm.quat_mult([0.0, 0.0, F(5.0) / F(60.0)], q)
    assert c1.tobytes() == (c0 + step).astype(F).tobytes()
    c2, _ = mm.move_camera(sc.nodes, c1, q, [1], fps=60.0)                 # S undoes it (up to rounding)
    assert np.allclose(c2, c0, atol=1e-6)
    c3, _ = mm.move_camera(sc.nodes, c0, q, [99], fps=60.0)                # unknown key: no move
    assert c3.tobytes() == c0.tobytes()
    near_wall = np.array([-5.0, 0.0, -49.6], dtype=F)                     # 0.4 from the z = -50 outer wall, facing it
    c4, blocked = mm.move_camera(sc.nodes, near_wall, q, [1], fps=60.0)    # S moves backwards into the wall
    assert blocked and c4.tobytes() == near_wall.tobytes()


def test_blur_reference_model():
    from oracle import np_oracle

    img = np.zeros((3, 4, 4), dtype=F)
    img[1, 1] = [3.0, 6.0, 9.0, 1.0]
    out = np_oracle.present_blur(img)
    assert out[1, 1, :3].tolist() == [1.0, 2.0, 3.0]                      # centre / 3
    assert out[1, 2, :3].tolist() == [0.5, 1.0, 1.5] and out[0, 1, :3].tolist() == [0.5, 1.0, 1.5]   # neighbour / 2 / 3
    assert out[0, 0, :3].tolist() == [0.0, 0.0, 0.0] and (out[..., 3] == 1).all()


@pytest.mark.gpu
def test_present_blur_matches_model(mm, renderer):
    import torch
    from oracle import np_oracle

    rng = np.random.default_rng(3)
    for (H, W) in ((1, 1), (5, 7), (33, 257), (270, 480)):
        img = rng.random((H, W, 4), dtype=np.float32)
        src = torch.from_numpy(img).cuda()
        dst = torch.empty_like(src)
        torch.cuda.synchronize()
        renderer.present_blur_device(src.data_ptr(), dst.data_ptr(), W, H)
        renderer.sync()
        assert dst.cpu().numpy().tobytes() == np_oracle.present_blur(img).tobytes()


@pytest.mark.gpu
def test_progressive_refresh_frame_loop(mm, oracle, noise, scenes):
    """Three frames of the reference's loop (main.rs:778-784, 867-893): pop a bag of chunks, render them into the
    persistent screen, blur the whole screen — GPU against the oracle + numpy composition, bit for bit."""
    from oracle import np_oracle

    sc, u, p, _ = build_case(mm, "yaw", scenes)
    W, H = int(u.view_width), int(u.view_height)
    r = mm.Renderer(0)
    r.upload_scene(sc, noise)
    bag_gpu, bag_cpu = mm.ChunkBag(W, H, 4, seed=9), mm.ChunkBag(W, H, 4, seed=9)
    q = mm.Params.from_buffer_copy(bytes(p))
    q.grid_x, q.grid_y = 16, 6                                             # 96 of 768 chunks per frame
    screen = np.zeros((H, W, 4), dtype=F)
    out = np.zeros((H, W, 4), dtype=F)
    for frame in range(3):
        u.time = frame
        chunks = bag_gpu.next(96)
        r.render(u, q, chunks)
        r.present(out)
        ref_chunks = bag_cpu.next(96)
        assert chunks.tobytes() == ref_chunks.tobytes()
        oracle.render(sc, noise, u, q, ref_chunks, out=screen)             # writes only the rendered chunks
        screen = np_oracle.present_blur(screen)
        assert out.tobytes() == screen.tobytes(), f"frame {frame}"
    r.close()


def test_rgba8_screen_quantisation_model(mm, oracle, noise, scenes):
    """MM_FLAG_SCREEN_RGBA8 (the reference's screen is RGBA8Unorm, main.rs:702-709): the oracle's stored pixels are the fp32
    frame pushed through rte(clamp(v) * 255) / 255, i.e. exact 8-bit values; the blur model quantises its output the same way."""
    from cases import build_case
    from oracle import np_oracle

    sc, u, p, ch = build_case(mm, "yaw", scenes)
    plain = oracle.render(sc, noise, u, p, ch)[0]
    p.flags = mm.FLAG_SCREEN_RGBA8
    q = oracle.render(sc, noise, u, p, ch)[0]
    assert q.tobytes() == np_oracle.quant8(plain).tobytes()
    k = q * np.float32(255.0)
    assert np.array_equal(np.rint(k), k.round(3)) and q.min() >= 0 and q.max() <= 1          # values are k / 255
    assert np_oracle.quant8(np.array([np.nan, -1.0, 0.5, 2.0, 0.5 / 255, 1.5 / 255], np.float32)).tolist() == \
        [0.0, 0.0, float(np.float32(128.0) / np.float32(255.0)), 1.0, 0.0, float(np.float32(2.0) / np.float32(255.0))]   # ties to even
    b = np_oracle.present_blur(q, rgba8=True)
    assert b.tobytes() == np_oracle.quant8(np_oracle.present_blur(q)).tobytes()


@pytest.mark.gpu
def test_rgba8_screen_mode_matches_oracle_and_blur_model(mm, oracle, noise, scenes):
    """The CUDA path with an RGBA8Unorm screen: dispatch == oracle under the same flag; two frames of progressive refresh + the
    quantising present blur == oracle + model; the byte frame is the texels."""
    from cases import build_case
    from oracle import np_oracle

    sc, u, p, ch = build_case(mm, "yaw", scenes)
    H, W = int(u.view_height), int(u.view_width)
    r = mm.Renderer(0)
    r.upload_scene(sc, noise)
    for flags in (mm.FLAG_SCREEN_RGBA8, mm.FLAG_SCREEN_RGBA8 | mm.FLAG_POOL_KERNEL):
        p.flags = flags
        q = mm.Params.from_buffer_copy(bytes(p)); q.flags = mm.FLAG_SCREEN_RGBA8
        assert r.render(u, p, ch)[0].tobytes() == oracle.render(sc, noise, u, q, ch)[0].tobytes()
    r.close()
    r = mm.Renderer(0)
    r.upload_scene(sc, noise)
    screen = np.zeros((H, W, 4), dtype=np.float32)
    half = mm.Params.from_buffer_copy(bytes(p)); half.flags = mm.FLAG_SCREEN_RGBA8
    n = p.grid_x * p.grid_y
    for frame in range(2):
        half.group_first, half.group_step, half.group_count = frame, 2, (n + 1 - frame) // 2
        u.time = frame
        r.render(u, half, ch)
        oracle.render(sc, noise, u, half, ch, out=screen)
        out, out8 = np.empty((H, W, 4), np.float32), np.empty((H, W, 4), np.uint8)
        r.present_rgba8(out, out8)
        screen = np_oracle.present_blur(screen, rgba8=True)
        assert out.tobytes() == screen.tobytes()
        assert np.array_equal(out8, np.rint(screen * np.float32(255.0)).astype(np.uint8))
    # the asynchronous form: one more frame, texels read back on the second stream into pinned memory
    half.group_first, half.group_count = 0, (n + 1) // 2
    u.time = 2
    r.render(u, half, ch)
    oracle.render(sc, noise, u, half, ch, out=screen)
    hf = mm.HostFrame(H, W)                                   # pinned; its first H*W*4 bytes receive the texels
    r.present_async_rgba8(hf.ptr)
    r.wait_present()
    screen = np_oracle.present_blur(screen, rgba8=True)
    texels = np.frombuffer(hf.array.tobytes()[: H * W * 4], dtype=np.uint8).reshape(H, W, 4)
    assert np.array_equal(texels, np.rint(screen * np.float32(255.0)).astype(np.uint8))
    hf.close()
    r.close()


@pytest.mark.gpu
def test_present_async_delivers_the_frames_of_the_synchronous_present(mm, noise, scenes):
    """mm_present_async (blur, device snapshot, read-back on a second stream while the next dispatch runs) into two alternating
    pinned frames == mm_present frame by frame; an unpinned buffer is refused."""
    from cases import build_case

    sc, u, p, ch = build_case(mm, "yaw", scenes)
    H, W = int(u.view_height), int(u.view_width)
    ra, rb = mm.Renderer(0), mm.Renderer(0)
    ra.upload_scene(sc, noise); rb.upload_scene(sc, noise)
    frames = [mm.HostFrame(H, W), mm.HostFrame(H, W)]
    half = mm.Params.from_buffer_copy(bytes(p))
    n = p.grid_x * p.grid_y
    expect = []
    for f in range(6):
        half.group_first, half.group_step, half.group_count = f % 3, 3, (n - f % 3 + 2) // 3
        u.time = f
        ra.render(u, half, ch)
        out = np.empty((H, W, 4), np.float32)
        ra.present(out)
        expect.append(out)
        rb.render(u, half, ch)                             # frame f's dispatch overlaps frame f - 1's read-back (even f - 1)
        rb.present_async(frames[f & 1].ptr)
        if f & 1:                                          # every second frame: wait, then both buffers hold finished frames
            rb.wait_present()
            assert frames[1].array.tobytes() == expect[f].tobytes(), f
            assert frames[0].array.tobytes() == expect[f - 1].tobytes(), f
    with pytest.raises(mm.MMError):
        rb.present_async(np.empty((H, W, 4), np.float32).ctypes.data)
    ra.close(); rb.close()
    for fr in frames:
        fr.close()


@pytest.mark.gpu
def test_blur_divide_by_three_is_exact_for_every_float(renderer):
    """blur_kernel's x / 3 (three FMA-pipe operations, guarded for tiny values) == __fdiv_rn(x, 3) on all 2^32 bit patterns."""
    assert renderer.selftest_div3() == 0
