"""Host-side mirror of the reference's data surface (reference src/main.rs:357-588, 732-755; src/maths.rs).

Everything here is a thin view over the C++ restatement exported by the C-ABI (csrc/maze_scene.cpp, bvh_build.cpp,
host_api.cpp): the maze, the wall lists, the plane / material / emission vectors, the BVH node and index arrays —
the exact buffers the reference hands to Metal with make_buf (src/main.rs:723-730).
"""
import ctypes as C
import gzip
import os

import numpy as np

from . import abi
from .abi import Float3, Float4, Uniform, Params, MMError

PLANE_DTYPE = np.dtype([("origin", "<f4", 3), ("v", "<f4", 3), ("u", "<f4", 3), ("color", "<f4", 3)])
NODE_DTYPE = np.dtype([("aabb_min", "<f4", 3), ("aabb_max", "<f4", 3), ("left_first", "<u4"), ("tri_count", "<u4")])
CHUNK_DTYPE = np.dtype([("x", "<u4"), ("y", "<u4")])
assert PLANE_DTYPE.itemsize == 48 and NODE_DTYPE.itemsize == 32 and CHUNK_DTYPE.itemsize == 8

_ASSETS = os.path.join(os.path.dirname(os.path.abspath(__file__)), "assets")


def _check(rc, what):
    if rc != 0:
        raise MMError(rc, what)


def _copy(ptr, count, dtype):
    if count == 0 or not ptr:
        return np.zeros(0, dtype=dtype)
    nbytes = count * np.dtype(dtype).itemsize
    return np.frombuffer(C.string_at(ptr, nbytes), dtype=dtype).copy()


class MazeScene:
    """main(): Kruskal maze -> walls -> planes/materials/emissions -> build_bvh (src/main.rs:357-588), n x n."""

    def __init__(self, maze_n, seed=0, fast_bvh=True):
        lib = abi.load_library()
        h = C.c_void_p()
        _check(lib.mm_scene_build(maze_n, seed, 1 if fast_bvh else 0, C.byref(h)), "mm_scene_build")
        try:
            self.maze_n = int(maze_n)
            self.seed = int(seed)
            P, N = lib.mm_scene_n_planes(h), lib.mm_scene_n_nodes(h)
            self.planes = _copy(lib.mm_scene_planes(h), P, PLANE_DTYPE)
            self.nodes = _copy(lib.mm_scene_nodes(h), N, NODE_DTYPE)
            self.indices = _copy(lib.mm_scene_indices(h), P, np.uint32)
            self.materials = _copy(lib.mm_scene_materials(h), P, np.uint8)
            self.emissions = _copy(lib.mm_scene_emissions(h), 4 * P, np.float32).reshape(P, 4)
            self.grid = _copy(lib.mm_scene_grid(h), maze_n * maze_n, np.uint8).reshape(maze_n, maze_n)
            self.vert_walls = _copy(lib.mm_scene_vert_walls(h), 3 * lib.mm_scene_n_vert_walls(h), np.float32).reshape(-1, 3)
            self.hori_walls = _copy(lib.mm_scene_hori_walls(h), 3 * lib.mm_scene_n_hori_walls(h), np.float32).reshape(-1, 3)
        finally:
            lib.mm_scene_free(h)

    @property
    def n_planes(self):
        return len(self.planes)

    @property
    def n_nodes(self):
        return len(self.nodes)


def build_bvh(planes, fast=True):
    """build_bvh (src/main.rs:247-263) on caller planes -> (nodes, indices)."""
    lib = abi.load_library()
    planes = np.ascontiguousarray(planes, dtype=PLANE_DTYPE)
    n = len(planes)
    nodes = np.zeros(max(2 * n - 1, 1), dtype=NODE_DTYPE)
    indices = np.zeros(n, dtype=np.uint32)
    nn = C.c_uint32()
    _check(lib.mm_build_bvh(planes.ctypes.data, n, 1 if fast else 0, nodes.ctypes.data, C.byref(nn), indices.ctypes.data),
           "mm_build_bvh")
    return nodes[: nn.value].copy(), indices


class StdRng:
    """rand 0.8.5 StdRng::seed_from_u64 (ChaCha12), as used at src/main.rs:381."""

    def __init__(self, seed):
        self._lib = abi.load_library()
        self._h = C.c_void_p()
        _check(self._lib.mm_stdrng_new(seed, C.byref(self._h)), "mm_stdrng_new")

    def next_u32(self):
        return int(self._lib.mm_stdrng_next_u32(self._h))

    def gen_f32(self):
        return float(self._lib.mm_stdrng_gen_f32(self._h))

    def gen_range(self, low, high):
        return int(self._lib.mm_stdrng_gen_range_u32(self._h, low, high))

    def __del__(self):
        if getattr(self, "_h", None):
            self._lib.mm_stdrng_free(self._h)
            self._h = None


def chacha_block(key32, counter, stream, rounds):
    lib = abi.load_library()
    key = (C.c_uint8 * 32)(*bytes(key32))
    out = (C.c_uint32 * 16)()
    _check(lib.mm_chacha_block(key, counter, stream, rounds, out), "mm_chacha_block")
    return np.array(out[:], dtype=np.uint32)


def calculate_quaternion(direction):
    q = abi.load_library().mm_calculate_quaternion(Float3(*[float(v) for v in direction]))
    return np.array([q.x, q.y, q.z, q.w], dtype=np.float32)


def update_quat_angle(q, theta):
    r = abi.load_library().mm_update_quat_angle(Float4(*[float(v) for v in q]), float(theta))
    return np.array([r.x, r.y, r.z, r.w], dtype=np.float32)


def quat_mult(v, q):
    r = abi.load_library().mm_quat_mult(Float3(*[float(x) for x in v]), Float4(*[float(x) for x in q]))
    return np.array([r.x, r.y, r.z], dtype=np.float32)


def gen_chunks(view_width, view_height, chunk_width):
    """gen_pixels without the shuffle (src/main.rs:293-302): x-major outer, y inner."""
    lib = abi.load_library()
    n = lib.mm_gen_chunks(float(view_width), float(view_height), chunk_width, None, 0)
    out = np.zeros(n, dtype=CHUNK_DTYPE)
    lib.mm_gen_chunks(float(view_width), float(view_height), chunk_width, out.ctypes.data, n)
    return out


def default_uniform(maze_n, view_width, view_height, chunk_width=4, time=0, camera_center=None, half_theta=None):
    """The uniform main() starts with (src/main.rs:732-755), generalised to n x n; optional pose override:
    camera_center (3 floats) and half_theta (the mouse-X yaw state, src/main.rs:741,829)."""
    u = Uniform()
    _check(abi.load_library().mm_default_uniform(maze_n, float(view_width), float(view_height), chunk_width, time, C.byref(u)),
           "mm_default_uniform")
    if camera_center is not None:
        u.cam.camera_center = Float3(*[float(v) for v in camera_center])
    if half_theta is not None:
        q = update_quat_angle([u.cam.rotation.x, u.cam.rotation.y, u.cam.rotation.z, u.cam.rotation.w], half_theta)
        u.cam.rotation = Float4(*[float(v) for v in q])
    return u


def full_frame_params(uniform, spp, bounce_limit, mirror_limit=15, flags=0):
    """Virtual dispatch covering the whole frame: one group per chunk of gen_chunks, grid = (W/chunk, H/chunk)."""
    gx = int(uniform.view_width) // uniform.chunk_width
    gy = int(uniform.view_height) // uniform.chunk_width
    return Params(spp=spp, bounce_limit=bounce_limit, mirror_limit=mirror_limit, grid_x=gx, grid_y=gy, group_first=0,
                  group_step=1, group_count=0, flags=flags)


def load_noise():
    """512x512 RGBA8 decode of the reference's textures/noiseTexture-2.png (embedded at src/main.rs:354, uploaded at
    :667-695).  Committed as a gzip of the raw bytes; tests/golden/make_noise_fixture.py regenerates it."""
    with gzip.open(os.path.join(_ASSETS, "noiseTexture-2.rgba8.gz"), "rb") as f:
        raw = f.read()
    a = np.frombuffer(raw, dtype=np.uint8).copy()
    assert a.size == 512 * 512 * 4
    return a.reshape(512, 512, 4)


def check_collision(nodes, bmin, bmax):
    nodes = np.ascontiguousarray(nodes, dtype=NODE_DTYPE)
    return int(abi.load_library().mm_check_collision(nodes.ctypes.data, len(nodes), Float3(*[float(v) for v in bmin]),
                                                     Float3(*[float(v) for v in bmax])))


def axis_rect(plane):
    """The axis-aligned form of one rect (mm_axis_rect): (c, lo_a, hi_a, lo_b, hi_b, k) or None when it is not axis-aligned."""
    import ctypes as C
    rec = np.ascontiguousarray(plane, dtype=PLANE_DTYPE).reshape(1)
    out, k = (C.c_float * 5)(), C.c_uint32()
    rc = abi.load_library().mm_axis_rect(rec.ctypes.data_as(C.POINTER(abi.Plane)), out, C.byref(k))
    if rc == -5:
        return None
    _check(rc, "mm_axis_rect")
    return tuple(np.float32(v) for v in out) + (int(k.value),)


def rect_edge_thresholds(length):
    """Interval [lo, up] on x = dot(rv, edge) equivalent to `0 <= RN(x / length) <= length` (mm_rect_edge_thresholds);
    None when the length is outside the guarded range."""
    import ctypes as C
    lo, up = C.c_float(), C.c_float()
    rc = abi.load_library().mm_rect_edge_thresholds(C.c_float(float(length)), C.byref(lo), C.byref(up))
    if rc == -5:                      # MM_ERR_UNSUPPORTED
        return None
    _check(rc, "mm_rect_edge_thresholds")
    return np.float32(lo.value), np.float32(up.value)


class ChunkBag:
    """The progressive-refresh bag of chunk origins: gen_pixels + random_pixels (src/main.rs:293-326, 713-720, 778-784),
    with a seeded StdRng in place of the reference's non-deterministic thread_rng."""

    def __init__(self, view_width, view_height, chunk_width=4, seed=0):
        self._lib = abi.load_library()
        self._h = C.c_void_p()
        _check(self._lib.mm_bag_new(float(view_width), float(view_height), chunk_width, seed, C.byref(self._h)), "mm_bag_new")

    def next(self, n):
        out = np.zeros(n, dtype=CHUNK_DTYPE)
        _check(self._lib.mm_bag_next(self._h, n, out.ctypes.data), "mm_bag_next")
        return out

    def reshuffle(self):
        _check(self._lib.mm_bag_reshuffle(self._h), "mm_bag_reshuffle")

    def __len__(self):
        return int(self._lib.mm_bag_size(self._h))

    def __del__(self):
        if getattr(self, "_h", None):
            self._lib.mm_bag_free(self._h)
            self._h = None


def move_camera(nodes, center, quat, keys, fps=60.0):
    """One frame of WASD movement with collision (src/main.rs:786-826).  Returns (new_center, blocked)."""
    nodes = np.ascontiguousarray(nodes, dtype=NODE_DTYPE)
    k = np.ascontiguousarray(keys, dtype=np.uint16)
    out = Float3()
    rc = abi.load_library().mm_move_camera(nodes.ctypes.data, len(nodes), Float3(*[float(v) for v in center]),
                                           Float4(*[float(v) for v in quat]), k.ctypes.data, len(k), float(fps), C.byref(out))
    if rc < 0:
        raise MMError(rc, "mm_move_camera")
    return np.array([out.x, out.y, out.z], dtype=np.float32), bool(rc)
