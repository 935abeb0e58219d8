"""ctypes wrapper of oracle/_ref/libref_shader.so: the reference's OWN compute_shader (reference src/shaders.metal),
compiled unmodified as C++ through oracle/ref_shader/msl_shim.h and run on the CPU (fibers emulate the threadgroup).
TEST INFRASTRUCTURE — it pins oracle/mm_oracle.cpp to the reference's shader text.  Built by oracle/Makefile where
/root/reference exists; the binary travels to the GPU box, the reference tree does not."""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = os.path.join(_HERE, "_ref", "libref_shader.so")
_lib = None

BOUNCE_LIMIT, MIRROR_LIMIT = 5, 15          # literals of shaders.metal:294-295, fixed in the unmodified shader


def available():
    return os.path.exists(_LIB)


def lib():
    global _lib
    if _lib is None:
        _lib = C.CDLL(_LIB)
        _lib.ref_compute_shader.restype = C.c_int
        _lib.ref_compute_shader.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32,
                                            C.c_uint32, C.c_void_p, C.c_void_p, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32,
                                            C.c_void_p, C.c_int]
        b, m = C.c_int(), C.c_int()
        _lib.ref_shader_limits(C.byref(b), C.byref(m))
        assert (b.value, m.value) == (BOUNCE_LIMIT, MIRROR_LIMIT)
    return _lib


def dispatch_shape(uniform, spp):
    """The virtual-dispatch threadgroup shape for `spp` samples: T = chunk^2 * spp, dims = (min(32,T), T/min(32,T))."""
    T = int(uniform.chunk_width) ** 2 * int(spp)
    dx = min(32, T)
    return dx, T // dx


def addressable(uniform, params):
    """True when the unmodified shader can run this dispatch: its literals for the limits, its chunk lookup
    `tgid.x + tgid.y * ((width / 2) / ppc)` (shaders.metal:266), its three unconditional pairwise phases (spp >= 8)
    and its 1024-entry threadgroup array."""
    ppc = int(uniform.chunk_width) ** 2
    return (params.bounce_limit == BOUNCE_LIMIT and params.mirror_limit == MIRROR_LIMIT and ppc * params.spp <= 1024
            and params.spp >= 8 and int((np.float32(uniform.view_width) / np.float32(2.0)) / np.float32(ppc)) == params.grid_x
            and params.group_first == 0 and params.group_step in (0, 1) and params.group_count in (0, params.grid_x * params.grid_y))


def render(scene, noise, uniform, params, chunks, threads=0, out=None):
    """One dispatch of the reference's compute_shader -> image[H, W, 4] float32 (pixels not written stay 0 / `out`)."""
    if not addressable(uniform, params):
        raise ValueError("the unmodified reference shader cannot address this dispatch")
    planes = np.ascontiguousarray(scene.planes)
    nodes = np.ascontiguousarray(scene.nodes)
    indices = np.ascontiguousarray(scene.indices, dtype=np.uint32)
    materials = np.ascontiguousarray(scene.materials, dtype=np.uint8)
    emissions = np.ascontiguousarray(scene.emissions, dtype=np.float32)
    noise = np.ascontiguousarray(noise, dtype=np.uint8)
    chunks = np.ascontiguousarray(chunks)
    nh, nw = noise.shape[:2]
    H, W = int(uniform.view_height), int(uniform.view_width)
    img = out if out is not None else np.zeros((H, W, 4), dtype=np.float32)
    dx, dy = dispatch_shape(uniform, params.spp)
    rc = lib().ref_compute_shader(planes.ctypes.data, nodes.ctypes.data, indices.ctypes.data, materials.ctypes.data,
                                  emissions.ctypes.data, noise.ctypes.data, nw, nh, C.addressof(uniform), chunks.ctypes.data,
                                  params.grid_x, params.grid_y, dx, dy, img.ctypes.data, threads)
    if rc != 0:
        raise ValueError(f"ref_compute_shader refused the dispatch ({rc})")
    return img
