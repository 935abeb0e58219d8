"""The C-ABI library loads on a CPU-only box, exports every symbol include/mirror_maze_cuda.h declares, keeps the
reference's byte layouts (reference src/main.rs:32-90) and fails loudly — no fallback — without a GPU."""
import ctypes as C
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "mirror_maze_cuda.h")


def declared_functions():
    text = open(HEADER).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(mm_[a-z0-9_]+)\s*\(", text)))


def test_header_declares_expected_surface():
    names = declared_functions()
    for must in ("mm_create", "mm_destroy", "mm_upload_scene", "mm_render", "mm_render_device", "mm_scatter_tiles_device",
                 "mm_set_chunks", "mm_sync", "mm_last_ms", "mm_last_error", "mm_scene_build", "mm_build_bvh"):
        assert must in names


def test_library_exports_every_declared_symbol(mm):
    lib = C.CDLL(mm.library_path())
    missing = [n for n in declared_functions() if not hasattr(lib, n)]
    assert not missing, f"declared in the header but not exported: {missing}"


def test_binding_covers_every_declared_symbol(mm):
    from mirror_maze_b200 import abi

    assert sorted(abi.PROTOTYPES) == declared_functions()


def test_struct_layouts_match_reference(mm):
    assert C.sizeof(mm.Float2) == 8 and C.sizeof(mm.Float3) == 12 and C.sizeof(mm.Float4) == 16   # maths.rs:3-16,50-52
    assert C.sizeof(mm.Plane) == 48                                                              # main.rs:51-58
    assert [mm.Plane.origin.offset, mm.Plane.v.offset, mm.Plane.u.offset, mm.Plane.color.offset] == [0, 12, 24, 36]
    assert C.sizeof(mm.BVHNode) == 32                                                            # main.rs:74-81
    assert [mm.BVHNode.aabb_min.offset, mm.BVHNode.aabb_max.offset, mm.BVHNode.left_first.offset,
            mm.BVHNode.tri_count.offset] == [0, 12, 24, 28]
    assert C.sizeof(mm.Camera) == 40                                                             # main.rs:32-39
    assert [mm.Camera.camera_center.offset, mm.Camera.focal_length.offset, mm.Camera.rotation.offset,
            mm.Camera.viewport.offset] == [0, 12, 16, 32]
    assert C.sizeof(mm.Uniform) == 56                                                            # main.rs:41-49
    assert [mm.Uniform.cam.offset, mm.Uniform.view_width.offset, mm.Uniform.view_height.offset,
            mm.Uniform.chunk_width.offset, mm.Uniform.time.offset] == [0, 40, 44, 48, 52]
    assert C.sizeof(mm.Chunk) == 8


def test_version_string(mm):
    assert b"sm_100a" in mm.load_library().mm_version()


def test_no_cpu_fallback_without_gpu(mm):
    """On a box without a GPU the render path must refuse to exist, not degrade."""
    import torch

    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(mm.MMError) as e:
        mm.Renderer(0)
    assert e.value.code == -2          # MM_ERR_CUDA
    assert "no CUDA device" in str(e.value) or "CUDA" in str(e.value)


def test_null_arguments_are_errors_not_crashes(mm):
    lib = mm.load_library()
    assert lib.mm_create(0, None) == -1
    assert lib.mm_scene_build(0, 0, 1, None) == -1
    h = C.c_void_p()
    assert lib.mm_scene_build(0, 0, 1, C.byref(h)) == -1      # empty maze
    assert lib.mm_sync(None) == -1
    assert lib.mm_destroy(None) == 0
    assert lib.mm_last_error(None) is not None


def test_product_never_imports_oracle():
    """The product path must not route through oracle/ (parity claims would be void)."""
    pkg = os.path.join(ROOT, "mirror_maze_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cpp", ".cu", ".h", ".cuh")) or f == "Makefile":
                text = open(os.path.join(dirpath, f), errors="ignore").read()
                for needle in ("libmm_oracle", "from oracle", "import oracle", "np_oracle", "mmo_render", "oracle/_ref", "dlopen"):
                    if needle == "dlopen" and f == "multi.cu":
                        # the one run-time load in the product: libnccl for MM_EXCHANGE_NCCL, and nothing else
                        libs = re.findall(r'"([^"]*\.so[^"]*)"', text)
                        assert libs and all(name.startswith("libnccl.so") for name in libs), libs
                        continue
                    assert needle not in text, f"{f} references the oracle ({needle})"
                assert not re.search(r'#include\s*["<][^">]*oracle', text), f"{f} includes oracle code"


def test_headless_driver_fails_loudly_without_gpu(mm):
    import subprocess
    import torch

    if torch.cuda.is_available():
        pytest.skip("GPU present")
    exe = os.path.join(os.path.dirname(mm.library_path()), "mm_headless")
    out = subprocess.run([exe, "--maze", "4", "--width", "16", "--height", "16", "--spp", "1"], capture_output=True, text=True, timeout=120)
    assert out.returncode != 0 and "mm_create failed" in out.stderr


def test_packed_fp32_instruction_mix_shows_no_contraction():
    """ptxas contracts mul.rn.f32x2 + add.rn.f32x2 into FFMA2 regardless of the rounding modifiers (docs/exact_quotient.md),
    which would round once where the contract rounds twice.  The slab sequence is add, mul, fma, fma, fma per quotient pair
    (add, mul in the reciprocal-multiply mode), six pairs per visit, two unrolled visits per traversal instantiation.  Every
    trace kernel holds the exact traversal in 3 instantiations (plain, mixed-literal, axis-aligned rects; trace_kernel_rg: 2)
    and the reciprocal-multiply one in 2, so its SASS must hold exactly FADD2 : FMUL2 : FFMA2 = 60 : 60 : 108 (48 : 48 : 72).
    A fused or dropped instruction changes the counts."""
    import shutil
    import subprocess
    obj = os.path.join(ROOT, "mirror_maze_b200", "build", "render_kernel.o")
    tool = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    if not (os.path.exists(obj) and os.path.exists(tool)):
        pytest.skip("needs the built render_kernel.o and cuobjdump")
    sass = subprocess.run([tool, "-sass", obj], capture_output=True, text=True, check=True).stdout
    counts, name = {}, None
    for line in sass.splitlines():
        if "Function :" in line:
            name = line.split("Function :")[1].strip()
            counts[name] = [0, 0, 0]
        elif name:
            for i, op in enumerate(("FADD2", "FMUL2", "FFMA2")):
                if f" {op} " in line:
                    counts[name][i] += 1
    kernels = {n: c for n, c in counts.items() if "trace_kernel" in n}
    assert len(kernels) >= 6
    for n, c in kernels.items():
        assert tuple(c) == ((48, 48, 72) if "trace_kernel_rg" in n else (60, 60, 108)), (n, c)


def test_new_entry_points_fail_cleanly_without_a_gpu(mm):
    """mm_multi_create / mm_host_alloc / mm_render_async on a box without a CUDA device: error codes, no crash, no fallback."""
    import torch

    if torch.cuda.is_available():
        pytest.skip("GPU present")
    lib = mm.load_library()
    devs = (C.c_int * 2)(0, 1)
    h = C.c_void_p()
    assert lib.mm_multi_create(devs, 2, mm.EXCHANGE_PEER, C.byref(h)) < 0 and not h.value
    assert b"device" in lib.mm_multi_last_error(None).lower() or lib.mm_multi_last_error(None) != b""
    assert lib.mm_multi_create(devs, 0, mm.EXCHANGE_PEER, C.byref(h)) == -1
    assert lib.mm_multi_create(devs, 2, 7, C.byref(h)) == -1
    dup = (C.c_int * 2)(0, 0)
    assert lib.mm_multi_create(dup, 2, mm.EXCHANGE_NCCL, C.byref(h)) == -1
    assert lib.mm_multi_destroy(None) == 0 and lib.mm_multi_n_devices(None) == 0
    p = C.c_void_p()
    assert lib.mm_host_alloc(1 << 20, C.byref(p)) < 0 and not p.value
    assert lib.mm_host_alloc(0, C.byref(p)) == -1
    assert lib.mm_host_free(None) == 0 and lib.mm_host_unregister(None) == 0
    assert lib.mm_host_free(C.c_void_p(12345)) == -1               # never allocated here
    assert lib.mm_render_async(None, None, None, None, 0, None, None) == -1
    assert lib.mm_wait(None, None) == -1
