// ref_shader.cpp — runs the reference's own compute_shader on the CPU.
//
// The reference's device source (reference src/shaders.metal, Metal Shading Language = a C++14 dialect) is #included
// below, unmodified, from where it lies under /root/reference (REF_SHADER_PATH is set by oracle/Makefile); msl_shim.h
// supplies the slice of <metal_stdlib> it uses.  Built only where /root/reference exists, into oracle/_ref/ (git-ignored,
// travels to the GPU box as a binary).  TEST INFRASTRUCTURE: tests compare oracle/mm_oracle.cpp (and, on the GPU box, the
// CUDA kernel) with this, image for image.
//
// Dispatch emulation (reference src/main.rs:867-886: dispatch_thread_groups(grid, threads_per_threadgroup)): every Metal
// thread of a threadgroup is a fiber (ucontext) on one OS thread; threadgroup_barrier yields to the next fiber, so one
// round-robin pass brings every thread to the same barrier (the shader's barriers are unconditional).  `threadgroup`
// arrays are static thread_local: shared by the fibers of a group, private to the OS thread.  Threadgroups are
// independent and run in parallel over OpenMP threads.
//
// What the unmodified shader fixes: bounce_limit = 5, mirror_limit = 15 (shaders.metal:294-295) and the chunk lookup
// `tgid.x + tgid.y * ((width / 2) / ppc)` (:266), i.e. grid_x must equal (width / 2) / chunk_width^2 — true for the
// reference's own dispatch (1024 wide, chunk 4, 32 groups per row).  ref_compute_shader refuses other shapes.
#include <omp.h>
#include <ucontext.h>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "msl_shim.h"

// ---- fiber scheduler -------------------------------------------------------------------------------------------------
namespace {
struct Fiber {
    ucontext_t ctx;
    bool done = false;
};
struct GroupRun {
    ucontext_t main_ctx;
    std::vector<Fiber> fibers;
    std::vector<char> stacks;
    size_t current = 0;
};
thread_local GroupRun *g_run = nullptr;
}  // namespace

namespace metal {
void threadgroup_barrier(mem_flags) {
    GroupRun *r = g_run;
    swapcontext(&r->fibers[r->current].ctx, &r->main_ctx);      // back to the scheduler; resumed on the next pass
}
}  // namespace metal

// ---- the reference's shader, compiled as C++ -------------------------------------------------------------------------
#define thread
#define device
#define constant
#define threadgroup static thread_local
#define kernel
#define vertex
#define fragment
#define uint metal::U32
// Constructor calls become braced initialisations: the shader draws random numbers inside constructor arguments
// (`float3(random(state), random(state), random(state))`, shaders.metal:303,315-317).  The order of evaluation of
// parenthesised arguments is unspecified in C++; the Metal compiler (clang) evaluates them left to right, g++ right to
// left.  Braced initialiser lists are evaluated left to right by rule, so this reproduces clang's order without touching
// the reference source.  (Function-like macros: only `float3(` is rewritten, declarations such as `float3 v;` are not.)
#define float2(...) metal::float2{__VA_ARGS__}
#define float3(...) metal::float3{__VA_ARGS__}
#define float4(...) metal::float4{__VA_ARGS__}
#define uint2(...) metal::uint2{__VA_ARGS__}
#include REF_SHADER_PATH
#undef float2
#undef float3
#undef float4
#undef uint2
#undef uint
#undef thread
#undef device
#undef constant
#undef threadgroup
#undef kernel
#undef vertex
#undef fragment

// ---- dispatch ----------------------------------------------------------------------------------------------------------
namespace {
struct Dispatch {
    metal::texture2d<float, metal::access::read_write> texout;
    metal::texture2d<float, metal::access::sample> noise;
    const metal::uint2 *chunks;
    const rect *mirrors;
    const bvh_node *nodes;
    const metal::U32 *indices;
    const uni *uniforms;
    const bool *materials;
    const metal::float4 *emissions;
    uint32_t dim_x, dim_y;
};
struct ThreadArgs {
    const Dispatch *d;
    metal::uint2 tgid, gid, texid, dims;
};
thread_local std::vector<ThreadArgs> *g_args = nullptr;

void fiber_entry(int index) {
    const ThreadArgs &a = (*g_args)[(size_t)index];
    const Dispatch &d = *a.d;
    compute_shader(d.texout, d.noise, d.chunks, d.mirrors, d.nodes, d.indices, d.uniforms, d.materials, d.emissions,
                   a.tgid, a.gid, a.texid, a.dims);
    g_run->fibers[(size_t)index].done = true;
    swapcontext(&g_run->fibers[(size_t)index].ctx, &g_run->main_ctx);
}

void run_group(const Dispatch &d, uint32_t gx, uint32_t gy, GroupRun &run, std::vector<ThreadArgs> &args) {
    const size_t T = (size_t)d.dim_x * d.dim_y, kStack = 32 * 1024;
    run.fibers.assign(T, Fiber());
    if (run.stacks.size() < T * kStack) run.stacks.resize(T * kStack);
    args.resize(T);
    g_run = &run;
    g_args = &args;
    for (size_t i = 0; i < T; i++) {
        const uint32_t x = (uint32_t)(i % d.dim_x), y = (uint32_t)(i / d.dim_x);
        ThreadArgs &a = args[i];
        a.d = &d;
        a.tgid = metal::uint2(metal::U32(gx), metal::U32(gy));
        a.gid = metal::uint2(metal::U32(x), metal::U32(y));
        a.texid = metal::uint2(metal::U32(gx * d.dim_x + x), metal::U32(gy * d.dim_y + y));
        a.dims = metal::uint2(metal::U32(d.dim_x), metal::U32(d.dim_y));
        Fiber &f = run.fibers[i];
        getcontext(&f.ctx);
        f.ctx.uc_stack.ss_sp = run.stacks.data() + i * kStack;
        f.ctx.uc_stack.ss_size = kStack;
        f.ctx.uc_link = &run.main_ctx;
        makecontext(&f.ctx, (void (*)())fiber_entry, 1, (int)i);
    }
    for (bool any = true; any;) {            // one pass = every live thread runs to its next barrier (or to the end)
        any = false;
        for (size_t i = 0; i < T; i++) {
            if (run.fibers[i].done) continue;
            run.current = i;
            swapcontext(&run.main_ctx, &run.fibers[i].ctx);
            any = any || !run.fibers[i].done;
        }
    }
}
}  // namespace

extern "C" {

// Layout checks against the byte layouts of reference src/main.rs:32-90 (what the host uploads).
static_assert(sizeof(rect) == 48 && sizeof(bvh_node) == 32 && sizeof(camera) == 40 && sizeof(uni) == 56, "reference layouts");

// One dispatch of the reference's compute_shader.  out_rgba: view_height x view_width x 4 floats, only the pixels the
// dispatch writes are touched.  noise_rgba8: nw x nh RGBA8Unorm (converted to c / 255 as the texture unit does).
// Returns 0, or -1 when the shape is one the unmodified shader cannot address (see the header comment).
int ref_compute_shader(const void *planes, const void *nodes, const uint32_t *indices, const uint8_t *materials,
                       const float *emissions, const uint8_t *noise_rgba8, uint32_t nw, uint32_t nh, const void *uniform56,
                       const uint32_t *chunks_xy, uint32_t grid_x, uint32_t grid_y, uint32_t dim_x, uint32_t dim_y,
                       float *out_rgba, int threads) {
    const uni *u = static_cast<const uni *>(uniform56);
    const uint32_t chunk = u->chunk_width, ppc = chunk * chunk;
    const uint32_t T = dim_x * dim_y;
    if (ppc == 0 || T == 0 || T > 1024 || T % ppc != 0) return -1;                  // threadgroup float3 test[1024] (:343)
    if ((uint32_t)((u->view_width / 2.0f) / (float)ppc) != grid_x) return -1;        // :266
    std::vector<float> noise_f((size_t)nw * nh * 4);
    for (size_t i = 0; i < noise_f.size(); i++) noise_f[i] = (float)noise_rgba8[i] / 255.0f;
    Dispatch d;
    d.texout.rgba = out_rgba; d.texout.width = (unsigned)u->view_width; d.texout.height = (unsigned)u->view_height;
    d.noise.rgba = noise_f.data(); d.noise.width = nw; d.noise.height = nh;
    d.chunks = reinterpret_cast<const metal::uint2 *>(chunks_xy);
    d.mirrors = static_cast<const rect *>(planes);
    d.nodes = static_cast<const bvh_node *>(nodes);
    d.indices = reinterpret_cast<const metal::U32 *>(indices);
    d.uniforms = u;
    d.materials = reinterpret_cast<const bool *>(materials);
    d.emissions = reinterpret_cast<const metal::float4 *>(emissions);
    d.dim_x = dim_x; d.dim_y = dim_y;
    const int n_groups = (int)(grid_x * grid_y);
    if (threads <= 0) threads = omp_get_max_threads();
#pragma omp parallel num_threads(threads)
    {
        GroupRun run;
        std::vector<ThreadArgs> args;
#pragma omp for schedule(dynamic, 1)
        for (int g = 0; g < n_groups; g++) run_group(d, (uint32_t)g % grid_x, (uint32_t)g / grid_x, run, args);
    }
    return 0;
}

// Unit entry points onto the reference's helper functions (same signatures as the mmo_* hooks of mm_oracle.cpp).
float ref_random(uint32_t *state) {                                                           // shaders.metal:181-186
    metal::U32 st(*state);
    const float r = random(st);
    *state = st;
    return r;
}
float ref_intersect_aabb(const float *ori, const float *dir, float t, const float *bmin, const float *bmax) {   // :87-95
    ray beam;
    beam.ori = metal::float3(ori[0], ori[1], ori[2]); beam.dir = metal::float3(dir[0], dir[1], dir[2]); beam.t = t;
    return intersect_aabb(beam, metal::float3(bmin[0], bmin[1], bmin[2]), metal::float3(bmax[0], bmax[1], bmax[2]));
}
int ref_ray_rect(const float *ori, const float *dir, float t, const void *plane48, float *t_out) {               // :51-67
    ray beam;
    beam.ori = metal::float3(ori[0], ori[1], ori[2]); beam.dir = metal::float3(dir[0], dir[1], dir[2]); beam.t = t;
    beam.index = 0xFFFFFFFFu;
    ray_rect_intersect(beam, *static_cast<const rect *>(plane48), 7);
    *t_out = beam.t;
    return (unsigned int)beam.index == 7u ? 1 : 0;
}
void ref_quat_mult(const float *vec3, const float *quat4, float *out3) {                                         // :159-172
    const metal::float3 r = quat_mult(metal::float3(vec3[0], vec3[1], vec3[2]), metal::float4(quat4[0], quat4[1], quat4[2], quat4[3]));
    out3[0] = r.x; out3[1] = r.y; out3[2] = r.z;
}

// The reference's fragment_shader (shaders.metal:214-225) evaluated for every pixel of an image, each against the
// UNBLURRED image (the shader blurs in place, which makes a pass order-dependent; the product defines the pass as
// ping-pong).  The pixel's own texel is restored after each call; dst receives the returned colour.
void ref_fragment_blur(const float *src, uint32_t W, uint32_t H, float *dst) {
    std::vector<float> scratch(src, src + (size_t)W * H * 4);
    metal::texture2d<float, metal::access::read_write> image;
    image.rgba = scratch.data(); image.width = W; image.height = H;
    for (uint32_t y = 0; y < H; y++)
        for (uint32_t x = 0; x < W; x++) {
            ColorInOut in;
            in.position = metal::float4((float)x + 0.5f, (float)y + 0.5f, 0.0f, 1.0f);      // [[position]]: pixel centres
            float *texel = scratch.data() + 4 * ((size_t)y * W + x);
            const float keep[4] = {texel[0], texel[1], texel[2], texel[3]};
            const metal::float4 c = fragment_shader(in, image);
            std::memcpy(texel, keep, sizeof(keep));
            float *o = dst + 4 * ((size_t)y * W + x);
            o[0] = c.x; o[1] = c.y; o[2] = c.z; o[3] = c.w;
        }
}

int ref_shader_limits(int *bounce_limit, int *mirror_limit) {   // the literals of shaders.metal:294-295, for the tests' parameters
    *bounce_limit = 5; *mirror_limit = 15;
    return 0;
}

}  // extern "C"
