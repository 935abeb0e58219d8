// render_kernel.cu — the hot path: mirror-maze's per-pixel path-tracing kernel, hand-written for sm_100a.
//
// One CUDA thread = one Metal thread of the reference's `compute_shader` (reference src/shaders.metal:245-368):
// one (pixel, sample) path through ray generation (:281-284), seed + jitter (:291,298,303), the bounce loop
// (:306-340) around intersect_bvh_iterative (:115-156) with intersect_aabb (:87-95) and ray_rect_intersect
// (:51-67), per-sample tone-map (:344) and the per-pixel reduction in the reference's summation order (:347-366).
//
// Arithmetic contract: every fp32 + - * / sqrt is one IEEE round-to-nearest operation, written
// with explicit __f*_rn intrinsics so that no compiler flag can contract or approximate them.  Results are
// bit-identical to oracle/mm_oracle.cpp.
//
// Slab test without twelve divides.  The literal test divides (bound - origin) by the ray direction twelve times
// per interior node.  In the common case the same correctly rounded quotients are produced from one IEEE
// reciprocal per axis per ray: with r = RN(1/d), rho = 1 - d*r (exact by FMA), rl = RN(rho*r),
//     q1 = RN(x*r + RN(x*rl))      error < 1/2 ulp + 2^-23 ulp  (faithful)
//     e  = x - d*q1                exact by FMA
//     q  = RN(q1 + e*r)            == RN(x/d)   (Markstein's theorem: r = RN(1/d), q1 faithful, no over/underflow)
// The no-over/underflow side conditions are guaranteed by range checks: the scene's box coordinates are 0 or in
// [2^-10, 2^30] (checked at upload) and a ray uses this path only when every |d| is in [2^-60, 2^60] and every
// |origin| is 0 or in [2^-40, 2^30]; any other ray (zero / denormal / huge / NaN components) takes the literal
// __fdiv_rn traversal, so the union is exact for all inputs.  MM_FLAG_FORCE_LITERAL disables the fast path.
#include "trace_device.cuh"

namespace mmk {
namespace {

// Block epilogue shared by the trace kernels.  On entry red[] holds every thread's tone-mapped sample (3 planes of
// kBlockThreads floats, indexed by the thread the path started on) and the block has synchronised.
template <bool CNT, int kBlockThreads>
__device__ __forceinline__ void block_epilogue(const KParams &P, float *red, bool active, uint32_t flat, uint32_t pxx, uint32_t pxy, uint32_t k,
                                               uint32_t seg, uint32_t nhits, uint32_t nliteral, Tally tl) {
    // Per-pixel reduction in the reference's order (shaders.metal:343-366): pairs, quads, octets (the phases whose
    // stride is < spp), then the pixel's first thread adds the octets serially and divides by spp.
    const uint32_t tid = threadIdx.x;
    float *rx = red, *ry = red + kBlockThreads, *rz = red + 2 * kBlockThreads;
#pragma unroll
    for (uint32_t stride = 1; stride <= 4; stride *= 2) {
        if (stride < P.spp && (tid & (2 * stride - 1)) == 0) {
            rx[tid] = fadd(rx[tid], rx[tid + stride]);
            ry[tid] = fadd(ry[tid], ry[tid + stride]);
            rz[tid] = fadd(rz[tid], rz[tid + stride]);
        }
        __syncthreads();
    }
    if (active && (flat & (P.spp - 1)) == 0) {
        float sx = rx[tid], sy = ry[tid], sz = rz[tid];
        for (uint32_t i = 1; i < P.spp / 8; i++) {
            sx = fadd(sx, rx[tid + 8 * i]); sy = fadd(sy, ry[tid + 8 * i]); sz = fadd(sz, rz[tid + 8 * i]);
        }
        const float d = (float)(int)P.spp;
        float4 px = make_float4(fdiv(sx, d), fdiv(sy, d), fdiv(sz, d), 1.0f);
        if (P.quant8) px = quant8(px);
        if (P.image && pxx < P.W && pxy < P.H) reinterpret_cast<float4 *>(P.image)[(size_t)pxy * P.W + pxx] = px;
        if (P.tiles) reinterpret_cast<float4 *>(P.tiles)[(size_t)k * P.ppc + (flat >> P.log2_spp)] = px;
        if ((P.n_peers || P.host_out) && pxx < P.W && pxy < P.H) {
            const size_t at = (size_t)pxy * P.W + pxx;
            if (P.host_out) reinterpret_cast<float4 *>(P.host_out)[at] = px;        // zero-copy output: mapped pinned host frame, over PCIe
            if (P.peers_multicast) {                                    // fused exchange, NVSwitch multicast: one store, replicated by the switch
                float4 *mc = reinterpret_cast<float4 *>(P.peers[0]) + at;
                asm volatile("multimem.st.weak.global.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(mc), "f"(px.x), "f"(px.y), "f"(px.z), "f"(px.w) : "memory");
            } else {                                                    // fused exchange, NVLink peer mappings: one store per rank's frame
                for (uint32_t i = 0; i < P.n_peers; i++) reinterpret_cast<float4 *>(P.peers[i])[at] = px;
            }
        }
    }

    // Event counts: warp-reduce, one atomic per warp and counter.
    {
        unsigned long long v_rays = seg, v_hits = nhits, v_lit = nliteral, v_paths = active ? 1u : 0u;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            v_rays += __shfl_xor_sync(0xFFFFFFFFu, v_rays, o);
            v_hits += __shfl_xor_sync(0xFFFFFFFFu, v_hits, o);
            v_lit += __shfl_xor_sync(0xFFFFFFFFu, v_lit, o);
            v_paths += __shfl_xor_sync(0xFFFFFFFFu, v_paths, o);
        }
        unsigned long long v_inner = tl.inner, v_leaf = tl.leaf, v_rect = tl.rect;
        uint32_t v_ms = tl.max_stack;
        if (CNT) {
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                v_inner += __shfl_xor_sync(0xFFFFFFFFu, v_inner, o);
                v_leaf += __shfl_xor_sync(0xFFFFFFFFu, v_leaf, o);
                v_rect += __shfl_xor_sync(0xFFFFFFFFu, v_rect, o);
                v_ms = max(v_ms, __shfl_xor_sync(0xFFFFFFFFu, v_ms, o));
            }
        }
        if ((tid & 31u) == 0u) {
            atomicAdd(&P.counters->rays, v_rays);
            atomicAdd(&P.counters->hits, v_hits);
            atomicAdd(&P.counters->paths, v_paths);
            if (v_lit) atomicAdd(&P.counters->literal_rays, v_lit);
            if (CNT) {
                atomicAdd(&P.counters->inner_visits, v_inner);
                atomicAdd(&P.counters->leaf_visits, v_leaf);
                atomicAdd(&P.counters->rect_tests, v_rect);
                atomicMax(&P.counters->max_stack, (unsigned long long)v_ms);
            }
        }
    }
}

template <bool CNT, bool DBG, int kBlockThreads>
__global__ void __launch_bounds__(kBlockThreads, 1024 / kBlockThreads)
trace_kernel(const __grid_constant__ KParams P) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float *red = reinterpret_cast<float *>(smem_raw);                      // 3 * kBlockThreads floats

    const uint64_t path = (uint64_t)blockIdx.x * kBlockThreads + threadIdx.x;
    const bool active = path < P.total_paths;
    Tally tl = {0u, 0u, 0u, 0u};
    uint32_t seg = 0, nhits = 0, nliteral = 0;
    V3 sample = mk(0.0f, 0.0f, 0.0f);
    uint32_t pxx = 0, pxy = 0, k = 0, flat = 0;
    const uint32_t root = P.root_count ? (kLeafBit | P.root_link | (P.root_count << 24)) : P.root_link;   // pair 0 is at byte offset 0
    V3 st_ori = mk(0.0f, 0.0f, 0.0f), st_dir = mk(1.0f, 1.0f, 1.0f), st_color = mk(1.0f, 1.0f, 1.0f), st_light = mk(0.0f, 0.0f, 0.0f);
    float st_t = 1e30f;
    uint32_t st_slot = 0xFFFFFFFFu, st_state = 0u, first_hit = 0xFFFFFFFFu;
    int mirror_hits = 0;
    bool n_alive = false;

    if (active) {
        const PathStart ps = start_path(P, path);
        k = ps.k; flat = ps.flat; pxx = ps.pxx; pxy = ps.pxy;
        n_alive = 0 < P.bounce_limit;                                      // :306, n = 0
        st_ori = ps.ori; st_dir = ps.dir; st_state = ps.state;
    }

    // Bounce loop (shaders.metal:306-340).  The lanes of a warp go through it together, segment by segment, so that the
    // traversal's votes can use the full warp; a lane whose path has ended (or that has no path) idles with alive = false.
    {
        V3 ori = st_ori, dir = st_dir, color = st_color, light = st_light;
        float t = st_t;
        uint32_t slot = st_slot, state = st_state;
        int n = 0;
        bool alive = active && n_alive;
        while (__any_sync(0xFFFFFFFFu, alive)) {
            const bool lit = ray_is_literal(P, ori, dir);
            const bool any_lit = __any_sync(0xFFFFFFFFu, alive && lit) || !P.rect_fast_ok;   // MIXED also means literal rect tests
            Hit h;
            if (P.rcp_mode) {
                if (!any_lit) h = traverse<false, CNT, true>(P.pairs, P.rects, root, alive, false, ori, dir, t, slot, &tl);
                else h = traverse<true, CNT, true>(P.pairs, P.rects, root, alive, lit, ori, dir, t, slot, &tl);
            } else {
                if (!any_lit && P.rects_axis) h = traverse<false, CNT, false, true>(P.pairs, P.rects_axis, root, alive, false, ori, dir, t, slot, &tl);
                else if (!any_lit) h = traverse<false, CNT, false>(P.pairs, P.rects, root, alive, false, ori, dir, t, slot, &tl);
                else h = traverse<true, CNT, false>(P.pairs, P.rects, root, alive, lit, ori, dir, t, slot, &tl);
            }
            t = h.t; slot = h.slot;
            if (alive) {
                if (lit) nliteral++;
                seg++;
                if (!(t < 1e30f)) {                                        // :308, :336-339 (sky term is * 0.0)
                    alive = false;
                } else {
                    nhits++;
                    uint32_t orig = 0xFFFFFFFFu;
                    if (!shade_hit(P, slot, t, ori, dir, color, light, state, mirror_hits, (DBG && n == 0) ? &orig : nullptr)) alive = false;
                    if (DBG && n == 0) first_hit = orig;
                    t = 1e30f;                                             // :323, :330
                    n++;
                    alive = alive && (n < P.bounce_limit + mirror_hits);   // :306
                }
            }
        }
        sample = mk(fsqrt(fmaxf(light.x, 0.0f)), fsqrt(fmaxf(light.y, 0.0f)), fsqrt(fmaxf(light.z, 0.0f)));   // :344
        if (DBG && active) {
            if (P.dbg_first_hit) P.dbg_first_hit[path] = first_hit;
            if (P.dbg_segments) P.dbg_segments[path] = seg;
            if (P.dbg_mirror_hits) P.dbg_mirror_hits[path] = (uint32_t)mirror_hits;
            if (P.dbg_radiance) { P.dbg_radiance[3 * path] = light.x; P.dbg_radiance[3 * path + 1] = light.y; P.dbg_radiance[3 * path + 2] = light.z; }
        }
    }

    red[threadIdx.x] = sample.x; red[kBlockThreads + threadIdx.x] = sample.y; red[2 * kBlockThreads + threadIdx.x] = sample.z;
    __syncthreads();
    block_epilogue<CNT, kBlockThreads>(P, red, active, flat, pxx, pxy, k, seg, nhits, nliteral, tl);
}

// ---- trace_kernel_rg: the same kernel with the block's paths RE-FORMED INTO WARPS at every segment boundary ---------------------
// After the first diffuse bounce the rays of a warp have nothing in common, and the warp runs every segment for as long as
// its slowest ray needs (17.4 of 32 lanes busy in the interior body).  How long a ray travels — and with it how many nodes it
// visits — correlates with how steeply it points at the floor or the roof (|dir.y|: the maze is a 10-unit-high slab, rays near
// the vertical end after a few nodes, near-horizontal ones cross many cells; trace replays in tools/sched_sim_r2.py).  So at
// the end of every segment the block pushes the state of its live paths through shared memory, ordered by a 5-bit key on
// |dir.y| (a counting sort: one shared-memory atomic per path, one warp scan), and each warp continues with 32 paths of similar
// steepness; ended paths drop out, so the tail segments (paths kept alive by mirror hits) run in fewer, fuller warps.  Which lane
// traces which path changes, nothing else: every path keeps its own RNG state, counters and home thread (the thread it started
// on, where its sample is delivered for the pixel reduction), so every observable stays bit-identical.
template <bool CNT, bool DBG, int kBlockThreads>
__global__ void __launch_bounds__(kBlockThreads, 1024 / kBlockThreads)
trace_kernel_rg(const __grid_constant__ KParams P) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float *red = reinterpret_cast<float *>(smem_raw);                                  // 3 * kBlockThreads floats: samples by home thread
    float4 *xch = reinterpret_cast<float4 *>(smem_raw + 3 * kBlockThreads * sizeof(float));   // 4 planes of kBlockThreads float4: path state in flight
    uint32_t *hist = reinterpret_cast<uint32_t *>(xch + 4 * kBlockThreads);            // 2 x 32 bin counters (double-buffered)

    const uint32_t tid = threadIdx.x, lane = tid & 31u;
    const uint64_t path = (uint64_t)blockIdx.x * kBlockThreads + tid;
    const bool active = path < P.total_paths;
    Tally tl = {0u, 0u, 0u, 0u};
    uint32_t seg = 0, nhits = 0, nliteral = 0;                                         // per-lane event counts (summed over the grid at the end)
    uint32_t pxx = 0, pxy = 0, k = 0, flat = 0;                                        // of the path this thread STARTED (its home pixel)
    const uint32_t root = P.root_count ? (kLeafBit | P.root_link | (P.root_count << 24)) : P.root_link;
    // the path this lane currently carries
    V3 ori = mk(0.0f, 0.0f, 0.0f), dir = mk(1.0f, 1.0f, 1.0f), color = mk(1.0f, 1.0f, 1.0f), light = mk(0.0f, 0.0f, 0.0f);
    uint32_t state = 0u, first_hit = 0xFFFFFFFFu, home = tid, pseg = 0;
    int n = 0, mirror_hits = 0;
    bool has = false;

    if (active) {
        const PathStart ps = start_path(P, path);
        k = ps.k; flat = ps.flat; pxx = ps.pxx; pxy = ps.pxy;
        ori = ps.ori; dir = ps.dir; state = ps.state;
        has = 0 < P.bounce_limit;                                                      // :306, n = 0
    }
    red[tid] = 0.0f; red[kBlockThreads + tid] = 0.0f; red[2 * kBlockThreads + tid] = 0.0f;   // sqrt(max(0, 0)) of a path that never runs (:344)
    if (DBG && active && !has) {
        if (P.dbg_first_hit) P.dbg_first_hit[path] = 0xFFFFFFFFu;
        if (P.dbg_segments) P.dbg_segments[path] = 0u;
        if (P.dbg_mirror_hits) P.dbg_mirror_hits[path] = 0u;
        if (P.dbg_radiance) { P.dbg_radiance[3 * path] = 0.0f; P.dbg_radiance[3 * path + 1] = 0.0f; P.dbg_radiance[3 * path + 2] = 0.0f; }
    }
    if (tid < 64) hist[tid] = 0u;          // kBlockThreads >= 64
    uint32_t live = (uint32_t)__syncthreads_count(has);                                // paths of the block still running
    uint32_t round = 0, n_rg = 0;
    // P.rg_mask: bit r set = re-form the warps after segment r.  After the mask's last round the warps run on independently.
    const uint32_t last_rg = P.rg_mask ? 32u - (uint32_t)__clz(P.rg_mask) : 0u;        // first round index without any later regroup

    while (live > 0u) {
        // ---- one segment for the lanes that carry a path (whole warps without one skip it) ----
        if (__any_sync(0xFFFFFFFFu, has)) {
            const bool lit = ray_is_literal(P, ori, dir);
            const bool any_lit = __any_sync(0xFFFFFFFFu, has && lit) || !P.rect_fast_ok;
            Hit h;
            if (P.rcp_mode) {
                if (!any_lit) h = traverse<false, CNT, true>(P.pairs, P.rects, root, has, false, ori, dir, 1e30f, 0xFFFFFFFFu, &tl);
                else h = traverse<true, CNT, true>(P.pairs, P.rects, root, has, lit, ori, dir, 1e30f, 0xFFFFFFFFu, &tl);
            } else {
                if (!any_lit) h = traverse<false, CNT, false>(P.pairs, P.rects, root, has, false, ori, dir, 1e30f, 0xFFFFFFFFu, &tl);
                else h = traverse<true, CNT, false>(P.pairs, P.rects, root, has, lit, ori, dir, 1e30f, 0xFFFFFFFFu, &tl);
            }
            if (has) {
                if (lit) nliteral++;
                seg++; pseg++;
                bool alive = false;
                if (h.t < 1e30f) {                                                     // :308, :336-339 (sky term is * 0.0)
                    nhits++;
                    uint32_t orig = 0xFFFFFFFFu;
                    alive = shade_hit(P, h.slot, h.t, ori, dir, color, light, state, mirror_hits, (DBG && n == 0) ? &orig : nullptr);
                    if (DBG && n == 0) first_hit = orig;
                    n++;
                    alive = alive && (n < P.bounce_limit + mirror_hits);               // :306
                }
                if (!alive) {                                                          // the path is over: its sample goes home (:344)
                    red[home] = fsqrt(fmaxf(light.x, 0.0f)); red[kBlockThreads + home] = fsqrt(fmaxf(light.y, 0.0f));
                    red[2 * kBlockThreads + home] = fsqrt(fmaxf(light.z, 0.0f));
                    if (DBG) {
                        const uint64_t hp = (uint64_t)blockIdx.x * kBlockThreads + home;
                        if (P.dbg_first_hit) P.dbg_first_hit[hp] = first_hit;
                        if (P.dbg_segments) P.dbg_segments[hp] = pseg;
                        if (P.dbg_mirror_hits) P.dbg_mirror_hits[hp] = (uint32_t)mirror_hits;
                        if (P.dbg_radiance) { P.dbg_radiance[3 * hp] = light.x; P.dbg_radiance[3 * hp + 1] = light.y; P.dbg_radiance[3 * hp + 2] = light.z; }
                    }
                    has = false;
                }
            }
        }
        if (round >= 32u || !((P.rg_mask >> round) & 1u)) {                            // no regroup after this segment
            round++;
            if (round >= last_rg && !__any_sync(0xFFFFFFFFu, has)) break;              // nothing left to wait for: this warp is done
            continue;
        }
        // ---- re-form the warps: counting sort of the live paths by |dir.y| (dir is normalised: :321, :329) ----
        uint32_t *h_now = hist + 32u * (n_rg & 1u), *h_next = hist + 32u * ((n_rg + 1u) & 1u);
        n_rg++;
        const uint32_t bin = min(31u, (uint32_t)(fabsf(dir.y) * 32.0f));
        uint32_t pos = 0;
        if (has) pos = atomicAdd(h_now + bin, 1u);
        if (tid < 32) h_next[tid] = 0u;
        __syncthreads();
        const uint32_t c = h_now[lane];
        uint32_t incl = c;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t v = __shfl_up_sync(0xFFFFFFFFu, incl, o);
            if (lane >= (uint32_t)o) incl += v;
        }
        live = __shfl_sync(0xFFFFFFFFu, incl, 31);
        const uint32_t start = __shfl_sync(0xFFFFFFFFu, incl - c, bin);
        if (has) {
            const uint32_t r = start + pos;
            xch[r] = make_float4(ori.x, ori.y, ori.z, dir.x);
            xch[kBlockThreads + r] = make_float4(dir.y, dir.z, color.x, color.y);
            xch[2 * kBlockThreads + r] = make_float4(color.z, light.x, light.y, light.z);
            xch[3 * kBlockThreads + r] = make_float4(__uint_as_float(state), __uint_as_float((uint32_t)n | ((uint32_t)mirror_hits << 16)),
                                                     __uint_as_float(home | (pseg << 12)), __uint_as_float(first_hit));
        }
        __syncthreads();
        has = tid < live;
        if (has) {
            const float4 a = xch[tid], b = xch[kBlockThreads + tid], cc = xch[2 * kBlockThreads + tid], d = xch[3 * kBlockThreads + tid];
            ori = mk(a.x, a.y, a.z); dir = mk(a.w, b.x, b.y); color = mk(b.z, b.w, cc.x); light = mk(cc.y, cc.z, cc.w);
            state = __float_as_uint(d.x);
            const uint32_t nm = __float_as_uint(d.y), hs = __float_as_uint(d.z);
            n = (int)(nm & 0xFFFFu); mirror_hits = (int)(nm >> 16);
            home = hs & 0xFFFu; pseg = hs >> 12;
            first_hit = __float_as_uint(d.w);
        }
        round++;
    }
    __syncthreads();
    block_epilogue<CNT, kBlockThreads>(P, red, active, flat, pxx, pxy, k, seg, nhits, nliteral, tl);
}

// De-interleave gathered tiles into the frame (consumer side of the multi-GPU tile gather).
__global__ void scatter_kernel(const float4 *__restrict__ tiles, float4 *__restrict__ image, const mm_chunk *__restrict__ chunks,
                               uint32_t group_first, uint32_t group_step, uint32_t group_count, uint32_t chunk, uint32_t ppc,
                               uint32_t W, uint32_t H) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (uint64_t)group_count * ppc) return;
    const uint32_t k = (uint32_t)(i / ppc), pn = (uint32_t)(i % ppc);
    const mm_chunk ch = chunks[group_first + k * group_step];
    const uint32_t x = ch.x + pn / chunk, y = ch.y + pn % chunk;
    if (x < W && y < H) image[(size_t)y * W + x] = tiles[i];
}

// The whole all-gather result in one launch: rank r's tile k is group r + k*world (the interleaved partition).
__global__ void scatter_all_kernel(const float4 *__restrict__ gathered, float4 *__restrict__ image, const mm_chunk *__restrict__ chunks,
                                   uint32_t world, uint32_t max_count, uint32_t n_groups, uint32_t chunk, uint32_t ppc, uint32_t W, uint32_t H) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (uint64_t)world * max_count * ppc) return;
    const uint32_t pn = (uint32_t)(i % ppc);
    const uint64_t tile = i / ppc;
    const uint32_t r = (uint32_t)(tile / max_count), k = (uint32_t)(tile % max_count);
    const uint64_t g = (uint64_t)r + (uint64_t)k * world;
    if (g >= n_groups) return;                         // padding rows of ranks that own one group less
    const mm_chunk ch = chunks[g];
    const uint32_t x = ch.x + pn / chunk, y = ch.y + pn % chunk;
    if (x < W && y < H) image[(size_t)y * W + x] = gathered[i];
}

// Self-test of the shared-reciprocal quotient against __fdiv_rn on pseudo-random operands inside the guarded ranges
// (|d| in [2^-60, 2^60], x = 0 or |x| in [2^-40, 2^31]); half of the samples are built to land within a few
// 2^-24 ulp of a rounding midpoint, the only place where a faithful-but-not-exact quotient could differ.
__device__ __forceinline__ uint32_t mix32(uint64_t &s) {
    s = s * 6364136223846793005ull + 1442695040888963407ull;
    uint32_t x = (uint32_t)(((s >> 18) ^ s) >> 27), r = (uint32_t)(s >> 59);
    return (x >> r) | (x << ((32 - r) & 31));
}
__global__ void quot_selftest_kernel(uint64_t n, uint64_t seed, unsigned long long *mismatches) {
    const uint64_t tid = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x, stride = (uint64_t)gridDim.x * blockDim.x;
    uint64_t s = seed ^ (tid * 0x9E3779B97F4A7C15ull);
    unsigned long long bad = 0;
    for (uint64_t i = tid; i < n; i += stride) {
        const uint32_t a = mix32(s), b = mix32(s), c = mix32(s);
        // d: random sign/mantissa, exponent in [-60, 59]
        const int ed = (int)(a % 120u) - 60;
        float d = __uint_as_float((b & 0x807FFFFFu) | ((uint32_t)(ed + 127) << 23));
        float x;
        if (c & 1u) {
            // adversarial: x ~ (q + half ulp(q)) * d for a random q, so x/d sits next to a rounding midpoint
            const uint32_t qm = mix32(s);
            const int eq = (int)(qm % 40u) - 20;
            float q = __uint_as_float((mix32(s) & 0x807FFFFFu) | ((uint32_t)(eq + 127) << 23));
            double mid = (double)q + 0.5 * (double)(__uint_as_float(__float_as_uint(fabsf(q)) + 1u) - fabsf(q)) * (q < 0 ? -1.0 : 1.0);
            x = (float)(mid * (double)d);
            if ((c >> 1) & 1u) x = __uint_as_float(__float_as_uint(x) + ((c >> 2) & 3u) - 1u);   // +-1 ulp neighbours
        } else {
            const int ex = (int)((c >> 1) % 71u) - 40;
            x = __uint_as_float((mix32(s) & 0x807FFFFFu) | ((uint32_t)(ex + 127) << 23));
            if (((c >> 8) & 63u) == 0u) x = 0.0f;
        }
        const float ax = fabsf(x);
        if (!(ax == 0.0f || (ax >= 9.094947017729282e-13f && ax <= 2147483648.0f))) continue;
        Axis ax_;
        ax_.o = 0.0f; ax_.d = d;
        ax_.r = __frcp_rn(d); ax_.rl = fmul(__fmaf_rn(-d, ax_.r, 1.0f), ax_.r);
        const float qf = quot<true>(x, ax_), ql = quot<false>(x, ax_);
        const bool same = (__float_as_uint(qf) == __float_as_uint(ql)) || (qf == 0.0f && ql == 0.0f);
        if (!same) bad++;
    }
    if (bad) atomicAdd(mismatches, bad);
}

// fragment_shader's 5-tap blur (shaders.metal:214-225), ping-pong: one thread per pixel, one float4 per tap.  HBM-bound:
// 16 B read + 16 B written per pixel (the four neighbour taps hit L1/L2).
// x / 3 without the divide: q0 = RN(x * r), e = x - 3 q0 (exact by FMA), q = RN(q0 + e * r) with r = RN(1/3).  Equal to __fdiv_rn(x, 3)
// for every float whose quotient is a normal number — established by exhaustion over all 2^32 bit patterns (div3_selftest_kernel,
// mm_selftest_div3; tests/test_next_rows.py) — and guarded for the rest (tiny values take the IEEE divide).  x / 2 is x * 0.5: the
// same real number, hence the same rounding.
__device__ __forceinline__ float div3(float x) {
    if (!(fabsf(x) >= 1e-30f && fabsf(x) <= 3.4028234663852886e38f)) return fdiv(x, 3.0f);   // zeros, denormal quotients, infinities, NaN: the literal divide
    const float r = 0.3333333432674407958984375f;                          // RN(1/3)
    const float q0 = fmul(x, r);
    return __fmaf_rn(__fmaf_rn(-3.0f, q0, x), r, q0);
}
__global__ void div3_selftest_kernel(unsigned long long *mismatches) {
    const uint64_t tid = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x, stride = (uint64_t)gridDim.x * blockDim.x;
    unsigned long long bad = 0;
    for (uint64_t i = tid; i < (1ull << 32); i += stride) {
        const float x = __uint_as_float((uint32_t)i);
        const float a = div3(x), b = fdiv(x, 3.0f);
        if (__float_as_uint(a) != __float_as_uint(b) && !(a != a && b != b)) bad++;   // NaN payloads aside, bit for bit
    }
    if (bad) atomicAdd(mismatches, bad);
}

// fragment_shader's 5-tap blur (shaders.metal:214-225), ping-pong.  HBM-bound: 16 B read + 16 B written per pixel (the four
// neighbour taps hit L1/L2).  One thread handles two horizontally adjacent pixels (8 loads for 2 pixels instead of 10).
__global__ void __launch_bounds__(256) blur_kernel(const float4 *__restrict__ src, float4 *__restrict__ dst, uint32_t W, uint32_t H, bool q8,
                                                   uchar4 *__restrict__ bytes) {
    const uint32_t x0 = 2u * (blockIdx.x * blockDim.x + threadIdx.x), y = blockIdx.y;
    if (x0 >= W || y >= H) return;
    const float4 zero = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
    const size_t row = (size_t)y * W;
    const bool two = x0 + 1 < W;
    const float4 c0 = __ldg(src + row + x0), c1 = two ? __ldg(src + row + x0 + 1) : zero;
    const float4 l0 = x0 > 0 ? __ldg(src + row + x0 - 1) : zero, r1 = x0 + 2 < W ? __ldg(src + row + x0 + 2) : zero;
    const float4 d0 = y + 1 < H ? __ldg(src + row + W + x0) : zero, u0 = y > 0 ? __ldg(src + row - W + x0) : zero;
    const float4 d1 = (two && y + 1 < H) ? __ldg(src + row + W + x0 + 1) : zero, u1 = (two && y > 0) ? __ldg(src + row - W + x0 + 1) : zero;
    auto tap = [](float c, float r, float l, float d, float u) {             // :217-222, one channel
        return div3(fadd(fadd(c, fmul(fadd(r, l), 0.5f)), fmul(fadd(d, u), 0.5f)));
    };
    float4 o0 = make_float4(tap(c0.x, c1.x, l0.x, d0.x, u0.x), tap(c0.y, c1.y, l0.y, d0.y, u0.y), tap(c0.z, c1.z, l0.z, d0.z, u0.z), 1.0f);
    float4 o1 = make_float4(tap(c1.x, r1.x, c0.x, d1.x, u1.x), tap(c1.y, r1.y, c0.y, d1.y, u1.y), tap(c1.z, r1.z, c0.z, d1.z, u1.z), 1.0f);
    if (q8) { o0 = quant8(o0); o1 = quant8(o1); }
    dst[row + x0] = o0;
    if (two) dst[row + x0 + 1] = o1;
    if (bytes) {
        bytes[row + x0] = make_uchar4((unsigned char)unorm8(o0.x), (unsigned char)unorm8(o0.y), (unsigned char)unorm8(o0.z), (unsigned char)unorm8(o0.w));
        if (two) bytes[row + x0 + 1] = make_uchar4((unsigned char)unorm8(o1.x), (unsigned char)unorm8(o1.y), (unsigned char)unorm8(o1.z), (unsigned char)unorm8(o1.w));
    }
}

// Micro-benchmarks for the two rooflines the path is measured against (SURVEY §8 d): how fast can this GPU fetch 56 useful
// bytes (a 32-B, a 16-B and an 8-B load, the traversal's per-visit pattern) from random 128-B records of a table of the
// scene's size, and how fast can it issue FP32 FMAs.  Independent addresses / chains: these are peaks, not models.
__global__ void __launch_bounds__(256) mb_gather_kernel(const float4 *__restrict__ table, uint32_t n_records, uint32_t iters,
                                                        float *__restrict__ sink) {
    uint32_t idx = (blockIdx.x * blockDim.x + threadIdx.x) * 2654435761u;
    float acc = 0.0f;
    for (uint32_t i = 0; i < iters; i++) {
        idx = idx * 747796405u + 2891336453u;
        const float4 *rec = table + (size_t)((idx >> 8) % n_records) * 12;
        const uint32_t sxy = (idx & 3u) * 2u, sz = ((idx >> 2) & 1u) * 2u;      // any travel order, like the kernel
        const Line32 ab = ldg256(rec + sxy);
        const float4 z = __ldg(rec + 8 + sz);
        const float2 l = __ldg(reinterpret_cast<const float2 *>(rec + 9 + sz));
        acc += lo2f(ab.x) + hi2f(ab.y) + lo2f(ab.w) + z.z + l.x;
    }
    if (acc == 12345.678f) sink[0] = acc;                                       // keep the loads alive
}
__global__ void __launch_bounds__(256) mb_ffma_kernel(uint32_t iters, float *__restrict__ sink) {
    float a0 = threadIdx.x * 1e-3f, a1 = a0 + 1.0f, a2 = a0 + 2.0f, a3 = a0 + 3.0f, a4 = a0 + 4.0f, a5 = a0 + 5.0f, a6 = a0 + 6.0f, a7 = a0 + 7.0f;
    const float m = 1.000001f, c = 1e-7f;
    for (uint32_t i = 0; i < iters; i++) {
        a0 = __fmaf_rn(a0, m, c); a1 = __fmaf_rn(a1, m, c); a2 = __fmaf_rn(a2, m, c); a3 = __fmaf_rn(a3, m, c);
        a4 = __fmaf_rn(a4, m, c); a5 = __fmaf_rn(a5, m, c); a6 = __fmaf_rn(a6, m, c); a7 = __fmaf_rn(a7, m, c);
    }
    const float s = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
    if (s == 12345.678f) sink[0] = s;
}

template <bool C, bool D>
const void *kptr_rg(int bt) {
    if (bt == 64) return reinterpret_cast<const void *>(&trace_kernel_rg<C, D, 64>);
    if (bt == 128) return reinterpret_cast<const void *>(&trace_kernel_rg<C, D, 128>);
    return reinterpret_cast<const void *>(&trace_kernel_rg<C, D, kLargeBlock>);
}

template <bool C, bool D>
const void *kptr(int bt) {
    return bt == kSmallBlock ? reinterpret_cast<const void *>(&trace_kernel<C, D, kSmallBlock>)
                             : reinterpret_cast<const void *>(&trace_kernel<C, D, kLargeBlock>);
}

}  // namespace

const void *kernel_ptr(KernelChoice c) {
    if (c.regroup) {
        if (c.debug) return kptr_rg<true, true>(c.block_threads);
        return c.counters ? kptr_rg<true, false>(c.block_threads) : kptr_rg<false, false>(c.block_threads);
    }
    if (c.debug) return kptr<true, true>(c.block_threads);
    return c.counters ? kptr<true, false>(c.block_threads) : kptr<false, false>(c.block_threads);
}

cudaError_t launch_trace(const KParams &p, KernelChoice c, unsigned blocks, size_t smem_bytes, cudaStream_t stream) {
    const void *fn = kernel_ptr(c);
    void *args[] = {const_cast<KParams *>(&p)};
    return cudaLaunchKernel(fn, dim3(blocks), dim3(c.block_threads), args, smem_bytes, stream);
}

cudaError_t launch_mb_gather(const void *table, uint32_t n_records, uint32_t iters, unsigned blocks, float *sink, cudaStream_t stream) {
    mb_gather_kernel<<<blocks, 256, 0, stream>>>(reinterpret_cast<const float4 *>(table), n_records, iters, sink);
    return cudaGetLastError();
}
cudaError_t launch_mb_ffma(uint32_t iters, unsigned blocks, float *sink, cudaStream_t stream) {
    mb_ffma_kernel<<<blocks, 256, 0, stream>>>(iters, sink);
    return cudaGetLastError();
}

cudaError_t launch_scatter_all(const float *gathered, float *image, const mm_chunk *chunks, uint32_t world, uint32_t max_count,
                               uint32_t n_groups, uint32_t chunk, uint32_t W, uint32_t H, cudaStream_t stream) {
    const uint32_t ppc = chunk * chunk;
    const uint64_t n = (uint64_t)world * max_count * ppc;
    if (n == 0) return cudaSuccess;
    scatter_all_kernel<<<(unsigned)((n + 255) / 256), 256, 0, stream>>>(reinterpret_cast<const float4 *>(gathered),
                                                                       reinterpret_cast<float4 *>(image), chunks, world, max_count,
                                                                       n_groups, chunk, ppc, W, H);
    return cudaGetLastError();
}

cudaError_t launch_blur(const float *src, float *dst, uint32_t W, uint32_t H, cudaStream_t stream, bool quant8, uint8_t *bytes) {
    if (W == 0 || H == 0) return cudaSuccess;
    dim3 grid((W + 511) / 512, H);                                            // two pixels per thread
    blur_kernel<<<grid, 256, 0, stream>>>(reinterpret_cast<const float4 *>(src), reinterpret_cast<float4 *>(dst), W, H, quant8,
                                          reinterpret_cast<uchar4 *>(bytes));
    return cudaGetLastError();
}

cudaError_t launch_div3_selftest(unsigned long long *d_mismatches, cudaStream_t stream) {
    div3_selftest_kernel<<<148 * 16, 256, 0, stream>>>(d_mismatches);
    return cudaGetLastError();
}

cudaError_t launch_quot_selftest(uint64_t n, uint64_t seed, unsigned long long *d_mismatches, cudaStream_t stream) {
    quot_selftest_kernel<<<148 * 8, 256, 0, stream>>>(n, seed, d_mismatches);
    return cudaGetLastError();
}

cudaError_t launch_scatter(const float *tiles, float *image, const mm_chunk *chunks, uint32_t /*grid_groups*/, uint32_t group_first,
                           uint32_t group_step, uint32_t group_count, uint32_t chunk, uint32_t W, uint32_t H, cudaStream_t stream) {
    const uint32_t ppc = chunk * chunk;
    const uint64_t n = (uint64_t)group_count * ppc;
    if (n == 0) return cudaSuccess;
    const unsigned blocks = (unsigned)((n + 255) / 256);
    scatter_kernel<<<blocks, 256, 0, stream>>>(reinterpret_cast<const float4 *>(tiles), reinterpret_cast<float4 *>(image), chunks,
                                               group_first, group_step, group_count, chunk, ppc, W, H);
    return cudaGetLastError();
}

}  // namespace mmk
