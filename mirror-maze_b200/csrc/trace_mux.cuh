// trace_mux.cuh — included by render_kernel.cu inside namespace mmk::{anonymous}, after the step functions.
//
// trace_kernel_mux<K>: the same path (reference src/shaders.metal:245-368) with K rays per lane.  Every thread owns K
// paths (slice k of a block = paths block_base + k*256 + tid, so each slice keeps the reference's thread -> pixel/sample
// mapping) whose ray and path state lives in shared memory (96 B per path, SoA float4 arrays: a lane always hits its
// own bank group).  At every vote a lane offers whichever of its rays stands at an interior node, so the interior body
// runs with more lanes than with one ray per lane (oracle-trace replay, tools/sched_sim.py: 17 -> 21-22 of 32 lanes).
// Per-ray arithmetic, visit order and RNG consumption are untouched: results are bit-identical to trace_kernel.
//
// Descriptor bits here: 0..23 link, 24..28 leaf count (<= 30, checked on the host), 29..31 travel octant of the ray
// (bit a set = the ray goes down axis a), so that the pair addresses of a visit come from the register alone.

constexpr uint32_t kMuxCountMask = 0x1Fu;

template <int K>
struct MuxState {
    float4 *A, *B, *C, *D, *E, *F;   // [K * kBlockThreads] each
    // A = (ori.xyz, t)   B = (dir.xyz, slot)   C = (r.xyz, head | lit << 31)   D = (rl.xyz, -)
    // E = (throughput.rgb, rng state)   F = (radiance.rgb, mirror_hits | n << 8 | seg << 16 | nhits << 24)
    __device__ __forceinline__ explicit MuxState(unsigned char *base) {
        A = reinterpret_cast<float4 *>(base);
        B = A + K * kBlockThreads; C = B + K * kBlockThreads; D = C + K * kBlockThreads;
        E = D + K * kBlockThreads; F = E + K * kBlockThreads;
    }
};

template <int K, bool CNT, bool DBG>
__global__ void __launch_bounds__(kBlockThreads, (K <= 2 ? 4 : 3))
trace_kernel_mux(const __grid_constant__ KParams P) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    MuxState<K> S(smem_raw);
    const uint32_t tid = threadIdx.x;
    const uint64_t block_base = (uint64_t)blockIdx.x * (uint64_t)(K * kBlockThreads);
    const uint32_t root = P.root_link | (P.root_count << 24);
    const char *pair_base = reinterpret_cast<const char *>(P.pairs);

    uint32_t cur[K];
    uint32_t stack[K][MM_MAX_STACK];
    Tally tl = {0u, 0u, 0u, 0u};
    uint32_t nliteral = 0, alive = 0;

    // ---- ray generation for the K paths of this thread (shaders.metal:261-304) ------------------------------------
#pragma unroll
    for (int k = 0; k < K; k++) {
        const uint32_t r = k * kBlockThreads + tid;
        const uint64_t path = block_base + r;
        cur[k] = CUR_END;
        if (path < P.total_paths) {
            const uint32_t kg = (uint32_t)(path / P.T);
            const uint32_t flat = (uint32_t)(path - (uint64_t)kg * P.T);
            const uint32_t g = P.group_first + kg * P.group_step;
            const uint32_t tgx = g % P.grid_x, tgy = g / P.grid_x;
            const mm_chunk ch = P.chunks[g];                                   // :266-267
            const uint32_t gx = flat % P.dim_x, gy = flat / P.dim_x;           // inverse of :271
            const uint32_t chunk = P.uni.chunk_width;
            const uint32_t pixel_number = flat >> P.log2_spp;                  // :272
            const uint32_t pxx = ch.x + pixel_number / chunk, pxy = ch.y + pixel_number % chunk;   // :273-275
            const uint32_t texid_x = tgx * P.dim_x + gx, texid_y = tgy * P.dim_y + gy;
            const mm_camera &cam = P.uni.cam;
            const V3 center = mk(cam.camera_center.x, cam.camera_center.y, cam.camera_center.z);
            const float pnx = fdiv(__uint2float_rn(pxx), P.uni.view_width), pny = fdiv(__uint2float_rn(pxy), P.uni.view_height);   // :281
            const V3 corner = sub3(center, mk(fdiv(cam.viewport.x, 2.0f), fdiv(cam.viewport.y, 2.0f), -cam.focal_length));        // :282
            V3 ray_dir = normalize3(sub3(add3(corner, mk(fmul(pnx, cam.viewport.x), fmul(pny, cam.viewport.y), 0.0f)), center));  // :283
            const Q4 rot = {cam.rotation.x, cam.rotation.y, cam.rotation.z, cam.rotation.w};
            ray_dir = quat_mult(ray_dir, rot);                                 // :284
            float nx, ny;
            sample_noise_xy(P.noise, P.noise_w, P.noise_h, __uint2float_rn(gx), __uint2float_rn(gy), nx, ny);   // :291
            const float seed_f = fadd(fadd(fadd(fadd(nx, ny), __uint2float_rn(texid_x * 15823u)), __uint2float_rn(texid_y * 9737333u)),
                                      __uint2float_rn(P.uni.time));           // :298
            uint32_t state = __float2uint_rz(seed_f);
            const float j1 = rnd_pm1(state), j2 = rnd_pm1(state);
            const V3 dir = add3(ray_dir, scale3(mk(j1, j2, 0.0f), 0.001f));    // :303
            S.A[r] = make_float4(center.x, center.y, center.z, 1e30f);         // :302, :304
            S.B[r] = make_float4(dir.x, dir.y, dir.z, __uint_as_float(0xFFFFFFFFu));
            S.E[r] = make_float4(1.0f, 1.0f, 1.0f, __uint_as_float(state));    // :289
            S.F[r] = make_float4(0.0f, 0.0f, 0.0f, __uint_as_float(0u));       // :290, mirror_hits = n = 0
            if (0 < P.bounce_limit) alive |= 1u << k;                          // :306, n = 0
            if (DBG && P.dbg_first_hit) P.dbg_first_hit[path] = 0xFFFFFFFFu;
        } else {
            S.F[r] = make_float4(0.0f, 0.0f, 0.0f, __uint_as_float(0u));
        }
    }

    // ---- bounce loop: the warp goes through it segment by segment, every lane with up to K live rays -----------------
    while (__any_sync(0xFFFFFFFFu, alive != 0u)) {
        uint32_t litmask = 0;
#pragma unroll
        for (int k = 0; k < K; k++) {
            cur[k] = CUR_END;
            if (alive & (1u << k)) {
                const uint32_t r = k * kBlockThreads + tid;
                const float4 a = S.A[r], b = S.B[r];
                const bool lit = P.force_literal || !P.scene_fast_ok || !(axis_safe(a.x, b.x) && axis_safe(a.y, b.y) && axis_safe(a.z, b.z));
                const float rx = __frcp_rn(b.x), ry = __frcp_rn(b.y), rz = __frcp_rn(b.z);
                S.C[r] = make_float4(rx, ry, rz, __uint_as_float(lit ? 0x80000000u : 0u));      // head = 0
                S.D[r] = make_float4(fmul(__fmaf_rn(-b.x, rx, 1.0f), rx), fmul(__fmaf_rn(-b.y, ry, 1.0f), ry),
                                     fmul(__fmaf_rn(-b.z, rz, 1.0f), rz), 0.0f);
                const uint32_t oct = lit ? 0u : ((b.x < 0.0f ? 1u : 0u) | (b.y < 0.0f ? 2u : 0u) | (b.z < 0.0f ? 4u : 0u));
                cur[k] = root | (oct << 29);
                if (lit) { litmask |= 1u << k; nliteral++; }
            }
        }
        const bool mixed = __any_sync(0xFFFFFFFFu, litmask != 0u);

        // traversal of all live rays of the warp (shaders.metal:115-156 per ray)
        while (true) {
            int kI = -1, kL = -1;
#pragma unroll
            for (int k = K - 1; k >= 0; k--) {
                const uint32_t c = cur[k];
                const bool isI = ((c >> 24) & kMuxCountMask) == 0u;
                if (isI) kI = k;
                else if (c != CUR_END) kL = k;
            }
            const unsigned mI = __ballot_sync(0xFFFFFFFFu, kI >= 0), mL = __ballot_sync(0xFFFFFFFFu, kL >= 0);
            if ((mI | mL) == 0u) break;
            const unsigned mLonly = mL & ~mI;
            if (mI != 0u && __popc(mI) >= kLeafWeight * __popc(mLonly)) {
#pragma unroll 1
                for (uint32_t rep = 0; rep < kInnerReps; rep++) {
                    int k = -1;
#pragma unroll
                    for (int q = K - 1; q >= 0; q--)
                        if (((cur[q] >> 24) & kMuxCountMask) == 0u) k = q;
                    if (k >= 0) {
                        uint32_t c = cur[0];
#pragma unroll
                        for (int q = 1; q < K; q++) c = (k == q) ? cur[q] : c;
                        const uint32_t r = k * kBlockThreads + tid;
                        // pair record: travel order from the descriptor's octant bits
                        const char *rec = pair_base + (size_t)(c & 0xFFFFFFu) * sizeof(PairRec);
                        const float4 bx = __ldg(reinterpret_cast<const float4 *>(rec + ((c >> 23) & 64u)));
                        const float4 by = __ldg(reinterpret_cast<const float4 *>(rec + 16 + ((c >> 24) & 64u)));
                        const float4 bz = __ldg(reinterpret_cast<const float4 *>(rec + 32 + ((c >> 25) & 64u)));
                        const uint2 lk = __ldg(reinterpret_cast<const uint2 *>(rec + 48));
                        const float4 a = S.A[r], b = S.B[r], cc = S.C[r], dd = S.D[r];
                        Axis ax, ay, az;
                        ax.o = a.x; ax.d = b.x; ax.r = cc.x; ax.rl = dd.x;
                        ay.o = a.y; ay.d = b.y; ay.r = cc.y; ay.rl = dd.y;
                        az.o = a.z; az.d = b.z; az.r = cc.z; az.rl = dd.z;
                        uint32_t hw = __float_as_uint(cc.w), head = hw & 0x7FFFFFFFu, nc = c;
                        if (CNT) tl.inner++;
                        if (!mixed || !(hw >> 31)) inner_step<true, CNT>(bx, by, bz, lk, ax, ay, az, a.w, nc, head, stack[k], tl);
                        else inner_step<false, CNT>(bx, by, bz, lk, ax, ay, az, a.w, nc, head, stack[k], tl);
                        S.C[r].w = __uint_as_float(head | (hw & 0x80000000u));
                        nc |= c & 0xE0000000u;                                  // keep the octant (CUR_END stays all ones)
#pragma unroll
                        for (int q = 0; q < K; q++)
                            if (k == q) cur[q] = nc;
                    }
                }
            } else {
                if (kL >= 0) {
                    uint32_t c = cur[0];
#pragma unroll
                    for (int q = 1; q < K; q++) c = (kL == q) ? cur[q] : c;
                    const uint32_t r = kL * kBlockThreads + tid;
                    const float4 a = S.A[r], b = S.B[r];
                    const uint32_t hw = __float_as_uint(S.C[r].w);
                    float t = a.w;
                    uint32_t slot = __float_as_uint(b.w), head = hw & 0x7FFFFFFFu, nc = c & 0x1FFFFFFFu;
                    if (CNT) tl.leaf++;
                    leaf_step<CNT>(P.rects, mk(a.x, a.y, a.z), mk(b.x, b.y, b.z), t, slot, nc, head, stack[kL], tl);
                    S.A[r].w = t;
                    S.B[r].w = __uint_as_float(slot);
                    S.C[r].w = __uint_as_float(head | (hw & 0x80000000u));
                    nc |= c & 0xE0000000u;
#pragma unroll
                    for (int q = 0; q < K; q++)
                        if (kL == q) cur[q] = nc;
                }
            }
        }

        // shading of the finished segment (shaders.metal:308-339), ray by ray
#pragma unroll 1
        for (int k = 0; k < K; k++) {
            if (!(alive & (1u << k))) continue;
            const uint32_t r = k * kBlockThreads + tid;
            const float4 a = S.A[r], b = S.B[r];
            float4 e = S.E[r], f = S.F[r];
            uint32_t pk = __float_as_uint(f.w);
            int mirror_hits = (int)(pk & 0xFFu), n = (int)((pk >> 8) & 0xFFu);
            uint32_t seg = (pk >> 16) & 0xFFu, nhits = pk >> 24;
            const float t = a.w;
            const uint32_t slot = __float_as_uint(b.w);
            const V3 ori = mk(a.x, a.y, a.z), dir = mk(b.x, b.y, b.z);
            bool live = true;
            seg++;
            if (!(t < 1e30f)) {                                            // :308, :336-339 (sky term is * 0.0)
                live = false;
            } else {
                nhits++;
                const float4 *rp = reinterpret_cast<const float4 *>(P.rects + slot);
                const float4 r1 = __ldg(rp + 1), r2 = __ldg(rp + 2), r3 = __ldg(rp + 3);
                if (DBG && n == 0 && P.dbg_first_hit) P.dbg_first_hit[block_base + r] = __float_as_uint(r2.w);
                const V3 nrm = mk(r1.x, r1.y, r1.z);                       // :309
                const float side = -sign1(dot3(dir, nrm));                 // :310
                const float4 *sp4 = reinterpret_cast<const float4 *>(P.shade + slot);
                if (__float_as_uint(r3.w) == 0u || side == -1.0f) {        // :311
                    const float4 col = __ldg(sp4), emi = __ldg(sp4 + 1);
                    const V3 color = mk(e.x, e.y, e.z);
                    const V3 light = add3(mk(f.x, f.y, f.z), mul3(mk(emi.x, emi.y, emi.z), color));   // :312-313
                    const V3 c2 = mul3(color, mk(col.x, col.y, col.z));    // :314
                    uint32_t state = __float_as_uint(e.w);
                    V3 rd;
                    do {                                                   // :315-318
                        const float x = rnd_pm1(state), y = rnd_pm1(state), z = rnd_pm1(state);
                        rd = mk(x, y, z);
                    } while (length3(rd) > 1.0f);
                    rd = normalize3(rd);                                   // :319
                    const V3 o2 = add3(ori, scale3(dir, t));               // :320
                    const V3 d2 = normalize3(add3(rd, scale3(nrm, side))); // :321
                    e = make_float4(c2.x, c2.y, c2.z, __uint_as_float(state));
                    f.x = light.x; f.y = light.y; f.z = light.z;
                    S.A[r] = make_float4(o2.x, o2.y, o2.z, 1e30f);         // :323
                    S.B[r] = make_float4(d2.x, d2.y, d2.z, b.w);
                } else {
                    mirror_hits++;                                         // :325
                    if (mirror_hits < P.mirror_limit) {                    // :326
                        const float4 col = __ldg(sp4);
                        const V3 light = add3(mk(f.x, f.y, f.z), scale3(mk(col.x, col.y, col.z), 0.005f));   // :327
                        const V3 o2 = add3(ori, scale3(dir, t));           // :328
                        const V3 d2 = normalize3(reflect3(dir, nrm));      // :329
                        f.x = light.x; f.y = light.y; f.z = light.z;
                        S.A[r] = make_float4(o2.x, o2.y, o2.z, 1e30f);     // :330
                        S.B[r] = make_float4(d2.x, d2.y, d2.z, b.w);
                    } else {
                        live = false;                                      // :333
                    }
                }
                n++;
                live = live && (n < P.bounce_limit + mirror_hits);         // :306
            }
            f.w = __uint_as_float((uint32_t)mirror_hits | ((uint32_t)n << 8) | (seg << 16) | (nhits << 24));
            S.E[r] = e;
            S.F[r] = f;
            if (!live) alive &= ~(1u << k);
        }
    }

    // ---- tone-map, per-pixel reduction in the reference's order, store (shaders.metal:343-366), slice by slice ------------
    unsigned long long v_rays = 0, v_hits = 0, v_paths = 0;
    float *red = reinterpret_cast<float *>(S.A);                            // ray state is dead now: reuse it as scratch
    float *rx = red, *ry = red + kBlockThreads, *rz = red + 2 * kBlockThreads;
    __syncthreads();
#pragma unroll 1
    for (int k = 0; k < K; k++) {
        const uint32_t r = k * kBlockThreads + tid;
        const uint64_t path = block_base + r;
        const bool active = path < P.total_paths;
        const float4 f = S.F[r];
        const uint32_t pk = __float_as_uint(f.w);
        if (active) {
            v_paths++; v_rays += (pk >> 16) & 0xFFu; v_hits += pk >> 24;
            if (DBG) {
                if (P.dbg_segments) P.dbg_segments[path] = (pk >> 16) & 0xFFu;
                if (P.dbg_mirror_hits) P.dbg_mirror_hits[path] = pk & 0xFFu;
                if (P.dbg_radiance) { P.dbg_radiance[3 * path] = f.x; P.dbg_radiance[3 * path + 1] = f.y; P.dbg_radiance[3 * path + 2] = f.z; }
            }
        }
        rx[tid] = fsqrt(fmaxf(f.x, 0.0f)); ry[tid] = fsqrt(fmaxf(f.y, 0.0f)); rz[tid] = fsqrt(fmaxf(f.z, 0.0f));   // :344
        __syncthreads();
#pragma unroll
        for (uint32_t stride = 1; stride <= 4; stride *= 2) {
            if (stride < P.spp && (tid & (2 * stride - 1)) == 0) {
                rx[tid] = fadd(rx[tid], rx[tid + stride]);
                ry[tid] = fadd(ry[tid], ry[tid + stride]);
                rz[tid] = fadd(rz[tid], rz[tid + stride]);
            }
            __syncthreads();
        }
        if (active) {
            const uint32_t kg = (uint32_t)(path / P.T);
            const uint32_t flat = (uint32_t)(path - (uint64_t)kg * P.T);
            if ((flat & (P.spp - 1)) == 0) {
                float sx = rx[tid], sy = ry[tid], sz = rz[tid];
                for (uint32_t i = 1; i < P.spp / 8; i++) {
                    sx = fadd(sx, rx[tid + 8 * i]); sy = fadd(sy, ry[tid + 8 * i]); sz = fadd(sz, rz[tid + 8 * i]);
                }
                const float d = (float)(int)P.spp;
                const float4 px = make_float4(fdiv(sx, d), fdiv(sy, d), fdiv(sz, d), 1.0f);
                const uint32_t g = P.group_first + kg * P.group_step;
                const mm_chunk ch = P.chunks[g];
                const uint32_t chunk = P.uni.chunk_width, pixel_number = flat >> P.log2_spp;
                const uint32_t pxx = ch.x + pixel_number / chunk, pxy = ch.y + pixel_number % chunk;
                if (P.image && pxx < P.W && pxy < P.H) reinterpret_cast<float4 *>(P.image)[(size_t)pxy * P.W + pxx] = px;
                if (P.tiles) reinterpret_cast<float4 *>(P.tiles)[(size_t)kg * P.ppc + pixel_number] = px;
            }
        }
        __syncthreads();
    }

    // event counts
    {
        unsigned long long v_lit = nliteral, v_inner = tl.inner, v_leaf = tl.leaf, v_rect = tl.rect;
        uint32_t v_ms = tl.max_stack;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            v_rays += __shfl_xor_sync(0xFFFFFFFFu, v_rays, o);
            v_hits += __shfl_xor_sync(0xFFFFFFFFu, v_hits, o);
            v_lit += __shfl_xor_sync(0xFFFFFFFFu, v_lit, o);
            v_paths += __shfl_xor_sync(0xFFFFFFFFu, v_paths, o);
            if (CNT) {
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
