// pool_kernel.cu — the hot path as persistent warps over a ray pool in shared memory (sm_100a).
//
// Same per-path arithmetic as render_kernel.cu (both call trace_device.cuh: reference src/shaders.metal:245-368 under the
// SURVEY 8 a-0 contract), different scheduling.  render_kernel.cu binds one path to one lane for its whole life, so a lane
// idles whenever its ray is in another phase than the warp's majority: ncu shows 17.4 of 32 lanes active in the interior
// body, 9.8 in the leaf body, 9.7 in the shading's rejection loop.  Here a warp owns a POOL of M paths whose state lives in
// shared memory (one record per path: ray constants, traversal stack, path state), and runs one of four bodies at a time on
// up to 32 paths that are all READY for that body:
//   G  generate 32 new paths of the virtual dispatch into free records                 (shaders.metal:261-303)
//   I  four interior visits of intersect_bvh_iterative for 32 rays standing at interior nodes  (:131-154)
//   L  one leaf visit (rect tests) for 32 rays standing at a leaf                       (:126-130, :51-67)
//   S  shade 32 finished traversals, set up the next segment or finish the path         (:308-344)
// Ready paths wait in per-body ring queues; every body pops its lanes from one queue and pushes each path into the queue of
// its next state.  Each ray still performs exactly the reference's visit sequence — only which lane executes which visit,
// and when, changes — so every observable and every counter stays bit-identical (tests/test_gpu_parity.py).  Finished samples
// are collected per pixel in shared memory and reduced in the reference's summation order (:347-364) by the lane that
// delivers a pixel's last sample.  Rays whose operands fall outside the guarded ranges of the exact shared-reciprocal slab
// quotient (about 49 per 268 M) run the literal-divide traversal of trace_device.cuh to completion in a fifth, rare body X.
//
// Work distribution: persistent grid (blocks = SMs x resident blocks), units of max(32, spp) consecutive paths (whole
// pixels) handed out by a global atomic counter that is prefetched one unit ahead.
#include "trace_device.cuh"

namespace mmk {
namespace {

// ---- record layout (32-bit words; stride = P.pool_slot_words, a multiple of 4) ------------------------------------------------
//  0..3   -o.x -o.y -d.x -d.y        one 16-B load -> the packed (x, y) constants of the slab test
//  4..7    r.x  r.y -o.z -d.z        r = RN(1/d)
//  8..11   r.z  t    cur  hit        t = beam.t, cur = node descriptor (CUR_END when the traversal is finished), hit = rect slot
// 12       sp | prec << 8 | sample << 16      stack depth, pixel record, sample index inside the pixel
// 13       RNG state     14  n | mirror_hits << 16     15  path index
// 16..18   throughput ("color")      19..21  radiance ("light")      22  first-hit id (DBG)      23  segments
// 24..     traversal stack, entry 0 = CUR_END sentinel
constexpr uint32_t W_T = 9, W_CUR = 10, W_HIT = 11, W_SP = 12, W_STACK = 24;
constexpr uint32_t kRing = 128;                  // ring capacity (record ids are bytes; M <= 128)
constexpr uint32_t kPrecHdr = 4;                 // pixel record header: samples done, x, y, tile index

struct Ring {
    uint32_t head = 0, cnt = 0;
    uint8_t *buf;
    __device__ __forceinline__ uint32_t pop(uint32_t n, uint32_t lane) {            // n <= cnt, uniform
        const uint32_t id = buf[(head + lane) & (kRing - 1)];
        head += n; cnt -= n;
        return id;
    }
    __device__ __forceinline__ void push(bool pred, uint32_t id, uint32_t lane_lt) {
        const unsigned m = __ballot_sync(0xFFFFFFFFu, pred);
        if (pred) buf[(head + cnt + __popc(m & lane_lt)) & (kRing - 1)] = (uint8_t)id;
        cnt += __popc(m);
    }
};

template <bool CNT, bool DBG>
__global__ void __launch_bounds__(128) pool_kernel(const __grid_constant__ KParams P) {
    extern __shared__ __align__(16) uint32_t smem_pool[];
    const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5, lane_lt = (1u << lane) - 1u;
    uint32_t *const region = smem_pool + warp * P.pool_region_words;
    uint32_t *const slots = region;                                                  // M records
    const uint32_t SW = P.pool_slot_words, M = P.pool_M, spp = P.spp;
    uint8_t *const rings = reinterpret_cast<uint8_t *>(region + M * SW);             // 5 rings of kRing bytes
    uint32_t *const precs = region + M * SW + 5 * kRing / 4;                         // R pixel records
    const uint32_t PW = kPrecHdr + 3 * spp, R = P.pool_R;
    Ring qI, qL, qS, qX, qF;
    qI.buf = rings; qL.buf = rings + kRing; qS.buf = rings + 2 * kRing; qX.buf = rings + 3 * kRing; qF.buf = rings + 4 * kRing;
    for (uint32_t i = lane; i < M; i += 32) qF.buf[i] = (uint8_t)i;
    qF.cnt = M;
    for (uint32_t i = lane; i < R; i += 32) precs[i * PW] = 0xFFFFFFFFu;               // free
    __syncwarp();

    const uint32_t root = P.root_count ? (kLeafBit | P.root_link | (P.root_count << 24)) : P.root_link;
    const uint32_t unit_paths = spp > 32u ? spp : 32u, gens_per_unit = unit_paths >> 5;
    const uint32_t n_units = (uint32_t)((P.total_paths + unit_paths - 1) / unit_paths);
    unsigned long long *const ticket = &P.counters->next_unit;
    // unit prefetch: `next` is the unit this warp works on after the current one
    uint32_t next = 0;
    if (lane == 0) next = (uint32_t)atomicAdd(ticket, 1ull);
    next = __shfl_sync(0xFFFFFFFFu, next, 0);
    uint32_t unit_base = 0, gens_left = 0, unit_prec = 0, prec_next = 0;
    bool prec_blocked = false;

    Tally tl = {0u, 0u, 0u, 0u};
    uint32_t c_rays = 0, c_hits = 0, c_lit = 0, c_paths = 0;
    const char *const pair_base = reinterpret_cast<const char *>(P.pairs);

    // A finished path: tone-map (:344), deliver the sample to its pixel record; the lane that delivers the last sample of a
    // pixel reduces it in the reference's order and stores the pixel.  Frees the path's record.
    auto finish_path = [&](bool doit, uint32_t sid, V3 light, uint32_t w12, uint32_t first_hit, uint32_t seg, uint32_t n_mh, uint32_t path) {
        uint32_t *rec = nullptr;
        if (doit) {
            rec = precs + ((w12 >> 8) & 0xFFu) * PW;
            const uint32_t si = (w12 >> 16) & 0xFFu;
            float *sm = reinterpret_cast<float *>(rec + kPrecHdr);
            sm[si] = fsqrt(fmaxf(light.x, 0.0f)); sm[spp + si] = fsqrt(fmaxf(light.y, 0.0f)); sm[2 * spp + si] = fsqrt(fmaxf(light.z, 0.0f));
            if (DBG) {
                if (P.dbg_first_hit) P.dbg_first_hit[path] = first_hit;
                if (P.dbg_segments) P.dbg_segments[path] = seg;
                if (P.dbg_mirror_hits) P.dbg_mirror_hits[path] = n_mh >> 16;
                if (P.dbg_radiance) { P.dbg_radiance[3 * (size_t)path] = light.x; P.dbg_radiance[3 * (size_t)path + 1] = light.y; P.dbg_radiance[3 * (size_t)path + 2] = light.z; }
            }
        }
        __syncwarp();
        bool last = false;
        if (doit) last = atomicAdd(rec, 1u) + 1u == spp;
        if (last) {
            const float *sm = reinterpret_cast<const float *>(rec + kPrecHdr);
            const float d = (float)(int)spp;
            float4 px = make_float4(fdiv(reduce_samples(sm, spp, 1), d), fdiv(reduce_samples(sm + spp, spp, 1), d),
                                    fdiv(reduce_samples(sm + 2 * spp, spp, 1), d), 1.0f);
            if (P.quant8) px = quant8(px);
            const uint32_t pxx = rec[1], pxy = rec[2];
            if (P.tiles) reinterpret_cast<float4 *>(P.tiles)[rec[3]] = px;
            if (pxx < P.W && pxy < P.H) {
                const size_t at = (size_t)pxy * P.W + pxx;
                if (P.image) reinterpret_cast<float4 *>(P.image)[at] = px;
                if (P.host_out) reinterpret_cast<float4 *>(P.host_out)[at] = px;
                if (P.peers_multicast) {
                    float4 *mc = reinterpret_cast<float4 *>(P.peers[0]) + at;
                    asm volatile("multimem.st.weak.global.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(mc), "f"(px.x), "f"(px.y), "f"(px.z), "f"(px.w) : "memory");
                } else {
                    for (uint32_t i = 0; i < P.n_peers; i++) reinterpret_cast<float4 *>(P.peers[i])[at] = px;
                }
            }
            rec[0] = 0xFFFFFFFFu;                                                     // pixel record free again
        }
        if (__any_sync(0xFFFFFFFFu, last)) prec_blocked = false;
        qF.push(doit, sid, lane_lt);
    };

    while (true) {
        __syncwarp();
        // ---- choose a body (everything here is warp-uniform) ----
        int body = -1;                                   // 0 G, 1 I, 2 L, 3 S, 4 X
        const bool can_gen = (gens_left > 0 || next < n_units) && !prec_blocked && qF.cnt >= 32u;
        if (can_gen) body = 0;
        else if (qI.cnt >= 32u) body = 1;
        else if (qL.cnt >= P.pool_th_leaf) body = 2;
        else if (qS.cnt >= P.pool_th_shade) body = 3;
        else if (qX.cnt) body = 4;
        else if (qI.cnt | qL.cnt | qS.cnt) body = (qI.cnt >= qL.cnt && qI.cnt >= qS.cnt) ? 1 : (qL.cnt >= qS.cnt ? 2 : 3);
        if (body < 0) break;                             // nothing in flight and nothing left to generate

        if (body == 0) {
            // ================= G: 32 new paths =================
            if (gens_left == 0) {
                unit_base = next * unit_paths;
                gens_left = gens_per_unit;
                uint32_t nn = 0;
                if (lane == 0) nn = (uint32_t)atomicAdd(ticket, 1ull);
                next = __shfl_sync(0xFFFFFFFFu, nn, 0);
            }
            const bool first_gen = gens_left == gens_per_unit;
            // pixel records this generation opens: one per pixel that starts here
            const uint32_t need = spp >= 32u ? (first_gen ? 1u : 0u) : 32u / spp;
            const bool rec_free = lane >= need || precs[((prec_next + lane) % R) * PW] == 0xFFFFFFFFu;
            if (!__all_sync(0xFFFFFFFFu, rec_free)) { prec_blocked = true; continue; }   // a straggler still owns the ring's next record
            const uint64_t path = (uint64_t)unit_base + (uint64_t)(gens_per_unit - gens_left) * 32u + lane;
            gens_left--;
            const bool active = path < P.total_paths;
            uint32_t my_prec;
            if (spp >= 32u) {
                if (first_gen) { unit_prec = prec_next % R; prec_next = (prec_next + 1u) % R; }
                my_prec = unit_prec;
            } else {
                my_prec = (prec_next + lane / spp) % R;
                prec_next = (prec_next + need) % R;
            }
            PathStart ps;
            if (active) ps = start_path(P, path);
            const uint32_t si = active ? (ps.flat & (spp - 1u)) : 0u;
            if (active && si == 0u) {                                                  // the pixel's first sample opens its record
                uint32_t *rec = precs + my_prec * PW;
                rec[0] = 0u; rec[1] = ps.pxx; rec[2] = ps.pxy; rec[3] = ps.k * P.ppc + (ps.flat >> P.log2_spp);
            }
            const uint32_t n_act = __popc(__ballot_sync(0xFFFFFFFFu, active));         // active lanes are 0 .. n_act - 1
            const uint32_t sid = qF.pop(n_act, lane);
            __syncwarp();
            c_paths += active ? 1u : 0u;
            const bool alive = active && 0 < P.bounce_limit;                           // :306, n = 0
            bool lit = false;
            uint32_t w12 = 0;
            if (active) {
                uint32_t *w = slots + sid * SW;
                w12 = 1u | (my_prec << 8) | (si << 16);
                lit = alive && ray_is_literal(P, ps.ori, ps.dir);
                const float rx = __frcp_rn(ps.dir.x), ry = __frcp_rn(ps.dir.y), rz = __frcp_rn(ps.dir.z);
                reinterpret_cast<float4 *>(w)[0] = make_float4(-ps.ori.x, -ps.ori.y, -ps.dir.x, -ps.dir.y);
                reinterpret_cast<float4 *>(w)[1] = make_float4(rx, ry, -ps.ori.z, -ps.dir.z);
                reinterpret_cast<float4 *>(w)[2] = make_float4(rz, 1e30f, __uint_as_float(root), __uint_as_float(0xFFFFFFFFu));
                reinterpret_cast<uint4 *>(w)[3] = make_uint4(w12, ps.state, 0u, (uint32_t)path);
                reinterpret_cast<float4 *>(w)[4] = make_float4(1.0f, 1.0f, 1.0f, 0.0f);               // color, light.x
                reinterpret_cast<float4 *>(w)[5] = make_float4(0.0f, 0.0f, __uint_as_float(0xFFFFFFFFu), __uint_as_float(0u));   // light.yz, first hit, segments
                w[W_STACK] = CUR_END;
            }
            const bool root_leaf = (root & kLeafBit) != 0u;
            qX.push(alive && lit, sid, lane_lt);
            qL.push(alive && !lit && root_leaf, sid, lane_lt);
            qI.push(alive && !lit && !root_leaf, sid, lane_lt);
            finish_path(active && !alive, sid, mk(0.0f, 0.0f, 0.0f), w12, 0xFFFFFFFFu, 0u, 0u, (uint32_t)path);
        } else if (body == 1) {
            // ================= I: interior visits =================
            const uint32_t n = min(32u, qI.cnt);
            const uint32_t sid = qI.pop(n, lane);
            const bool valid = lane < n;
            uint32_t cur = CUR_END;
            uint32_t *w = slots + (valid ? sid : 0u) * SW;
            if (valid) {
                const float4 a = reinterpret_cast<const float4 *>(w)[0], b = reinterpret_cast<const float4 *>(w)[1], c = reinterpret_cast<const float4 *>(w)[2];
                cur = __float_as_uint(c.z);
                const float t = c.y;
                const float rx = b.x, ry = b.y, rz = c.x, ndx = a.z, ndy = a.w, ndz = b.w;
                const float rlx = fmul(__fmaf_rn(ndx, rx, 1.0f), rx), rly = fmul(__fmaf_rn(ndy, ry, 1.0f), ry), rlz = fmul(__fmaf_rn(ndz, rz, 1.0f), rz);
                RayK k;
                k.no_xy = pack2(a.x, a.y); k.nd_xy = pack2(ndx, ndy); k.r_xy = pack2(rx, ry); k.rl_xy = pack2(rlx, rly);
                k.no_zz = pack2(b.z, b.z); k.nd_zz = pack2(ndz, ndz); k.r_zz = pack2(rz, rz); k.rl_zz = pack2(rlz, rlz);
                // travel order folded into the record pointers: -d > 0 means the ray travels down that axis
                const char *pAB = pair_base + (ndx > 0.0f ? 32 : 0) + (ndy > 0.0f ? 64 : 0);
                const char *pZ = pair_base + 128 + (ndz > 0.0f ? 32 : 0);
                uint32_t *const stack = w + W_STACK;
                uint32_t *sp = stack + (w[W_SP] & 0xFFu);
#pragma unroll 2
                for (uint32_t rep = 0; rep < 4u; rep++) {
                    if ((cur & kLeafBit) == 0u) {
                        if (CNT) tl.inner++;
                        const Line32 ab = ldg256(pAB + cur);
                        ulonglong2 A, B;
                        A.x = ab.x; A.y = ab.y; B.x = ab.z; B.y = ab.w;
                        const ulonglong2 Z = __ldg(reinterpret_cast<const ulonglong2 *>(pZ + cur));
                        const uint2 lk = __ldg(reinterpret_cast<const uint2 *>(pZ + cur + 16));
                        inner_step_packed<CNT, false>(A, B, Z, lk, k, t, cur, sp, stack, tl);
                    }
                }
                w[W_CUR] = cur;
                reinterpret_cast<uint8_t *>(w + W_SP)[0] = (uint8_t)(sp - stack);
            }
            const bool done = cur == CUR_END, leaf = !done && (cur & kLeafBit) != 0u;
            qS.push(valid && done, sid, lane_lt);
            qL.push(valid && leaf, sid, lane_lt);
            qI.push(valid && !done && !leaf, sid, lane_lt);
        } else if (body == 2) {
            // ================= L: one leaf visit =================
            const uint32_t n = min(32u, qL.cnt);
            const uint32_t sid = qL.pop(n, lane);
            const bool valid = lane < n;
            uint32_t cur = CUR_END;
            if (valid) {
                uint32_t *w = slots + sid * SW;
                const float4 a = reinterpret_cast<const float4 *>(w)[0], b = reinterpret_cast<const float4 *>(w)[1], c = reinterpret_cast<const float4 *>(w)[2];
                cur = __float_as_uint(c.z);
                float t = c.y;
                uint32_t hit = __float_as_uint(c.w);
                uint32_t *const stack = w + W_STACK;
                uint32_t *sp = stack + (w[W_SP] & 0xFFu);
                if (CNT) tl.leaf++;
                leaf_step<CNT, false>(P.rects, mk(-a.x, -a.y, -b.z), mk(-a.z, -a.w, -b.w), t, hit, cur, sp, tl);
                w[W_T] = __float_as_uint(t); w[W_CUR] = cur; w[W_HIT] = hit;
                reinterpret_cast<uint8_t *>(w + W_SP)[0] = (uint8_t)(sp - stack);
            }
            const bool done = cur == CUR_END, leaf = !done && (cur & kLeafBit) != 0u;
            qS.push(valid && done, sid, lane_lt);
            qL.push(valid && leaf, sid, lane_lt);
            qI.push(valid && !done && !leaf, sid, lane_lt);
        } else if (body == 3) {
            // ================= S: shade a finished traversal =================
            const uint32_t n = min(32u, qS.cnt);
            const uint32_t sid = qS.pop(n, lane);
            const bool valid = lane < n;
            bool alive = false, lit = false;
            V3 light = mk(0.0f, 0.0f, 0.0f);
            uint32_t w12 = 0, first_hit = 0xFFFFFFFFu, seg = 0, n_mh = 0, path = 0;
            if (valid) {
                uint32_t *w = slots + sid * SW;
                const float4 a = reinterpret_cast<const float4 *>(w)[0], b = reinterpret_cast<const float4 *>(w)[1], c = reinterpret_cast<const float4 *>(w)[2];
                const uint4 d = reinterpret_cast<const uint4 *>(w)[3];
                const float4 e = reinterpret_cast<const float4 *>(w)[4], f = reinterpret_cast<const float4 *>(w)[5];
                V3 ori = mk(-a.x, -a.y, -b.z), dir = mk(-a.z, -a.w, -b.w), color = mk(e.x, e.y, e.z);
                light = mk(e.w, f.x, f.y);
                const float t = c.y;
                const uint32_t hit = __float_as_uint(c.w);
                w12 = d.x; n_mh = d.z; path = d.w;
                uint32_t state = d.y;
                first_hit = __float_as_uint(f.z); seg = __float_as_uint(f.w) + 1u;
                int nb = (int)(n_mh & 0xFFFFu), mirror_hits = (int)(n_mh >> 16);
                c_rays++;
                if (t < 1e30f) {                                                        // :308
                    c_hits++;
                    uint32_t orig = 0xFFFFFFFFu;
                    alive = shade_hit(P, hit, t, ori, dir, color, light, state, mirror_hits, (DBG && nb == 0) ? &orig : nullptr);
                    if (DBG && nb == 0) first_hit = orig;
                    nb++;
                    alive = alive && (nb < P.bounce_limit + mirror_hits);               // :306
                }
                n_mh = (uint32_t)nb | ((uint32_t)mirror_hits << 16);
                if (alive) {
                    lit = ray_is_literal(P, ori, dir);
                    const float rx = __frcp_rn(dir.x), ry = __frcp_rn(dir.y), rz = __frcp_rn(dir.z);
                    reinterpret_cast<float4 *>(w)[0] = make_float4(-ori.x, -ori.y, -dir.x, -dir.y);
                    reinterpret_cast<float4 *>(w)[1] = make_float4(rx, ry, -ori.z, -dir.z);
                    reinterpret_cast<float4 *>(w)[2] = make_float4(rz, 1e30f, __uint_as_float(root), __uint_as_float(0xFFFFFFFFu));   // :323, :330
                    reinterpret_cast<uint4 *>(w)[3] = make_uint4((w12 & 0xFFFFFF00u) | 1u, state, n_mh, path);
                    reinterpret_cast<float4 *>(w)[4] = make_float4(color.x, color.y, color.z, light.x);
                    reinterpret_cast<float4 *>(w)[5] = make_float4(light.y, light.z, __uint_as_float(first_hit), __uint_as_float(seg));
                }
            }
            const bool root_leaf = (root & kLeafBit) != 0u;
            qX.push(alive && lit, sid, lane_lt);
            qL.push(alive && !lit && root_leaf, sid, lane_lt);
            qI.push(alive && !lit && !root_leaf, sid, lane_lt);
            finish_path(valid && !alive, sid, light, w12, first_hit, seg, n_mh, path);
        } else {
            // ================= X: whole literal-divide traversals (rays outside the guarded operand ranges) =================
            const uint32_t n = min(32u, qX.cnt);
            const uint32_t sid = qX.pop(n, lane);
            const bool valid = lane < n;
            uint32_t *w = slots + (valid ? sid : 0u) * SW;
            V3 ori = mk(0.0f, 0.0f, 0.0f), dir = mk(1.0f, 1.0f, 1.0f);
            if (valid) {
                const float4 a = reinterpret_cast<const float4 *>(w)[0], b = reinterpret_cast<const float4 *>(w)[1];
                ori = mk(-a.x, -a.y, -b.z); dir = mk(-a.z, -a.w, -b.w);
                c_lit++;
            }
            const Hit h = traverse<true, CNT, false>(P.pairs, P.rects, root, valid, true, ori, dir, 1e30f, 0xFFFFFFFFu, &tl);
            if (valid) { w[W_T] = __float_as_uint(h.t); w[W_HIT] = h.slot; w[W_CUR] = CUR_END; }
            qS.push(valid, sid, lane_lt);
        }
    }

    // Event counts: warp-reduce, one atomic per warp and counter.
    {
        unsigned long long v_rays = c_rays, v_hits = c_hits, v_lit = c_lit, v_paths = c_paths;
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
        if (lane == 0u) {
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

}  // namespace

const void *pool_kernel_ptr(KernelChoice c) {
    if (c.debug) return reinterpret_cast<const void *>(&pool_kernel<true, true>);
    return c.counters ? reinterpret_cast<const void *>(&pool_kernel<true, false>) : reinterpret_cast<const void *>(&pool_kernel<false, false>);
}

cudaError_t launch_pool(const KParams &p, KernelChoice c, unsigned blocks, unsigned threads, size_t smem_bytes, cudaStream_t stream) {
    void *args[] = {const_cast<KParams *>(&p)};
    return cudaLaunchKernel(pool_kernel_ptr(c), dim3(blocks), dim3(threads), args, smem_bytes, stream);
}

}  // namespace mmk
