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
#include "render_kernel.cuh"

namespace mmk {
namespace {

struct V3 { float x, y, z; };

__device__ __forceinline__ float fadd(float a, float b) { return __fadd_rn(a, b); }
__device__ __forceinline__ float fsub(float a, float b) { return __fsub_rn(a, b); }
__device__ __forceinline__ float fmul(float a, float b) { return __fmul_rn(a, b); }
__device__ __forceinline__ float fdiv(float a, float b) { return __fdiv_rn(a, b); }
__device__ __forceinline__ float fsqrt(float a) { return __fsqrt_rn(a); }

__device__ __forceinline__ V3 mk(float x, float y, float z) { V3 r; r.x = x; r.y = y; r.z = z; return r; }
__device__ __forceinline__ V3 add3(V3 a, V3 b) { return mk(fadd(a.x, b.x), fadd(a.y, b.y), fadd(a.z, b.z)); }
__device__ __forceinline__ V3 sub3(V3 a, V3 b) { return mk(fsub(a.x, b.x), fsub(a.y, b.y), fsub(a.z, b.z)); }
__device__ __forceinline__ V3 mul3(V3 a, V3 b) { return mk(fmul(a.x, b.x), fmul(a.y, b.y), fmul(a.z, b.z)); }
__device__ __forceinline__ V3 scale3(V3 a, float s) { return mk(fmul(a.x, s), fmul(a.y, s), fmul(a.z, s)); }
__device__ __forceinline__ float dot3(V3 a, V3 b) { return fadd(fadd(fmul(a.x, b.x), fmul(a.y, b.y)), fmul(a.z, b.z)); }
__device__ __forceinline__ V3 cross3(V3 a, V3 b) {
    return mk(fsub(fmul(a.y, b.z), fmul(a.z, b.y)), fsub(fmul(a.z, b.x), fmul(a.x, b.z)), fsub(fmul(a.x, b.y), fmul(a.y, b.x)));
}
__device__ __forceinline__ float length3(V3 a) { return fsqrt(dot3(a, a)); }
__device__ __forceinline__ V3 normalize3(V3 a) { float l = length3(a); return mk(fdiv(a.x, l), fdiv(a.y, l), fdiv(a.z, l)); }
__device__ __forceinline__ V3 reflect3(V3 i, V3 n) { return sub3(i, scale3(n, fmul(2.0f, dot3(n, i)))); }
__device__ __forceinline__ float sign1(float x) { return x > 0.0f ? 1.0f : (x < 0.0f ? -1.0f : 0.0f); }

// random() of shaders.metal:181-186 (state = state * 747796405 + 291336453; PCG output hash; float(result) / 2^32) is only
// ever used as (random(state) - 0.5) * 2.0 (:303, :315-317), evaluated here as one FMA: float(r) * 2^-32 is an exact scaling, and so is the
// final * 2, hence 2 * RN(f * 2^-32 - 0.5) == RN(f * 2^-31 - 1), which is what the FMA's single rounding returns
// (tests/test_oracle.py checks the identity over 2^24 random words and the edge words).
__device__ __forceinline__ float rnd_pm1(uint32_t &state) {
    state = state * 747796405u + 291336453u;
    uint32_t result = ((state >> ((state >> 28) + 4u)) ^ state) * 277803737u;
    result = (result >> 22) ^ result;
    return __fmaf_rn(__uint2float_rn(result), 4.656612873077393e-10f /*2^-31*/, -1.0f);
}

// shaders.metal:163-172
struct Q4 { float x, y, z, w; };
__device__ __forceinline__ Q4 quat_dot(Q4 q1, Q4 q2) {
    V3 a = mk(q1.x, q1.y, q1.z), b = mk(q2.x, q2.y, q2.z);
    float s = fsub(fmul(q1.w, q2.w), dot3(a, b));
    V3 v = add3(add3(cross3(a, b), scale3(b, q1.w)), scale3(a, q2.w));
    Q4 r = {v.x, v.y, v.z, s};
    return r;
}
__device__ __forceinline__ V3 quat_mult(V3 vec, Q4 q) {
    Q4 inv = {-q.x, -q.y, -q.z, q.w};
    Q4 v4 = {vec.x, vec.y, vec.z, 0.0f};
    Q4 r = quat_dot(quat_dot(inv, v4), q);
    return mk(r.x, r.y, r.z);
}

struct Axis { float o, d, r, rl; };

// RCP (MM_FLAG_RCP_SLAB): the opt-in reciprocal-multiply slab arithmetic t = (b - o) * RN(1/d) — what a fast-math compile
// of the reference's divide amounts to; the oracle implements the same rule under the same flag.
template <bool FAST, bool RCP = false>
__device__ __forceinline__ float quot(float b, const Axis &a) {
    float x = fsub(b, a.o);
    if (RCP) return fmul(x, a.r);
    if (FAST) {
        float p = fmul(x, a.rl);
        float q1 = __fmaf_rn(x, a.r, p);
        float e = __fmaf_rn(-a.d, q1, x);
        return __fmaf_rn(e, a.r, q1);
    } else {
        return fdiv(x, a.d);
    }
}

__device__ __forceinline__ bool axis_safe(float o, float d) {
    float ad = fabsf(d), ao = fabsf(o);
    bool dok = ad >= 8.673617379884035e-19f /*2^-60*/ && ad <= 1.152921504606847e18f /*2^60*/;
    bool ook = ao == 0.0f || (ao >= 9.094947017729282e-13f /*2^-40*/ && ao <= 1073741824.0f /*2^30*/);
    return dok && ook;
}

// Reciprocal-multiply mode: the travel-ordered form needs finite operands and a finite non-zero reciprocal (no NaN from
// 0 * inf, monotone products); anything else takes the general min/max form.
__device__ __forceinline__ bool rcp_safe(float o, float d) {
    const float ad = fabsf(d);
    return ad >= 1.1754943508222875e-38f /*2^-126*/ && ad <= 8.507059173023462e37f /*2^126*/ && fabsf(o) <= 3.4028234663852886e38f;
}

struct Tally { uint32_t inner, leaf, rect, max_stack; };

constexpr uint32_t CUR_END = 0xFFFFFFFFu;   // traversal finished

// ---- packed FP32 (Blackwell FADD2 / FMUL2 / FFMA2: two IEEE-RN fp32 operations per issued instruction) ------------------
// The kernel is bound by instruction issue, not by the FMA pipe (ncu: issue slots 85 % busy, FMA pipe 43 %), and the
// twelve slab quotients of a visit are six pairs of identical, independent operation chains — so they are issued as
// add.rn.f32x2 / mul.rn.f32x2 / fma.rn.f32x2 on register pairs: per lane the results are the bits of the scalar
// instructions, at half the issue slots.
typedef unsigned long long f2;   // two fp32 in a 64-bit register pair (low word = first component)
__device__ __forceinline__ f2 pack2(float lo, float hi) {
    f2 r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
    return r;
}
__device__ __forceinline__ float lo2f(f2 v) { float a, b; asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(v)); (void)b; return a; }
__device__ __forceinline__ float hi2f(f2 v) { float a, b; asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(v)); (void)a; return b; }
__device__ __forceinline__ f2 add2(f2 a, f2 b) { f2 d; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }
__device__ __forceinline__ f2 mul2(f2 a, f2 b) { f2 d; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }
__device__ __forceinline__ f2 fma2(f2 a, f2 b, f2 c) { f2 d; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c)); return d; }

// 32-byte read-only load (LDG.E.256, sm_100+): one instruction and one L1 request per half record.
struct Line32 { f2 x, y, z, w; };
__device__ __forceinline__ Line32 ldg256(const void *p) {
    Line32 q;
    asm("ld.global.nc.v4.u64 {%0, %1, %2, %3}, [%4];" : "=l"(q.x), "=l"(q.y), "=l"(q.z), "=l"(q.w) : "l"(p));
    return q;
}

// Per-ray constants of the packed slab test: x and y share register pairs, z is duplicated.
struct RayK {
    f2 no_xy, nd_xy, r_xy, rl_xy;   // (-o.x, -o.y), (-d.x, -d.y), (r.x, r.y), (rl.x, rl.y)
    f2 no_zz, nd_zz, r_zz, rl_zz;
};

// Two slab quotients at once: x = b - o, then either the exact shared-reciprocal sequence (== RN(x/d), see the file header
// and docs/exact_quotient.md) or, in the opt-in MM_FLAG_RCP_SLAB arithmetic, x * RN(1/d).
template <bool RCP>
__device__ __forceinline__ f2 quot2(f2 b, f2 no, f2 nd, f2 r, f2 rl) {
    const f2 x = add2(b, no);                  // b + (-o) == b - o
    if (RCP) return mul2(x, r);
    const f2 p = mul2(x, rl);
    const f2 q1 = fma2(x, r, p);
    const f2 e = fma2(nd, q1, x);              // x - d*q1, exact
    return fma2(e, r, q1);
}

// The decisions of one interior visit (shaders.metal:140-154) from the children's [lo, hi] slab intervals.
// The literal `dist` values are never materialised: with hit_k = (hi_k >= lo_k && lo_k < t && hi_k > 0) and
// dist_k = hit_k ? lo_k : 1e30 (lo_k < t <= 1e30 when hit), `dist1 > dist2` is hit2 && (!hit1 || lo1 > lo2),
// `dist_near == 1e30` is !hit1 && !hit2 and `dist_far != 1e30` is hit1 && hit2 — the same decisions, fewer instructions.
template <bool CNT>
__device__ __forceinline__ void descend(float lo1, float hi1, float lo2, float hi2, const uint2 &lk, float t, uint32_t &cur, uint32_t *&sp,
                                        const uint32_t *stack, Tally &tl) {
    const bool hit1 = (hi1 >= lo1) & (lo1 < t) & (hi1 > 0.0f);          // :94
    const bool hit2 = (hi2 >= lo2) & (lo2 < t) & (hi2 > 0.0f);
    const bool swap = hit2 & (!hit1 | (lo1 > lo2));                     // :140, ties keep the left child first
    if (!(hit1 | hit2)) {                                               // :149-150
        cur = *--sp;                                                    // the bottom entry is the CUR_END sentinel
    } else {                                                            // :151-154
        cur = swap ? lk.y : lk.x;
        if (hit1 & hit2) {
            *sp++ = swap ? lk.x : lk.y;
            if (CNT) tl.max_stack = max(tl.max_stack, (uint32_t)(sp - stack) - 1u);   // entries above the sentinel
        }
    }
}

// One interior visit, fast form (guarded operand ranges: no zero / NaN / inf anywhere).  The quotients are the exact
// RN((b - o)/d) of the literal code; the record is loaded in the ray's travel order on every axis, so each axis' near
// plane is the first and its far plane the second value (rounding is monotone: min(t1,t2) is the near plane's quotient):
//   A = (c0.near.x, c0.near.y | c0.far.x, c0.far.y)   B likewise for child 1   Z = (c0.near.z, c1.near.z | c0.far.z, c1.far.z)
template <bool CNT, bool RCP>
__device__ __forceinline__ void inner_step_packed(const ulonglong2 &A, const ulonglong2 &B, const ulonglong2 &Z, const uint2 &lk,
                                                  const RayK &k, float t, uint32_t &cur, uint32_t *&sp, const uint32_t *stack, Tally &tl) {
    const f2 an = quot2<RCP>(A.x, k.no_xy, k.nd_xy, k.r_xy, k.rl_xy), af = quot2<RCP>(A.y, k.no_xy, k.nd_xy, k.r_xy, k.rl_xy);
    const f2 bn = quot2<RCP>(B.x, k.no_xy, k.nd_xy, k.r_xy, k.rl_xy), bf = quot2<RCP>(B.y, k.no_xy, k.nd_xy, k.r_xy, k.rl_xy);
    const f2 zn = quot2<RCP>(Z.x, k.no_zz, k.nd_zz, k.r_zz, k.rl_zz), zf = quot2<RCP>(Z.y, k.no_zz, k.nd_zz, k.r_zz, k.rl_zz);
    const float lo1 = fmaxf(fmaxf(lo2f(an), hi2f(an)), lo2f(zn));
    const float hi1 = fminf(fminf(lo2f(af), hi2f(af)), lo2f(zf));
    const float lo2 = fmaxf(fmaxf(lo2f(bn), hi2f(bn)), hi2f(zn));
    const float hi2 = fminf(fminf(lo2f(bf), hi2f(bf)), hi2f(zf));
    descend<CNT>(lo1, hi1, lo2, hi2, lk, t, cur, sp, stack, tl);
}

// One interior visit, general form: the literal min/max of shaders.metal:88-93 with NaN-dropping fmin/fmax, for rays whose
// operands are outside the guarded ranges (zero / subnormal / huge / NaN components) or under MM_FLAG_FORCE_LITERAL.
// Reads the record in its "up" order: a = (c0.min.x, c0.min.y, c0.max.x, c0.max.y), zu = (c0.min.z, c1.min.z, c0.max.z, c1.max.z).
template <bool CNT, bool RCP>
__device__ __forceinline__ void inner_step_general(const float4 &a, const float4 &b, const float4 &zu, const uint2 &lk, const Axis &ax,
                                                   const Axis &ay, const Axis &az, float t, uint32_t &cur, uint32_t *&sp,
                                                   const uint32_t *stack, Tally &tl) {
    float t1 = quot<false, RCP>(a.x, ax), t2 = quot<false, RCP>(a.z, ax);
    float lo1 = fminf(t1, t2), hi1 = fmaxf(t1, t2);
    t1 = quot<false, RCP>(a.y, ay); t2 = quot<false, RCP>(a.w, ay);
    lo1 = fmaxf(lo1, fminf(t1, t2)); hi1 = fminf(hi1, fmaxf(t1, t2));
    t1 = quot<false, RCP>(zu.x, az); t2 = quot<false, RCP>(zu.z, az);
    lo1 = fmaxf(lo1, fminf(t1, t2)); hi1 = fminf(hi1, fmaxf(t1, t2));
    t1 = quot<false, RCP>(b.x, ax); t2 = quot<false, RCP>(b.z, ax);
    float lo2 = fminf(t1, t2), hi2 = fmaxf(t1, t2);
    t1 = quot<false, RCP>(b.y, ay); t2 = quot<false, RCP>(b.w, ay);
    lo2 = fmaxf(lo2, fminf(t1, t2)); hi2 = fminf(hi2, fmaxf(t1, t2));
    t1 = quot<false, RCP>(zu.y, az); t2 = quot<false, RCP>(zu.w, az);
    lo2 = fmaxf(lo2, fminf(t1, t2)); hi2 = fminf(hi2, fmaxf(t1, t2));
    descend<CNT>(lo1, hi1, lo2, hi2, lk, t, cur, sp, stack, tl);
}

// One leaf visit (shaders.metal:127-129 with ray_rect_intersect :51-67 inlined), then pop / finish.
// LITERAL = false: the two edge tests `0 <= RN(x / L) <= L` are evaluated as the equivalent interval test on x stored in
// the record (render_kernel.cuh, RectI) — no divides by the edge lengths.  LITERAL = true (scenes with an edge length
// outside the guarded range, or MM_FLAG_FORCE_LITERAL): the literal divides, with the lengths recomputed by the same
// operations the upload used.
template <bool CNT, bool LITERAL>
__device__ __forceinline__ void leaf_step(const RectI *__restrict__ rects, V3 ori, V3 dir, float &t, uint32_t &slot, uint32_t &cur,
                                          uint32_t *&sp, Tally &tl) {
    const uint32_t first = cur & 0xFFFFFFu, count = (cur >> 24) & 0x7Fu;
#pragma unroll 1                               // leaves hold one rect almost always (leaf visits ~ rect tests): no unrolled copy, -1.3 %
    for (uint32_t i = 0; i < count; i++) {
        const float4 *rp = reinterpret_cast<const float4 *>(rects + first + i);
        const float4 r0 = __ldg(rp), r1 = __ldg(rp + 1), r2 = __ldg(rp + 2), r3 = __ldg(rp + 3);
        if (CNT) tl.rect++;
        const V3 ro = mk(r0.x, r0.y, r0.z), n = mk(r1.x, r1.y, r1.z), v = mk(r2.x, r2.y, r2.z), u = mk(r3.x, r3.y, r3.z);
        const float norm_check = dot3(dir, n);                                   // :53
        const float a = fdiv(dot3(sub3(ro, ori), n), norm_check);                // :55
        const V3 isect = add3(ori, scale3(dir, a));                              // :56
        const V3 rv = sub3(isect, ro);                                           // :58
        const float xv = dot3(rv, v), xu = dot3(rv, u);
        bool inside;
        if (LITERAL) {
            const float len_v = length3(v), len_u = length3(u);
            const float d1 = fdiv(xv, len_v);                                    // :60
            const float d2 = fdiv(xu, len_u);                                    // :61
            inside = (0.0f <= d1 && d1 <= len_v) && (0.0f <= d2 && d2 <= len_u);
        } else {
            inside = (r2.w <= xv) & (xv <= r0.w) & (r3.w <= xu) & (xu <= r1.w);
        }
        if (inside && norm_check != 0.0f && a > 0.1f && a < t) {                 // :63
            t = a;
            slot = first + i;
        }
    }
    cur = *--sp;
}

// intersect_bvh_iterative (shaders.metal:115-156) for the rays of one warp.  Every lane of the warp calls this together
// (lanes without a ray pass alive = false) and the warp votes on which body to execute: the interior body runs (kInnerReps
// visits per vote) while the lanes standing at an interior node outweigh the lanes waiting at a leaf
// (nI >= kLeafWeight * nL), otherwise the waiting lanes test their rects.  Each lane still performs exactly the
// reference's sequence of visits for its own ray; only the interleaving between lanes changes.  (A plain while-while
// loop — all lanes descend to a leaf, then all test — left 12 of 32 lanes active in the interior body; see profiles/.)
// `lit` lanes (operands outside the guarded ranges, or MM_FLAG_FORCE_LITERAL) use the general slab form.
#ifndef MM_LEAF_WEIGHT
#define MM_LEAF_WEIGHT 6
#endif
constexpr uint32_t kLeafWeight = MM_LEAF_WEIGHT;   // measured best on B200 (profiles/r1_sched_sweep.txt)
#ifndef MM_INNER_REPS
#define MM_INNER_REPS 4
#endif
constexpr uint32_t kInnerReps = MM_INNER_REPS;
constexpr int kRepUnroll = 2;                      // visits per loop trip: 2 measured best (1: +0.6 %, 4: +1.3 %)

// MIXED = false: no lane of the warp is literal (the common case; the loop then contains no general-form code).
// Not inlined on purpose: the call boundary parks the path state that the traversal does not touch (throughput,
// radiance, RNG state, counters, pixel bookkeeping) in the caller's frame, so the traversal loop has the whole 64-register
// budget for its per-ray constants.
struct Hit { float t; uint32_t slot; };
template <bool MIXED, bool CNT, bool RCP>
__device__ __noinline__ Hit traverse(const PairRec *__restrict__ pairs, const RectI *__restrict__ rects, uint32_t root, bool alive,
                                     bool lit, V3 ori, V3 dir, float beam_t, uint32_t beam_slot, Tally *tlp) {
    Tally tl = {0u, 0u, 0u, 0u};
    uint32_t stack[MM_MAX_STACK];              // local memory (L1); BVH depth is validated against it at upload
    const float rx = __frcp_rn(dir.x), ry = __frcp_rn(dir.y), rz = __frcp_rn(dir.z);
    const float rlx = fmul(__fmaf_rn(-dir.x, rx, 1.0f), rx), rly = fmul(__fmaf_rn(-dir.y, ry, 1.0f), ry),
                rlz = fmul(__fmaf_rn(-dir.z, rz, 1.0f), rz);
    RayK k;
    k.no_xy = pack2(-ori.x, -ori.y); k.nd_xy = pack2(-dir.x, -dir.y); k.r_xy = pack2(rx, ry); k.rl_xy = pack2(rlx, rly);
    k.no_zz = pack2(-ori.z, -ori.z); k.nd_zz = pack2(-dir.z, -dir.z); k.r_zz = pack2(rz, rz); k.rl_zz = pack2(rlz, rlz);
    // per-ray record pointers with the travel order folded in: ab[sx + 2 sy] at +32 (sx + 2 sy), z|link[sz] at +128 + 32 sz
    const char *base = reinterpret_cast<const char *>(pairs);
    const char *pAB = base + (lit ? 0 : (dir.x < 0.0f ? 32 : 0) + (dir.y < 0.0f ? 64 : 0));
    const char *pZ = base + 128 + ((!lit && dir.z < 0.0f) ? 32 : 0);
    asm("" : "+l"(pAB)); asm("" : "+l"(pZ));      // keep them live: ptxas otherwise re-derives them from sign(dir) at every node
    // The pair table lies inside one 4-GB-aligned window (checked at upload), so a record address is {hi, lo + offset} with
    // no carry: one 32-bit add per pointer instead of a 64-bit add (two instructions).
    uint32_t ab_lo = (uint32_t)reinterpret_cast<uintptr_t>(pAB), z_lo = (uint32_t)reinterpret_cast<uintptr_t>(pZ);
    uint32_t hi_a = (uint32_t)(reinterpret_cast<uintptr_t>(pAB) >> 32), hi_z = (uint32_t)(reinterpret_cast<uintptr_t>(pZ) >> 32);   // two registers on purpose
    asm("" : "+r"(ab_lo)); asm("" : "+r"(z_lo)); asm("" : "+r"(hi_a)); asm("" : "+r"(hi_z));
    stack[0] = CUR_END;                        // sentinel: popping an empty stack ends the traversal, no emptiness test
    uint32_t *sp = stack + 1;                  // next free entry; a pointer, so push and pop need no address arithmetic
    uint32_t cur = alive ? root : CUR_END, slot = beam_slot;
    float t = beam_t;
    while (true) {
        const bool isI = (cur & kLeafBit) == 0u;
        const bool isL = !isI && cur != CUR_END;
        const unsigned mI = __ballot_sync(0xFFFFFFFFu, isI), mL = __ballot_sync(0xFFFFFFFFu, isL);
        if ((mI | mL) == 0u) break;
        if (mI != 0u && __popc(mI) >= kLeafWeight * __popc(mL)) {
#pragma unroll kRepUnroll
            for (uint32_t rep = 0; rep < kInnerReps; rep++) {
                if ((cur & kLeafBit) == 0u) {
                    const size_t off = cur;                      // interior descriptors are byte offsets
                    if (CNT) tl.inner++;
                    if (!MIXED || !lit) {
                        // 32-B load for (A, B): measured 1.5 % faster than two 16-B loads; folding (Z, link) into a second
                        // 32-B load gave nothing (profiles/r1_block_shape.txt)
                        uint64_t aAB, aZ;
                        asm("mov.b64 %0, {%1, %2};" : "=l"(aAB) : "r"(ab_lo + cur), "r"(hi_a));
                        asm("mov.b64 %0, {%1, %2};" : "=l"(aZ) : "r"(z_lo + cur), "r"(hi_z));
                        const Line32 ab = ldg256(reinterpret_cast<const void *>(aAB));
                        ulonglong2 A, B;
                        A.x = ab.x; A.y = ab.y; B.x = ab.z; B.y = ab.w;
                        const ulonglong2 Z = __ldg(reinterpret_cast<const ulonglong2 *>(aZ));
                        const uint2 lk = __ldg(reinterpret_cast<const uint2 *>(aZ + 16));
                        inner_step_packed<CNT, RCP>(A, B, Z, lk, k, t, cur, sp, stack, tl);
                    } else {
                        const uint2 lk = __ldg(reinterpret_cast<const uint2 *>(base + 144 + off));
                        const float4 a = __ldg(reinterpret_cast<const float4 *>(base + off)), b = __ldg(reinterpret_cast<const float4 *>(base + off + 16));
                        const float4 zu = __ldg(reinterpret_cast<const float4 *>(base + off + 128));
                        Axis ax, ay, az;
                        ax.o = ori.x; ax.d = dir.x; ax.r = rx; ax.rl = rlx;
                        ay.o = ori.y; ay.d = dir.y; ay.r = ry; ay.rl = rly;
                        az.o = ori.z; az.d = dir.z; az.r = rz; az.rl = rlz;
                        inner_step_general<CNT, RCP>(a, b, zu, lk, ax, ay, az, t, cur, sp, stack, tl);
                    }
                }
            }
        } else {
            if (isL) {
                if (CNT) tl.leaf++;
                leaf_step<CNT, MIXED>(rects, ori, dir, t, slot, cur, sp, tl);
            }
        }
    }
    if (CNT) { tlp->inner += tl.inner; tlp->leaf += tl.leaf; tlp->rect += tl.rect; tlp->max_stack = max(tlp->max_stack, tl.max_stack); }
    Hit h;
    h.t = t; h.slot = slot;
    return h;
}

// noise.sample(s, float2(gid)): normalised coordinates, address::repeat, filter::nearest (shaders.metal:288,291).
__device__ __forceinline__ void sample_noise_xy(const uint8_t *noise, uint32_t nw, uint32_t nh, float u, float v, float &nx, float &ny) {
    float fu = fsub(u, floorf(u)), fv = fsub(v, floorf(v));
    int ix = (int)floorf(fmul(fu, (float)nw)), iy = (int)floorf(fmul(fv, (float)nh));
    ix = min(max(ix, 0), (int)nw - 1);
    iy = min(max(iy, 0), (int)nh - 1);
    const uint8_t *tx = noise + 4 * ((size_t)iy * nw + (size_t)ix);
    uchar4 c = *reinterpret_cast<const uchar4 *>(tx);
    nx = fdiv((float)c.x, 255.0f);
    ny = fdiv((float)c.y, 255.0f);
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
        k = (uint32_t)(path / P.T);
        flat = (uint32_t)(path - (uint64_t)k * P.T);
        const uint32_t g = P.group_first + k * P.group_step;
        const uint32_t tgx = g % P.grid_x, tgy = g / P.grid_x;
        const mm_chunk ch = P.chunks[g];                                   // :266-267
        const uint32_t gx = flat % P.dim_x, gy = flat / P.dim_x;           // inverse of :271
        const uint32_t chunk = P.uni.chunk_width;
        const uint32_t pixel_number = flat >> P.log2_spp;                  // :272
        pxx = ch.x + pixel_number / chunk;                                 // :274-275
        pxy = ch.y + pixel_number % chunk;                                 // :273,275
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
        // :298 — float + uint promotes to float, products wrap in u32, left to right; float->uint saturates.
        const float seed_f = fadd(fadd(fadd(fadd(nx, ny), __uint2float_rn(texid_x * 15823u)), __uint2float_rn(texid_y * 9737333u)),
                                  __uint2float_rn(P.uni.time));
        uint32_t state = __float2uint_rz(seed_f);

        V3 ori = center;                                                   // :302
        const float j1 = rnd_pm1(state), j2 = rnd_pm1(state);
        V3 dir = add3(ray_dir, scale3(mk(j1, j2, 0.0f), 0.001f));          // :303
        float t = 1e30f;
        uint32_t slot = 0xFFFFFFFFu;
        V3 color = mk(1.0f, 1.0f, 1.0f), light = mk(0.0f, 0.0f, 0.0f);

        n_alive = 0 < P.bounce_limit;                                      // :306, n = 0
        st_ori = ori; st_dir = dir; st_t = t; st_slot = slot; st_color = color; st_light = light; st_state = state;
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
            // lit: this ray must use the general slab form (operands outside the guarded ranges / flags)
            const bool lit = P.force_literal || !P.scene_fast_ok ||
                             !(P.rcp_mode ? (rcp_safe(ori.x, dir.x) && rcp_safe(ori.y, dir.y) && rcp_safe(ori.z, dir.z))
                                          : (axis_safe(ori.x, dir.x) && axis_safe(ori.y, dir.y) && axis_safe(ori.z, dir.z)));
            const bool any_lit = __any_sync(0xFFFFFFFFu, alive && lit) || !P.rect_fast_ok;   // MIXED also means literal rect tests
            Hit h;
            if (P.rcp_mode) {
                if (!any_lit) h = traverse<false, CNT, true>(P.pairs, P.rects, root, alive, false, ori, dir, t, slot, &tl);
                else h = traverse<true, CNT, true>(P.pairs, P.rects, root, alive, lit, ori, dir, t, slot, &tl);
            } else {
                if (!any_lit) h = traverse<false, CNT, false>(P.pairs, P.rects, root, alive, false, ori, dir, t, slot, &tl);
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
                    const float4 *rp = reinterpret_cast<const float4 *>(P.rects + slot);
                    const float4 r1 = __ldg(rp + 1);
                    const V3 nrm = mk(r1.x, r1.y, r1.z);                   // :309 (per-rect constant, same operations)
                    const float side = -sign1(dot3(dir, nrm));             // :310
                    const float4 *sp4 = reinterpret_cast<const float4 *>(P.shade + slot);
                    const float4 col = __ldg(sp4);                         // albedo, material bits in .w
                    if (DBG && n == 0) first_hit = __float_as_uint(__ldg(sp4 + 1).w);
                    if (__float_as_uint(col.w) == 0u || side == -1.0f) {   // :311
                        const float4 emi = __ldg(sp4 + 1);
                        light = add3(light, mul3(mk(emi.x, emi.y, emi.z), color));   // :312-313
                        color = mul3(color, mk(col.x, col.y, col.z));      // :314
                        // :315-318 rejection loop `while (length(rd) > 1)`.  RN(sqrt(s)) > 1 <=> s > 1 + 2^-23: sqrt is
                        // monotone, sqrt(1 + 2^-23) = 1 + 2^-24 - ... lies below the midpoint and rounds to 1, and
                        // sqrt(1 + 2^-22) rounds above 1 (checked over every float in [0.5, 2) in tests/test_oracle.py),
                        // so the loop compares the squared length and the square root is taken once, after it.
                        V3 rd;
                        float s2;
                        do {
                            const float a = rnd_pm1(state), b = rnd_pm1(state), c = rnd_pm1(state);
                            rd = mk(a, b, c);
                            s2 = dot3(rd, rd);
                        } while (s2 > 1.00000011920928955f);
                        const float rl = fsqrt(s2);                        // :319 normalize = v / length(v)
                        rd = mk(fdiv(rd.x, rl), fdiv(rd.y, rl), fdiv(rd.z, rl));
                        ori = add3(ori, scale3(dir, t));                   // :320
                        dir = normalize3(add3(rd, scale3(nrm, side)));     // :321
                        t = 1e30f;                                         // :323
                    } else {
                        mirror_hits++;                                     // :325
                        if (mirror_hits < P.mirror_limit) {                // :326
                            light = add3(light, scale3(mk(col.x, col.y, col.z), 0.005f));   // :327
                            ori = add3(ori, scale3(dir, t));               // :328
                            dir = normalize3(reflect3(dir, nrm));          // :329
                            t = 1e30f;                                     // :330
                        } else {
                            alive = false;                                 // :333
                        }
                    }
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

    // Per-pixel reduction in the reference's order (shaders.metal:343-366): pairs, quads, octets (the phases whose
    // stride is < spp), then the pixel's first thread adds the octets serially and divides by spp.
    const uint32_t tid = threadIdx.x;
    float *rx = red, *ry = red + kBlockThreads, *rz = red + 2 * kBlockThreads;
    rx[tid] = sample.x; ry[tid] = sample.y; rz[tid] = sample.z;
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
    if (active && (flat & (P.spp - 1)) == 0) {
        float sx = rx[tid], sy = ry[tid], sz = rz[tid];
        for (uint32_t i = 1; i < P.spp / 8; i++) {
            sx = fadd(sx, rx[tid + 8 * i]); sy = fadd(sy, ry[tid + 8 * i]); sz = fadd(sz, rz[tid + 8 * i]);
        }
        const float d = (float)(int)P.spp;
        const float4 px = make_float4(fdiv(sx, d), fdiv(sy, d), fdiv(sz, d), 1.0f);
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
__global__ void __launch_bounds__(256) blur_kernel(const float4 *__restrict__ src, float4 *__restrict__ dst, uint32_t W, uint32_t H) {
    const uint32_t x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
    if (x >= W || y >= H) return;
    const float4 zero = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
    const size_t row = (size_t)y * W;
    const float4 c = __ldg(src + row + x);
    const float4 r = x + 1 < W ? __ldg(src + row + x + 1) : zero, l = x > 0 ? __ldg(src + row + x - 1) : zero;
    const float4 d = y + 1 < H ? __ldg(src + row + W + x) : zero, u = y > 0 ? __ldg(src + row - W + x) : zero;
    float4 o;
    o.x = fdiv(fadd(fadd(c.x, fdiv(fadd(r.x, l.x), 2.0f)), fdiv(fadd(d.x, u.x), 2.0f)), 3.0f);
    o.y = fdiv(fadd(fadd(c.y, fdiv(fadd(r.y, l.y), 2.0f)), fdiv(fadd(d.y, u.y), 2.0f)), 3.0f);
    o.z = fdiv(fadd(fadd(c.z, fdiv(fadd(r.z, l.z), 2.0f)), fdiv(fadd(d.z, u.z), 2.0f)), 3.0f);
    o.w = 1.0f;
    dst[row + x] = o;
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
const void *kptr(int bt) {
    return bt == kSmallBlock ? reinterpret_cast<const void *>(&trace_kernel<C, D, kSmallBlock>)
                             : reinterpret_cast<const void *>(&trace_kernel<C, D, kLargeBlock>);
}

}  // namespace

const void *kernel_ptr(KernelChoice c) {
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

cudaError_t launch_blur(const float *src, float *dst, uint32_t W, uint32_t H, cudaStream_t stream) {
    if (W == 0 || H == 0) return cudaSuccess;
    dim3 grid((W + 255) / 256, H);
    blur_kernel<<<grid, 256, 0, stream>>>(reinterpret_cast<const float4 *>(src), reinterpret_cast<float4 *>(dst), W, H);
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
