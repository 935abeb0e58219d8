// scene_prep.cpp — turns the reference's buffers (rect[], bvh_node[], indices[], materials[], emissions[];
// reference src/shaders.metal:249-254) into the device layout of render_kernel.cuh.  Host code, compiled by g++
// with -ffp-contract=off: the per-rect constants are computed with exactly the operations ray_rect_intersect
// performs per call (shaders.metal:52,60-61: normalize(cross(v,u)), length(v), length(u)) and the shader performs
// per hit (:312 emissions.rgb * emissions.a), one IEEE rounding each, so the kernel sees the literal values.
#include <cmath>
#include <cstring>
#include "scene_prep.h"

namespace mmk {

namespace {
inline float dot3(const float a[3], const float b[3]) { return (a[0] * b[0] + a[1] * b[1]) + a[2] * b[2]; }
inline float as_float(uint32_t u) { float f; std::memcpy(&f, &u, 4); return f; }
inline bool coord_ok(float b) {
    float a = std::fabs(b);
    return a == 0.0f || (a >= 9.765625e-4f /*2^-10*/ && a <= 1073741824.0f /*2^30*/);
}
}  // namespace

int prepare_scene(const mm_plane *planes, uint32_t n_planes, const mm_bvh_node *nodes, uint32_t n_nodes,
                  const uint32_t *indices, const uint8_t *materials, const mm_float4 *emissions, Prepared &out,
                  std::string &err) {
    uint32_t depth = 0, max_leaf = 0;
    if (!mmh::bvh_stats(nodes, n_nodes, n_planes, &depth, &max_leaf)) {
        err = "malformed BVH: child or leaf range out of bounds, shared child or cycle";
        return MM_ERR_BVH;
    }
    if (depth > MM_MAX_STACK) {   // stack occupancy <= depth - 1
        err = "BVH depth " + std::to_string(depth) + " exceeds MM_MAX_STACK";
        return MM_ERR_BVH;
    }
    if (max_leaf > kMaxLeafCount || n_planes >= (1u << 24) || n_nodes >= (1u << 24)) {
        err = "scene exceeds packed-descriptor limits (leaf > 126 planes or >= 2^24 planes/nodes)";
        return MM_ERR_UNSUPPORTED;
    }
    for (uint32_t i = 0; i < n_planes; i++)
        if (indices[i] >= n_planes) { err = "index out of range"; return MM_ERR_BVH; }

    // pair ids: interior nodes in node-index order (the root, node 0, gets pair 0)
    std::vector<uint32_t> pair_id(n_nodes, 0xFFFFFFFFu);
    uint32_t n_pairs = 0;
    for (uint32_t i = 0; i < n_nodes; i++)
        if (nodes[i].tri_count == 0) pair_id[i] = n_pairs++;
    auto desc = [&](uint32_t c) -> uint32_t {
        const mm_bvh_node &nd = nodes[c];
        return nd.tri_count > 0 ? (kLeafBit | nd.left_first | (nd.tri_count << 24)) : pair_id[c] * (uint32_t)sizeof(PairRec);
    };
    out.pairs.assign(n_pairs ? n_pairs : 1, PairRec());
    std::memset(out.pairs.data(), 0, out.pairs.size() * sizeof(PairRec));
    bool fast_ok = true;
    for (uint32_t i = 0; i < n_nodes; i++) {
        if (nodes[i].tri_count != 0) continue;
        const mm_bvh_node &a = nodes[nodes[i].left_first], &b = nodes[nodes[i].left_first + 1];
        PairRec &p = out.pairs[pair_id[i]];
        for (int sy = 0; sy < 2; sy++)
            for (int sx = 0; sx < 2; sx++) {
                const mm_bvh_node *ch[2] = {&a, &b};
                for (int k = 0; k < 2; k++) {
                    const mm_float3 &mn = ch[k]->aabb_min, &mx = ch[k]->aabb_max;
                    p.ab[sx + 2 * sy][k] = make_float4(sx ? mx.x : mn.x, sy ? mx.y : mn.y, sx ? mn.x : mx.x, sy ? mn.y : mx.y);
                }
            }
        p.zl[0].z = make_float4(a.aabb_min.z, b.aabb_min.z, a.aabb_max.z, b.aabb_max.z);
        p.zl[1].z = make_float4(a.aabb_max.z, b.aabb_max.z, a.aabb_min.z, b.aabb_min.z);
        p.zl[0].link = p.zl[1].link = make_uint4(desc(nodes[i].left_first), desc(nodes[i].left_first + 1), 0u, 0u);
        const float c[12] = {a.aabb_min.x, a.aabb_min.y, a.aabb_min.z, a.aabb_max.x, a.aabb_max.y, a.aabb_max.z,
                             b.aabb_min.x, b.aabb_min.y, b.aabb_min.z, b.aabb_max.x, b.aabb_max.y, b.aabb_max.z};
        for (int k = 0; k < 12; k++) fast_ok = fast_ok && coord_ok(c[k]);
        // the travel-ordered fast path also needs min <= max on every axis (true for every box the builder emits)
        for (int k = 0; k < 3; k++) fast_ok = fast_ok && c[k] <= c[k + 3] && c[k + 6] <= c[k + 9];
    }
    out.n_pairs = n_pairs;
    out.root_link = nodes[0].tri_count > 0 ? nodes[0].left_first : 0u;
    out.root_count = nodes[0].tri_count;
    out.depth = depth;
    out.max_leaf = max_leaf;
    out.fast_ok = fast_ok;

    out.rects.resize(n_planes);
    out.shade.resize(n_planes);
    for (uint32_t s = 0; s < n_planes; s++) {
        const uint32_t id = indices[s];
        const mm_plane &m = planes[id];
        const float v[3] = {m.v.x, m.v.y, m.v.z}, u[3] = {m.u.x, m.u.y, m.u.z};
        const float c[3] = {v[1] * u[2] - v[2] * u[1], v[2] * u[0] - v[0] * u[2], v[0] * u[1] - v[1] * u[0]};   // cross(v,u)
        const float lc = std::sqrt(dot3(c, c));
        const float n[3] = {c[0] / lc, c[1] / lc, c[2] / lc};                                                   // normalize
        const float lv = std::sqrt(dot3(v, v)), lu = std::sqrt(dot3(u, u));
        RectI &r = out.rects[s];
        r.o_lv = make_float4(m.origin.x, m.origin.y, m.origin.z, lv);
        r.n_lu = make_float4(n[0], n[1], n[2], lu);
        r.v_id = make_float4(v[0], v[1], v[2], as_float(id));
        r.u_mat = make_float4(u[0], u[1], u[2], as_float(materials[id] ? 1u : 0u));
        const mm_float4 &e = emissions[id];
        RectS &sh = out.shade[s];
        sh.color = make_float4(m.color.x, m.color.y, m.color.z, 0.0f);
        sh.emitted = make_float4(e.x * e.w, e.y * e.w, e.z * e.w, 0.0f);
    }
    return MM_OK;
}

}  // namespace mmk
