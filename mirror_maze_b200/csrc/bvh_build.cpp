// bvh_build.cpp — the reference's top-down BVH2 builder (reference src/main.rs:91-212, 247-263).
//
// build_bvh_literal follows the Rust line by line: exhaustive SAH over every primitive centroid on three
// axes (O(count^2) per node, :118-129 + :180-211), `cost <= best_cost` keeps the LAST minimum (:123), leaf when
// best_cost > count*area (:130-135), Hoare-style partition on centre < split (:141-157), degenerate split ->
// leaf (:158-161), children pushed adjacently before either is subdivided (:162-168), interior marked with
// tri_count = 0 (:176).  Node order and leaf contents decide traversal order on the device, which decides
// which of two coplanar rects wins a tie (shaders.metal:63 strict `a < beam.t`), so they are part of parity.
//
// build_bvh_fast evaluates the same fp32 cost of every candidate with a sorted sweep
// (prefix/suffix boxes; min/max are exact and order-independent, counts are integers) and then applies the
// reference's scan order and `<=` rule, so it emits the identical node and index arrays in O(n log n) per node.
// tests/test_host_surface.py checks literal == fast array-for-array.
#include <algorithm>
#include "host_surface.h"

namespace mmh {

namespace {

struct Box {
    float mn[3], mx[3];
    Box() { for (int a = 0; a < 3; a++) { mn[a] = 1e30f; mx[a] = -1e30f; } }   // aabb::default (main.rs:220-227)
    inline void grow(const float p[3]) {                                        // main.rs:229-232
        for (int a = 0; a < 3; a++) { mn[a] = p[a] < mn[a] ? p[a] : mn[a]; mx[a] = p[a] > mx[a] ? p[a] : mx[a]; }
    }
    inline void merge(const Box &o) {
        for (int a = 0; a < 3; a++) { mn[a] = o.mn[a] < mn[a] ? o.mn[a] : mn[a]; mx[a] = o.mx[a] > mx[a] ? o.mx[a] : mx[a]; }
    }
    inline float area() const {                                                 // main.rs:233-236
        float e0 = mx[0] - mn[0], e1 = mx[1] - mn[1], e2 = mx[2] - mn[2];
        return e0 * e1 + e1 * e2 + e2 * e0;
    }
};

struct Prim {
    float corner[3][3];   // origin, origin+u, origin+v  (main.rs:95-97)
    float center[3];      // origin + (u+v)*0.5          (main.rs:69-71)
    Box box;
};

inline void grow_prim(Box &b, const Prim &p) { b.grow(p.corner[0]); b.grow(p.corner[1]); b.grow(p.corner[2]); }

std::vector<Prim> make_prims(const std::vector<mm_plane> &planes) {
    std::vector<Prim> prims(planes.size());
    for (size_t i = 0; i < planes.size(); i++) {
        const mm_plane &pl = planes[i];
        const float o[3] = {pl.origin.x, pl.origin.y, pl.origin.z};
        const float u[3] = {pl.u.x, pl.u.y, pl.u.z};
        const float v[3] = {pl.v.x, pl.v.y, pl.v.z};
        Prim &p = prims[i];
        for (int a = 0; a < 3; a++) {
            p.corner[0][a] = o[a];
            p.corner[1][a] = o[a] + u[a];
            p.corner[2][a] = o[a] + v[a];
            p.center[a] = o[a] + (u[a] + v[a]) * 0.5f;
        }
        grow_prim(p.box, p);
    }
    return prims;
}

struct Builder {
    const std::vector<Prim> &prims;
    std::vector<mm_bvh_node> &nodes;
    std::vector<uint32_t> &idx;
    bool fast;
    // scratch for the sweep
    std::vector<uint32_t> order;
    std::vector<Box> suffix;
    std::vector<float> cost_of;

    mm_bvh_node new_node(uint32_t left_first, uint32_t tri_count) {             // BVHNode::new (main.rs:83-90)
        mm_bvh_node nd;
        nd.aabb_min = {1e30f, 1e30f, 1e30f};
        nd.aabb_max = {-1e30f, -1e30f, -1e30f};
        nd.left_first = left_first;
        nd.tri_count = tri_count;
        return nd;
    }
    void update_bounds(mm_bvh_node &nd) {                                       // main.rs:91-101
        Box b;
        for (uint32_t i = nd.left_first; i < nd.left_first + nd.tri_count; i++) grow_prim(b, prims[idx[i]]);
        nd.aabb_min = {b.mn[0], b.mn[1], b.mn[2]};
        nd.aabb_max = {b.mx[0], b.mx[1], b.mx[2]};
    }
    float eval_sah(const mm_bvh_node &nd, int axis, float pos) {                // main.rs:180-211
        Box lb, rb;
        uint32_t lc = 0, rc = 0;
        for (uint32_t i = nd.left_first; i < nd.left_first + nd.tri_count; i++) {
            const Prim &p = prims[idx[i]];
            if (p.center[axis] < pos) { lc++; grow_prim(lb, p); } else { rc++; grow_prim(rb, p); }
        }
        float cost = (float)lc * lb.area() + (float)rc * rb.area();
        return cost > 0.0f ? cost : 1e30f;                                      // NaN (0*inf) falls to 1e30
    }
    // Costs of all candidates on one axis, candidate k = centre of prim idx[first+k]; identical values to eval_sah.
    void sweep_costs(const mm_bvh_node &nd, int axis) {
        const uint32_t first = nd.left_first, count = nd.tri_count;
        order.resize(count);
        for (uint32_t k = 0; k < count; k++) order[k] = k;
        std::sort(order.begin(), order.end(), [&](uint32_t a, uint32_t b) {
            return prims[idx[first + a]].center[axis] < prims[idx[first + b]].center[axis];
        });
        suffix.assign(count + 1, Box());
        for (uint32_t k = count; k-- > 0;) {
            suffix[k] = suffix[k + 1];
            suffix[k].merge(prims[idx[first + order[k]]].box);
        }
        cost_of.resize(count);
        Box prefix;
        uint32_t k = 0;
        while (k < count) {
            // group of equal centres: all share left set = sorted[0..k)
            float c = prims[idx[first + order[k]]].center[axis];
            uint32_t e = k;
            while (e < count && !(c < prims[idx[first + order[e]]].center[axis])) e++;
            float cost = (float)k * prefix.area() + (float)(count - k) * suffix[k].area();
            cost = cost > 0.0f ? cost : 1e30f;
            for (uint32_t m = k; m < e; m++) {
                cost_of[order[m]] = cost;
                prefix.merge(prims[idx[first + order[m]]].box);
            }
            k = e;
        }
    }
    void subdivide(uint32_t self) {                                             // main.rs:102-179
        mm_bvh_node nd = nodes[self];
        if (nd.tri_count == 1) return;
        float best_pos = 0.0f, best_cost = 1e30f;
        int best_axis = 6;
        for (int axis = 0; axis <= 2; axis++) {
            if (fast) sweep_costs(nd, axis);
            for (uint32_t i = nd.left_first; i < nd.left_first + nd.tri_count; i++) {
                float candidate = prims[idx[i]].center[axis];
                float cost = fast ? cost_of[i - nd.left_first] : eval_sah(nd, axis, candidate);
                if (cost <= best_cost) { best_cost = cost; best_pos = candidate; best_axis = axis; }
            }
        }
        float d0 = nd.aabb_max.x - nd.aabb_min.x, d1 = nd.aabb_max.y - nd.aabb_min.y, d2 = nd.aabb_max.z - nd.aabb_min.z;
        float area = d0 * d1 + d1 * d2 + d2 * d0;
        float parent_cost = (float)nd.tri_count * area;
        if (best_cost > parent_cost) return;
        if (best_axis > 2) return;   // unreachable (every cost <= 1e30); the Rust would index out of bounds
        int axis = best_axis;
        float split_pos = best_pos;
        int64_t i = nd.left_first, j = i + (int64_t)nd.tri_count - 1;
        while (i <= j) {
            if (prims[idx[(size_t)i]].center[axis] < split_pos) {
                i++;
            } else {
                std::swap(idx[(size_t)i], idx[(size_t)j]);
                j--;
            }
        }
        uint32_t left_count = (uint32_t)i - nd.left_first;
        if (left_count == 0 || left_count == nd.tri_count) return;
        mm_bvh_node left = new_node(nd.left_first, left_count);
        update_bounds(left);
        uint32_t li = (uint32_t)nodes.size();
        nodes.push_back(left);
        mm_bvh_node right = new_node((uint32_t)i, nd.tri_count - left_count);
        update_bounds(right);
        nodes.push_back(right);
        subdivide(li);
        subdivide(li + 1);
        nodes[self].left_first = li;
        nodes[self].tri_count = 0;
    }
};

void build(const std::vector<mm_plane> &planes, std::vector<mm_bvh_node> &nodes, std::vector<uint32_t> &indices, bool fast) {
    const size_t n = planes.size();
    std::vector<Prim> prims = make_prims(planes);
    nodes.clear();
    nodes.reserve(n ? 2 * n - 1 : 0);
    indices.resize(n);
    for (size_t i = 0; i < n; i++) indices[i] = (uint32_t)i;
    if (n == 0) return;
    Builder b{prims, nodes, indices, fast, {}, {}, {}};
    mm_bvh_node root = b.new_node(0, (uint32_t)n);
    b.update_bounds(root);
    nodes.push_back(root);
    b.subdivide(0);
}

}  // namespace

void build_bvh_literal(const std::vector<mm_plane> &planes, std::vector<mm_bvh_node> &nodes, std::vector<uint32_t> &indices) {
    build(planes, nodes, indices, false);
}
void build_bvh_fast(const std::vector<mm_plane> &planes, std::vector<mm_bvh_node> &nodes, std::vector<uint32_t> &indices) {
    build(planes, nodes, indices, true);
}

bool bvh_stats(const mm_bvh_node *nodes, uint32_t n_nodes, uint32_t n_planes, uint32_t *depth_out, uint32_t *max_leaf_out,
               std::vector<uint8_t> *reachable) {
    if (n_nodes == 0) return false;
    std::vector<uint8_t> seen(n_nodes, 0);
    std::vector<std::pair<uint32_t, uint32_t>> stack;   // (node, depth)
    stack.push_back({0u, 1u});
    uint32_t depth = 0, max_leaf = 0;
    while (!stack.empty()) {
        auto [i, d] = stack.back();
        stack.pop_back();
        if (i >= n_nodes || seen[i]) return false;      // out of range, shared child or cycle
        seen[i] = 1;
        if (d > depth) depth = d;
        const mm_bvh_node &nd = nodes[i];
        if (nd.tri_count > 0) {
            if ((uint64_t)nd.left_first + nd.tri_count > n_planes) return false;
            if (nd.tri_count > max_leaf) max_leaf = nd.tri_count;
        } else {
            if ((uint64_t)nd.left_first + 1 >= n_nodes) return false;
            stack.push_back({nd.left_first, d + 1});
            stack.push_back({nd.left_first + 1, d + 1});
        }
    }
    if (depth_out) *depth_out = depth;
    if (max_leaf_out) *max_leaf_out = max_leaf;
    if (reachable) reachable->swap(seen);
    return true;
}

}  // namespace mmh
