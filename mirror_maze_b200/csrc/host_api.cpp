// host_api.cpp — extern "C" face of the kept host surface (no device work here).
// Camera math follows reference src/maths.rs:139-178; chunk order follows src/main.rs:293-302 without the
// non-deterministic thread_rng shuffle (:303-305); the start-of-run uniform follows src/main.rs:599-602,732-755.
#include <cmath>
#include <cstring>
#include <new>
#include "host_surface.h"

using namespace mmh;

struct mm_scene { Scene s; };
struct mm_stdrng { StdRng r; explicit mm_stdrng(uint64_t seed) : r(seed) {} };
struct mm_bag {
    float w, h; uint32_t chunk;
    StdRng rng;
    std::vector<mm_chunk> original, pixels;
    mm_bag(float w_, float h_, uint32_t c, uint64_t seed) : w(w_), h(h_), chunk(c), rng(seed) {}
    void gen() {                                       // gen_pixels (main.rs:293-307)
        uint32_t n = mm_gen_chunks(w, h, chunk, nullptr, 0);
        original.resize(n);
        mm_gen_chunks(w, h, chunk, original.data(), n);
        for (size_t i = original.size(); i-- > 1;) {   // out_pixels.shuffle(&mut rng)
            size_t j = rng.gen_range_u32(0, (uint32_t)(i + 1));
            mm_chunk t = original[i]; original[i] = original[j]; original[j] = t;
        }
    }
};

namespace {

inline float magnitude(mm_float3 v) {                     // maths.rs:21-23 (powf(2.0) == x*x)
    return std::sqrt(v.x * v.x + v.y * v.y + v.z * v.z);
}
inline mm_float3 normalized(mm_float3 v) {                // maths.rs:24-26
    float m = magnitude(v);
    mm_float3 r = {v.x / m, v.y / m, v.z / m};
    return r;
}
inline mm_float3 cross(mm_float3 a, mm_float3 b) {        // maths.rs:130-136
    mm_float3 r = {a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x};
    return r;
}
inline float dot3(mm_float3 a, mm_float3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }   // maths.rs:105-107

inline mm_float4 quat_dot(mm_float4 q1, mm_float4 q2) {   // maths.rs:169-173 (host grouping, not the device's)
    mm_float3 a = {q1.x, q1.y, q1.z}, b = {q2.x, q2.y, q2.z};
    float s = q1.w * q2.w - dot3(a, b);
    mm_float3 c = cross(a, b);
    mm_float4 r = {c.x + (b.x * q1.w + a.x * q2.w), c.y + (b.y * q1.w + a.y * q2.w), c.z + (b.z * q1.w + a.z * q2.w), s};
    return r;
}

}  // namespace

extern "C" {

const char *mm_version(void) { return "mirror-maze-b200 0.1 (sm_100a)"; }

int mm_scene_build(uint32_t maze_n, uint64_t seed, int fast_bvh, mm_scene **out) {
    if (!out || maze_n < 1 || maze_n > 4096) return MM_ERR_INVALID;
    mm_scene *sc = new (std::nothrow) mm_scene();
    if (!sc) return MM_ERR_NOMEM;
    try {
        Scene &s = sc->s;
        s.n = maze_n;
        StdRng rng(seed);                                                    // main.rs:381
        build_maze(maze_n, rng, s.grid);                                     // main.rs:357-396
        extract_walls(maze_n, s.grid, s.vert_walls, s.hori_walls);           // main.rs:397-438
        assemble_scene(maze_n, s.vert_walls, s.hori_walls, rng, s.planes, s.materials, s.emissions);   // :443-586
        if (fast_bvh) build_bvh_fast(s.planes, s.nodes, s.indices);          // main.rs:588
        else build_bvh_literal(s.planes, s.nodes, s.indices);
    } catch (...) {
        delete sc;
        return MM_ERR_NOMEM;
    }
    *out = sc;
    return MM_OK;
}
int mm_scene_free(mm_scene *s) { delete s; return MM_OK; }
uint32_t mm_scene_n_planes(const mm_scene *s) { return s ? (uint32_t)s->s.planes.size() : 0; }
uint32_t mm_scene_n_nodes(const mm_scene *s) { return s ? (uint32_t)s->s.nodes.size() : 0; }
const mm_plane *mm_scene_planes(const mm_scene *s) { return s ? s->s.planes.data() : nullptr; }
const mm_bvh_node *mm_scene_nodes(const mm_scene *s) { return s ? s->s.nodes.data() : nullptr; }
const uint32_t *mm_scene_indices(const mm_scene *s) { return s ? s->s.indices.data() : nullptr; }
const uint8_t *mm_scene_materials(const mm_scene *s) { return s ? s->s.materials.data() : nullptr; }
const mm_float4 *mm_scene_emissions(const mm_scene *s) { return s ? s->s.emissions.data() : nullptr; }
const uint8_t *mm_scene_grid(const mm_scene *s) { return s ? s->s.grid.data() : nullptr; }
uint32_t mm_scene_n_vert_walls(const mm_scene *s) { return s ? (uint32_t)s->s.vert_walls.size() : 0; }
uint32_t mm_scene_n_hori_walls(const mm_scene *s) { return s ? (uint32_t)s->s.hori_walls.size() : 0; }
const float *mm_scene_vert_walls(const mm_scene *s) { return s ? &s->s.vert_walls.data()->line : nullptr; }
const float *mm_scene_hori_walls(const mm_scene *s) { return s ? &s->s.hori_walls.data()->line : nullptr; }

int mm_build_bvh(const mm_plane *planes, uint32_t n, int fast, mm_bvh_node *nodes, uint32_t *n_nodes_out, uint32_t *indices) {
    if (!planes || !nodes || !indices || !n_nodes_out || n == 0) return MM_ERR_INVALID;
    try {
        std::vector<mm_plane> p(planes, planes + n);
        std::vector<mm_bvh_node> nd;
        std::vector<uint32_t> ix;
        if (fast) build_bvh_fast(p, nd, ix); else build_bvh_literal(p, nd, ix);
        std::memcpy(nodes, nd.data(), nd.size() * sizeof(mm_bvh_node));
        std::memcpy(indices, ix.data(), ix.size() * sizeof(uint32_t));
        *n_nodes_out = (uint32_t)nd.size();
    } catch (...) {
        return MM_ERR_NOMEM;
    }
    return MM_OK;
}

int mm_stdrng_new(uint64_t seed, mm_stdrng **out) {
    if (!out) return MM_ERR_INVALID;
    *out = new (std::nothrow) mm_stdrng(seed);
    return *out ? MM_OK : MM_ERR_NOMEM;
}
int mm_stdrng_free(mm_stdrng *r) { delete r; return MM_OK; }
uint32_t mm_stdrng_next_u32(mm_stdrng *r) { return r->r.next_u32(); }
float mm_stdrng_gen_f32(mm_stdrng *r) { return r->r.gen_f32(); }
uint32_t mm_stdrng_gen_range_u32(mm_stdrng *r, uint32_t low, uint32_t high) { return r->r.gen_range_u32(low, high); }

int mm_chacha_block(const uint8_t key[32], uint64_t counter, uint64_t stream, int rounds, uint32_t out[16]) {
    if (!key || !out || rounds <= 0 || (rounds & 1)) return MM_ERR_INVALID;
    uint32_t k[8];
    for (int i = 0; i < 8; i++)
        k[i] = (uint32_t)key[4 * i] | ((uint32_t)key[4 * i + 1] << 8) | ((uint32_t)key[4 * i + 2] << 16) | ((uint32_t)key[4 * i + 3] << 24);
    chacha_block(k, counter, stream, rounds, out);
    return MM_OK;
}

mm_float4 mm_calculate_quaternion(mm_float3 dir) {       // maths.rs:139-156
    mm_float3 default_rotation = {0.0f, 0.0f, 1.0f};
    mm_float3 camera_rotation = normalized(dir);
    mm_float3 rotation_axis = cross(default_rotation, camera_rotation);
    mm_float3 axis_n = normalized(rotation_axis);
    float half_theta = std::asin(magnitude(rotation_axis)) / 2.0f;
    float s = std::sin(half_theta);
    mm_float4 q = {axis_n.x * s, axis_n.y * s, axis_n.z * s, std::cos(half_theta)};
    return q;
}

mm_float4 mm_update_quat_angle(mm_float4 q, float theta) {   // maths.rs:159-162
    float new_ratio = std::sin(theta) / std::sin(std::acos(q.w));
    mm_float4 r = {q.x * new_ratio, q.y * new_ratio, q.z * new_ratio, std::cos(theta)};
    return r;
}

mm_float3 mm_quat_mult(mm_float3 v, mm_float4 q) {       // maths.rs:175-178
    mm_float4 inv = {-q.x, -q.y, -q.z, q.w};
    mm_float4 vv = {v.x, v.y, v.z, 0.0f};
    mm_float4 r = quat_dot(quat_dot(inv, vv), q);
    mm_float3 o = {r.x, r.y, r.z};
    return o;
}

uint32_t mm_gen_chunks(float view_width, float view_height, uint32_t chunk_width, mm_chunk *out, uint32_t cap) {
    if (chunk_width == 0) return 0;
    uint32_t width = (uint32_t)view_width / chunk_width, height = (uint32_t)view_height / chunk_width;   // main.rs:294-295
    uint32_t k = 0;
    for (uint32_t i = 0; i < width; i++)                  // main.rs:298-302
        for (uint32_t j = 0; j < height; j++) {
            if (out && k < cap) { out[k].x = chunk_width * i; out[k].y = chunk_width * j; }
            k++;
        }
    return k;
}

int mm_default_uniform(uint32_t maze_n, float view_width, float view_height, uint32_t chunk_width, uint32_t time, mm_uniform *out) {
    if (!out || maze_n == 0 || !(view_width >= 1.0f) || !(view_height >= 1.0f) || chunk_width == 0) return MM_ERR_INVALID;
    float viewport_height = 2.0f;                                                     // main.rs:732
    float viewport_width = viewport_height * (view_width / view_height);              // main.rs:733
    mm_float3 start_dir = {0.1f, 0.0f, 1.0f};                                         // main.rs:740
    // main.rs:735 Float3(-5.0, 0.0, -45.0) at n = 10  ->  (-5, 0, -half + 5)
    out->cam.camera_center = {-5.0f, 0.0f, -10.0f * ((float)maze_n / 2.0f) + 5.0f};
    out->cam.focal_length = 1.0f;                                                     // main.rs:736
    out->cam.rotation = mm_calculate_quaternion(start_dir);
    out->cam.viewport = {viewport_width, viewport_height};
    out->view_width = view_width;
    out->view_height = view_height;
    out->chunk_width = chunk_width;
    out->time = time;
    return MM_OK;
}

int mm_bag_new(float view_width, float view_height, uint32_t chunk_width, uint64_t seed, mm_bag **out) {
    if (!out || chunk_width == 0 || !(view_width >= 1.0f) || !(view_height >= 1.0f)) return MM_ERR_INVALID;
    mm_bag *b = new (std::nothrow) mm_bag(view_width, view_height, chunk_width, seed);
    if (!b) return MM_ERR_NOMEM;
    try {
        b->gen();
        b->pixels = b->original;                         // main.rs:713-714
    } catch (...) { delete b; return MM_ERR_NOMEM; }
    *out = b;
    return MM_OK;
}
int mm_bag_free(mm_bag *bag) { delete bag; return MM_OK; }
uint32_t mm_bag_size(const mm_bag *bag) { return bag ? (uint32_t)bag->pixels.size() : 0; }
int mm_bag_next(mm_bag *bag, uint32_t n, mm_chunk *out) {    // random_pixels (main.rs:309-326)
    if (!bag || !out || bag->original.empty()) return MM_ERR_INVALID;
    try {
        for (uint32_t i = 0; i < n; i++) {
            if (bag->pixels.empty()) bag->pixels = bag->original;   // pixels.append(&mut original.clone())
            out[i] = bag->pixels.back();
            bag->pixels.pop_back();
        }
    } catch (...) { return MM_ERR_NOMEM; }
    return MM_OK;
}
int mm_bag_reshuffle(mm_bag *bag) {                          // main.rs:838-839
    if (!bag) return MM_ERR_INVALID;
    try {
        bag->gen();
        bag->pixels = bag->original;
    } catch (...) { return MM_ERR_NOMEM; }
    return MM_OK;
}

int mm_move_camera(const mm_bvh_node *nodes, uint32_t n_nodes, mm_float3 center, mm_float4 quat, const uint16_t *keys,
                   uint32_t n_keys, float fps, mm_float3 *out_center) {
    if (!nodes || !out_center || (n_keys && !keys) || !(fps > 0.0f)) return MM_ERR_INVALID;
    const mm_float3 prev = center;
    const float step = 5.0f / fps;
    for (uint32_t i = 0; i < n_keys; i++) {                  // main.rs:787-815
        mm_float3 dx = {step, 0.0f, 0.0f}, dz = {0.0f, 0.0f, step}, m;
        switch (keys[i]) {
            case 0: m = mm_quat_mult(dx, quat); center = {center.x - m.x, center.y - m.y, center.z - m.z}; break;
            case 1: m = mm_quat_mult(dz, quat); center = {center.x - m.x, center.y - m.y, center.z - m.z}; break;
            case 2: m = mm_quat_mult(dx, quat); center = {center.x + m.x, center.y + m.y, center.z + m.z}; break;
            case 13: m = mm_quat_mult(dz, quat); center = {center.x + m.x, center.y + m.y, center.z + m.z}; break;
            default: break;
        }
    }
    const mm_float3 diag = {0.5f, 0.2f, 0.5f};               // main.rs:738
    mm_float3 bmin = {center.x - diag.x, center.y - diag.y, center.z - diag.z};
    mm_float3 bmax = {center.x + diag.x, center.y + diag.y, center.z + diag.z};
    int blocked = mm_check_collision(nodes, n_nodes, bmin, bmax) >= 0 ? 1 : 0;   // main.rs:817-826
    *out_center = blocked ? prev : center;
    return blocked;
}

int mm_check_collision(const mm_bvh_node *nodes, uint32_t n_nodes, mm_float3 bmin, mm_float3 bmax) {
    // main.rs:265-291, iteratively and with the same visiting order (left subtree first).  Like the reference,
    // only tri_count == 1 nodes are treated as leaves; a multi-plane leaf has tri_count > 1 and left_first
    // pointing into the index array, which the reference would mis-walk as a node index —
    // here such a leaf is tested against its own box instead of being followed.
    if (!nodes || n_nodes == 0) return -1;
    auto overlap = [&](const mm_bvh_node &nd) {           // aabb::intersect (main.rs:237-245), self = player box
        return bmin.x <= nd.aabb_max.x && bmax.x >= nd.aabb_min.x && bmin.y <= nd.aabb_max.y && bmax.y >= nd.aabb_min.y &&
               bmin.z <= nd.aabb_max.z && bmax.z >= nd.aabb_min.z;
    };
    std::vector<uint32_t> stack;
    stack.push_back(0);
    while (!stack.empty()) {
        uint32_t i = stack.back();
        stack.pop_back();
        if (i >= n_nodes) return -1;
        const mm_bvh_node &nd = nodes[i];
        if (nd.tri_count >= 1) {
            if (overlap(nd)) return (int)i;
            continue;
        }
        if (overlap(nd)) {
            stack.push_back(nd.left_first + 1);
            stack.push_back(nd.left_first);
        }
    }
    return -1;
}

}  // extern "C"
