// maze_scene.cpp — Kruskal maze, wall run-length extraction and scene assembly.
// Restates reference src/main.rs:328-352 (TreeBuilder), :357-396 (maze), :397-438 (walls), :443-586 (planes,
// materials, emissions) in C++, generalised from the hard-wired 10x10 maze to n x n:
//   half = 10*(n/2) in f32 exactly as `-10.0 * (height as f32 / 2.0)` (:452), outer box +-half with side 10n,
//   entry light z = -half + 0.1 (== -49.9f at n = 10, :560), everything else literal.
// All arithmetic is f32 with one rounding per operation (compile with -ffp-contract=off).
#include "host_surface.h"

namespace mmh {

namespace {

// main.rs:328-352 — parent-pointer forest.  The reference walks to the root without rank or path compression (minutes of
// chain walking at n = 256); which node is the root never reaches the output — Kruskal only asks whether two cells are in the
// same set (main.rs:388) — so the walk halves the path as it goes: the same accept / reject sequence, the same maze.
struct TreeBuilder {
    std::vector<int64_t> nodes;   // -1 == None
    void new_node() { nodes.push_back(-1); }
    size_t get_root(size_t index) {
        while (nodes[index] >= 0) {
            const size_t parent = (size_t)nodes[index];
            if (nodes[parent] >= 0) nodes[index] = nodes[parent];   // path halving
            index = parent;
        }
        return index;
    }
    bool connected(size_t left, size_t right) { return get_root(left) == get_root(right); }
    void connect(size_t parent, size_t child) {
        size_t root = get_root(child);
        nodes[root] = (int64_t)parent;
    }
};

struct Edge { uint32_t x, y; bool up; };

inline mm_float3 f3(float x, float y, float z) { mm_float3 r = {x, y, z}; return r; }
inline mm_float4 f4(float x, float y, float z, float w) { mm_float4 r = {x, y, z, w}; return r; }

inline void push_plane(std::vector<mm_plane> &planes, mm_float3 origin, mm_float3 side1, mm_float3 side2, mm_float3 color) {
    mm_plane p;   // Plane::new(origin, side1, side2, color): v = side1, u = side2 (main.rs:61-68)
    p.origin = origin; p.v = side1; p.u = side2; p.color = color;
    planes.push_back(p);
}

}  // namespace

void build_maze(uint32_t n, StdRng &rng, std::vector<uint8_t> &grid) {
    TreeBuilder builder;
    std::vector<Edge> edges;
    std::vector<size_t> sets((size_t)n * n);
    grid.assign((size_t)n * n, 0);
    // main.rs:364-379: y outer, x inner; the `up` edge is pushed before the `left` edge.
    for (uint32_t y = 0; y < n; y++) {
        for (uint32_t x = 0; x < n; x++) {
            if (y != 0) edges.push_back({x, y, true});
            if (x != 0) edges.push_back({x, y, false});
            sets[(size_t)y * n + x] = builder.nodes.size();
            builder.new_node();
        }
    }
    // main.rs:382 edges.shuffle(&mut rng): rand 0.8.5 SliceRandom::shuffle (Fisher-Yates from the top).
    for (size_t i = edges.size(); i-- > 1;) {
        size_t j = rng.gen_range_u32(0, (uint32_t)(i + 1));
        Edge t = edges[i]; edges[i] = edges[j]; edges[j] = t;
    }
    // main.rs:384-396
    for (const Edge &e : edges) {
        uint32_t nx = e.up ? e.x : e.x - 1, ny = e.up ? e.y - 1 : e.y;
        size_t a = sets[(size_t)e.y * n + e.x], b = sets[(size_t)ny * n + nx];
        if (!builder.connected(a, b)) {
            builder.connect(a, b);
            if (e.up) {
                grid[(size_t)e.y * n + e.x] |= 1;
                grid[(size_t)ny * n + nx] |= 2;
            } else {
                grid[(size_t)e.y * n + e.x] |= 4;
                grid[(size_t)ny * n + nx] |= 8;
            }
        }
    }
}

void extract_walls(uint32_t n, const std::vector<uint8_t> &grid, std::vector<Wall> &vert, std::vector<Wall> &hori) {
    auto g = [&](uint32_t y, uint32_t x) { return grid[(size_t)y * n + x]; };
    vert.clear();
    hori.clear();
    // main.rs:397-417.  The trailing run is pushed even when its length is 0 (:416): zero-area planes exist.
    for (uint32_t x = 0; x < n; x++) {
        uint32_t wall_start = 0, wall_height = 0;
        for (uint32_t y = 0; y < n; y++) {
            if (x == 0) {
                wall_height += 1;
                continue;
            } else if ((g(y, x) & 4) == 0 && (g(y, x - 1) & 8) == 0) {
                wall_height += 1;
            } else {
                if (wall_height > 0) vert.push_back({(float)x, (float)wall_start, (float)wall_height});
                wall_height = 0;
                wall_start = y + 1;
            }
        }
        vert.push_back({(float)x, (float)wall_start, (float)wall_height});
    }
    // main.rs:419-438
    for (uint32_t y = 0; y < n; y++) {
        uint32_t wall_start = 0, wall_length = 0;
        for (uint32_t x = 0; x < n; x++) {
            if (y == 0) {
                wall_length += 1;
                continue;
            } else if ((g(y, x) & 1) == 0 && (g(y - 1, x) & 2) == 0) {
                wall_length += 1;
            } else {
                if (wall_length > 0) hori.push_back({(float)y, (float)wall_start, (float)wall_length});
                wall_length = 0;
                wall_start = x + 1;
            }
        }
        hori.push_back({(float)y, (float)wall_start, (float)wall_length});
    }
}

void assemble_scene(uint32_t n, const std::vector<Wall> &vert, const std::vector<Wall> &hori, StdRng &rng,
                    std::vector<mm_plane> &planes, std::vector<uint8_t> &materials, std::vector<mm_float4> &emissions) {
    planes.clear();
    materials.clear();
    emissions.clear();
    const mm_float3 wall_color = f3(0.3f, 0.35f, 0.4f);             // main.rs:447
    const float base = -10.0f * ((float)n / 2.0f);                  // main.rs:452 `-10.0 * (height as f32 / 2.0)`
    const float half = 10.0f * ((float)n / 2.0f);                   // 50.0 at n = 10
    const float side = 10.0f * (float)n;                            // 100.0 at n = 10

    for (const Wall &w : vert) {                                    // main.rs:449-481
        push_plane(planes, f3(base + (w.line * 10.0f), 2.0f, base + (w.start * 10.0f)),
                   f3(0.0f, 0.0f, w.len * 10.0f), f3(0.0f, -10.0f, 0.0f), wall_color);
        materials.push_back(rng.gen_f32() < 0.85f ? 0 : 1);
        emissions.push_back(f4(1.0f, 0.0f, 0.0f, 0.0f));
        if (w.len <= 2.0f && rng.gen_f32() < 0.3f) {                // short-circuit: draw only when len <= 2
            push_plane(planes, f3(base + (w.line * 10.0f) + 0.1f, 2.0f, base + (w.start * 10.0f)),
                       f3(0.0f, 0.0f, 9.9f), f3(0.0f, -6.0f, 0.0f), wall_color);
            materials.push_back(0);
            emissions.push_back(f4(1.0f, 0.8f, 0.3f, 2.0f));
        }
    }
    for (const Wall &w : hori) {                                    // main.rs:483-515
        push_plane(planes, f3(base + (w.start * 10.0f), 2.0f, base + (w.line * 10.0f)),
                   f3(w.len * 10.0f, 0.0f, 0.0f), f3(0.0f, -10.0f, 0.0f), wall_color);
        materials.push_back(rng.gen_f32() < 0.90f ? 0 : 1);
        emissions.push_back(f4(1.0f, 0.0f, 0.0f, 0.0f));
        if (w.len <= 2.0f && rng.gen_f32() < 0.3f) {
            push_plane(planes, f3(base + (w.start * 10.0f), 2.0f, base + (w.line * 10.0f) + 0.1f),
                       f3(9.9f, 0.0f, 0.0f), f3(0.0f, -6.0f, 0.0f), wall_color);
            materials.push_back(0);
            emissions.push_back(f4(1.0f, 0.8f, 0.3f, 2.0f));
        }
    }
    // main.rs:517-556: four outer walls, then the floor.
    push_plane(planes, f3(-half, 2.0f, -half), f3(0.0f, -20.0f, 0.0f), f3(side, 0.0f, 0.0f), wall_color);
    materials.push_back(0); emissions.push_back(f4(1.0f, 1.0f, 1.0f, 0.0f));
    push_plane(planes, f3(-half, 2.0f, half), f3(side, 0.0f, 0.0f), f3(0.0f, -20.0f, 0.0f), wall_color);
    materials.push_back(0); emissions.push_back(f4(1.0f, 1.0f, 1.0f, 0.0f));
    push_plane(planes, f3(-half, 2.0f, -half), f3(0.0f, 0.0f, side), f3(0.0f, -20.0f, 0.0f), wall_color);
    materials.push_back(0); emissions.push_back(f4(1.0f, 1.0f, 1.0f, 0.0f));
    push_plane(planes, f3(half, 2.0f, -half), f3(0.0f, -20.0f, 0.0f), f3(0.0f, 0.0f, side), wall_color);
    materials.push_back(0); emissions.push_back(f4(1.0f, 1.0f, 1.0f, 0.0f));
    push_plane(planes, f3(-half, 2.0f, half), f3(0.0f, 0.0f, -side), f3(side, 0.0f, 0.0f), f3(0.4f, 0.45f, 0.3f));
    materials.push_back(0); emissions.push_back(f4(1.0f, 1.0f, 1.0f, 0.0f));
    // main.rs:559-566: entry light just inside the z = -half wall.
    push_plane(planes, f3(-5.0f, 2.0f, -half + 0.1f), f3(10.0f, 0.0f, 0.0f), f3(0.0f, -6.0f, 0.0f), f3(0.0f, 0.0f, 0.0f));
    materials.push_back(0); emissions.push_back(f4(1.0f, 0.8f, 0.3f, 2.0f));
    // main.rs:578-585: faintly emissive roof.
    push_plane(planes, f3(-half, -8.0f, half), f3(0.0f, 0.0f, -side), f3(side, 0.0f, 0.0f), f3(0.0f, 0.0f, 0.0f));
    materials.push_back(0); emissions.push_back(f4(1.0f, 0.8f, 0.3f, 0.02f));
}

}  // namespace mmh
