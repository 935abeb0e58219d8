// host_surface.h — C++ restatement of the reference's host data surface (maze, walls, scene, BVH, camera).
// The reference is Rust (src/main.rs, src/maths.rs); no Rust toolchain exists on the build or GPU boxes, so
// the surface is kept byte-for-byte in C++ and exported through include/mirror_maze_cuda.h.
#pragma once
#include <cstdint>
#include <vector>
#include "../../include/mirror_maze_cuda.h"

static_assert(sizeof(mm_float2) == 8 && sizeof(mm_float3) == 12 && sizeof(mm_float4) == 16, "maths.rs layouts");
static_assert(sizeof(mm_plane) == 48, "Plane layout (main.rs:51-58)");
static_assert(sizeof(mm_bvh_node) == 32, "BVHNode layout (main.rs:74-81)");
static_assert(sizeof(mm_camera) == 40, "Camera layout (main.rs:32-39)");
static_assert(sizeof(mm_uniform) == 56, "Uniform layout (main.rs:41-49)");
static_assert(sizeof(mm_chunk) == 8, "chunk = (u32, u32)");

namespace mmh {

void chacha_block(const uint32_t key[8], uint64_t counter, uint64_t stream, int rounds, uint32_t out[16]);

class StdRng {
public:
    explicit StdRng(uint64_t seed);
    uint32_t next_u32();
    float gen_f32();
    uint32_t gen_range_u32(uint32_t low, uint32_t high);   // [low, high)
private:
    uint32_t key_[8];
    uint64_t counter_;
    uint32_t buf_[16];
    int pos_;
};

struct Wall { float line, start, len; };   // (x|y, start, length) tuples of main.rs:409,416,431,437

struct Scene {
    uint32_t n = 0;                         // maze is n x n cells
    std::vector<uint8_t> grid;              // passage bits 1=N 2=S 4=W 8=E, [y*n + x]
    std::vector<Wall> vert_walls, hori_walls;
    std::vector<mm_plane> planes;
    std::vector<uint8_t> materials;         // Rust Vec<bool>
    std::vector<mm_float4> emissions;
    std::vector<mm_bvh_node> nodes;
    std::vector<uint32_t> indices;
};

// main.rs:357-396 Kruskal with StdRng; rng continues into build_scene_planes.
void build_maze(uint32_t n, StdRng &rng, std::vector<uint8_t> &grid);
// main.rs:397-438
void extract_walls(uint32_t n, const std::vector<uint8_t> &grid, std::vector<Wall> &vert, std::vector<Wall> &hori);
// main.rs:443-586
void assemble_scene(uint32_t n, const std::vector<Wall> &vert, const std::vector<Wall> &hori, StdRng &rng,
                    std::vector<mm_plane> &planes, std::vector<uint8_t> &materials, std::vector<mm_float4> &emissions);
// main.rs:247-263 (+ 91-212).  literal: the reference's O(n^2)-per-node exhaustive SAH.
// fast: sorted-sweep evaluation of the same fp32 costs with the same tie rule -> identical arrays.
void build_bvh_literal(const std::vector<mm_plane> &planes, std::vector<mm_bvh_node> &nodes, std::vector<uint32_t> &indices);
void build_bvh_fast(const std::vector<mm_plane> &planes, std::vector<mm_bvh_node> &nodes, std::vector<uint32_t> &indices);

// Structural facts used by upload-time validation.  Returns false on malformed trees.
// `reachable` (optional): n_nodes flags, 1 for every node the walk from the root visits.
bool bvh_stats(const mm_bvh_node *nodes, uint32_t n_nodes, uint32_t n_planes, uint32_t *depth, uint32_t *max_leaf,
               std::vector<uint8_t> *reachable = nullptr);

}  // namespace mmh
