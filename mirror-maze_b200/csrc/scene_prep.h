// scene_prep.h — host-side conversion of the reference buffers into the device layout (see scene_prep.cpp).
#pragma once
#include <string>
#include <vector>
#include "host_surface.h"
#include "render_kernel.cuh"

namespace mmk {

struct Prepared {
    std::vector<PairRec> pairs;
    std::vector<RectI> rects;
    std::vector<RectS> shade;
    uint32_t n_pairs = 0, root_link = 0, root_count = 0, depth = 0, max_leaf = 0;
    bool fast_ok = false;
};

int prepare_scene(const mm_plane *planes, uint32_t n_planes, const mm_bvh_node *nodes, uint32_t n_nodes,
                  const uint32_t *indices, const uint8_t *materials, const mm_float4 *emissions, Prepared &out,
                  std::string &err);

}  // namespace mmk
