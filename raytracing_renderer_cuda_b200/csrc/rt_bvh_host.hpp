// rt_bvh_host.hpp — host-side binned-SAH builder producing the flattened 64-byte
// two-child-box node array of rt_device.cuh (replaces the reference's device-side,
// single-thread, random-axis median split: bvh_node ctor, bvh.h:75-113).
#pragma once

#include <cstdint>
#include <vector>

namespace rth {

struct Box {
    float lo[3], hi[3];
};

struct NodeHost { // mirrors rtd::BvhNode (4 x float4)
    float lmin[3];
    int32_t left;
    float lmax[3];
    int32_t right;
    float rmin[3];
    float pad0;
    float rmax[3];
    float pad1;
};
static_assert(sizeof(NodeHost) == 64, "BVH node must be 64 bytes");

struct BvhStats {
    uint32_t depth = 0;
    float sah_cost = 0.f;
};

// One primitive per leaf; child < 0 encodes leaf ~prim.  boxes.size() >= 2.
void build_bvh_sah(const std::vector<Box>& boxes, std::vector<NodeHost>& nodes, BvhStats* stats);

// Sphere bounds as the reference computes them (sphere.h:142-146; moving: union of the
// boxes at c0 and c1, sphere.h:192-202), padded so the conservative slab test never
// rejects a ray the (rounded) sphere test accepts — see DESIGN.md "BVH == brute force".
Box sphere_box(const float c0[3], const float c1[3], float r);

} // namespace rth
