/* b2_bvh_build.h -- host-side CW8 builder interface (see b2_bvh_build.cpp). */
#ifndef B2_BVH_BUILD_H
#define B2_BVH_BUILD_H

#include "b2_bvh.h"

#include <cstddef>
#include <string>
#include <vector>

struct B2BvhStats
{
    size_t node_count = 0;
    size_t triangle_count = 0;
    size_t bytes = 0;
    int max_depth = 0;
    double sah_cost = 0.0; /* sum of wide-node surface areas relative to the root's */
};

struct B2BvhHost
{
    std::vector<B2Node> nodes;
    std::vector<B2Tri> tris;
    B2BvhStats stats;
};

/* indices: 3 per triangle; positions: 3 floats per vertex. Returns false and fills `error` on invalid input. */
/* max_leaf: triangles per leaf slot, 1..3; 0 = chosen from the scene size (see b2_bvh_build.cpp) */
bool b2_build_cw8(const int* indices, size_t n_tri, const float* positions, size_t n_vert, B2BvhHost& out, std::string& error, int max_leaf = 0);

#endif
