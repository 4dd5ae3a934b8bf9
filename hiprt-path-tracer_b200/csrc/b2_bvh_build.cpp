/*
 * b2_bvh_build.cpp -- host-side construction of the CW8 acceleration structure (layout: b2_bvh.h).
 *
 * Replaces `HIPRTGeometry::build_bvh` (/root/reference/src/HIPRT-Orochi/HIPRTScene.h:87-150), which hands the
 * mesh to HIPRT's GPU LBVH/PLOC/SBVH builders. Per the north star the B200 build does this on the host:
 *   1. binned-SAH binary BVH over triangle boxes (64 bins; leaves of <= 2 triangles, 1 on scenes of a million triangles and more), subtrees built in parallel;
 *   2. collapse to 8-wide nodes: a child whose whole subtree fits into the free slots is dissolved into them, otherwise the child
 *      with the largest surface area is opened;
 *   3. octant-ordered slot assignment, 8-bit quantisation of child boxes on a power-of-two grid anchored at the
 *      node's min corner (floor / ceil, i.e. conservative), breadth-first emission so inner children are contiguous.
 * Every input triangle is kept (degenerate ones included: the reference tests them too and they can never be hit).
 */
#include "b2_bvh_build.h"

#include <algorithm>
#include <atomic>
#include <cfloat>
#include <cmath>
#include <cstring>
#include <deque>
#include <numeric>

namespace
{
struct Box
{
    float lo[3], hi[3];
    void reset() { for (int a = 0; a < 3; a++) { lo[a] = FLT_MAX; hi[a] = -FLT_MAX; } }
    void grow(const float* p) { for (int a = 0; a < 3; a++) { lo[a] = std::min(lo[a], p[a]); hi[a] = std::max(hi[a], p[a]); } }
    void grow(const Box& b) { for (int a = 0; a < 3; a++) { lo[a] = std::min(lo[a], b.lo[a]); hi[a] = std::max(hi[a], b.hi[a]); } }
    float half_area() const
    {
        float dx = hi[0] - lo[0], dy = hi[1] - lo[1], dz = hi[2] - lo[2];
        if (dx < 0.0f) return 0.0f;
        return dx * dy + dy * dz + dz * dx;
    }
};

struct Bin2Node
{
    Box box;
    int left = -1, right = -1; /* children (inner) */
    int first = 0, count = 0;  /* range in the permuted primitive array (leaf) */
};

/* 64 bins: measured on the 10M-triangle scene, nodes visited per ray 28.7 (16 bins) -> 25.0 (32) -> 24.2 (64) -> 23.9 (128) */
constexpr int kBins = 64;
/* Triangles per leaf slot (<= 3, the CW8 leaf encoding). The traversal tests a leaf's triangles one after the other while
 * the lanes whose node had no leaf hit wait, so on big scenes smaller leaves pay although they make more nodes. Measured on
 * B200 (profiles/r01_ncu_megakernel_history.md): 10M triangles 3 / 2 / 1 per leaf = 25.2 / 27.3 / 28.9 Mpaths/s (nodes per
 * ray 24.2 / 24.7 / 25.6, triangles per ray 9.9 / 7.8 / 5.9); Cornell (8k triangles) 342.9 / 344.2 / 328.1. */
constexpr size_t kSmallLeafTriangleCount = 1u << 20;
#ifndef B2_COLLAPSE_FILL
#define B2_COLLAPSE_FILL 0 /* measured on the 10M-triangle scene: 3.03 M -> 2.0 M nodes but the same node visits per ray (the visits are upper-level nodes the ray grazes) and the same wf_trace time; off because a different tree resolves equal-t ties differently (frame checksum moved in the 8th digit) */
#endif

struct Builder
{
    const std::vector<Box>& prim_box;
    const std::vector<float>& centroid; /* 3 per prim */
    std::vector<int>& order;
    std::vector<Bin2Node> nodes;
    std::atomic<int> next_node{ 0 };

    int max_leaf;

    Builder(const std::vector<Box>& b, const std::vector<float>& c, std::vector<int>& o, int leaf) : prim_box(b), centroid(c), order(o), max_leaf(leaf) { nodes.resize(std::max<size_t>(1, 2 * o.size())); }

    int alloc() { return next_node.fetch_add(1); }

    void build(int node_index, int first, int count, int depth)
    {
        Bin2Node& n = nodes[node_index];
        n.box.reset();
        Box cbox;
        cbox.reset();
        for (int i = first; i < first + count; i++)
        {
            n.box.grow(prim_box[order[i]]);
            cbox.grow(&centroid[3 * (size_t)order[i]]);
        }
        if (count <= max_leaf)
        {
            n.first = first;
            n.count = count;
            return;
        }

        /* binned SAH over the three axes */
        int best_axis = -1, best_split = -1;
        float best_cost = FLT_MAX;
        for (int axis = 0; axis < 3; axis++)
        {
            float c0 = cbox.lo[axis], c1 = cbox.hi[axis];
            if (!(c1 > c0))
                continue;
            Box bins[kBins];
            int counts[kBins] = { 0 };
            for (int b = 0; b < kBins; b++) bins[b].reset();
            float scale = kBins / (c1 - c0);
            for (int i = first; i < first + count; i++)
            {
                int p = order[i];
                int b = std::min(kBins - 1, std::max(0, (int)((centroid[3 * (size_t)p + axis] - c0) * scale)));
                bins[b].grow(prim_box[p]);
                counts[b]++;
            }
            float right_area[kBins];
            int right_count[kBins];
            Box acc;
            acc.reset();
            int cnt = 0;
            for (int b = kBins - 1; b > 0; b--)
            {
                acc.grow(bins[b]);
                cnt += counts[b];
                right_area[b] = acc.half_area();
                right_count[b] = cnt;
            }
            acc.reset();
            cnt = 0;
            for (int b = 0; b < kBins - 1; b++)
            {
                acc.grow(bins[b]);
                cnt += counts[b];
                if (cnt == 0 || right_count[b + 1] == 0)
                    continue;
                float cost = acc.half_area() * cnt + right_area[b + 1] * right_count[b + 1];
                if (cost < best_cost) { best_cost = cost; best_axis = axis; best_split = b; }
            }
        }

        int mid;
        if (best_axis < 0)
        {
            mid = first + count / 2; /* all centroids coincide: split by index */
        }
        else
        {
            float c0 = cbox.lo[best_axis], scale = kBins / (cbox.hi[best_axis] - c0);
            auto it = std::partition(order.begin() + first, order.begin() + first + count, [&](int p) {
                int b = std::min(kBins - 1, std::max(0, (int)((centroid[3 * (size_t)p + best_axis] - c0) * scale)));
                return b <= best_split;
            });
            mid = (int)(it - order.begin());
            if (mid == first || mid == first + count)
                mid = first + count / 2;
        }

        int l = alloc(), r = alloc();
        nodes[node_index].left = l;
        nodes[node_index].right = r;
        int lc = mid - first, rc = count - lc;
        if (count > 1 << 16)
        {
#pragma omp task default(shared) firstprivate(l, first, lc, depth)
            build(l, first, lc, depth + 1);
#pragma omp task default(shared) firstprivate(r, mid, rc, depth)
            build(r, mid, rc, depth + 1);
#pragma omp taskwait
        }
        else
        {
            build(l, first, lc, depth + 1);
            build(r, mid, rc, depth + 1);
        }
    }
};

/* exponent e (unbiased) such that extent / 2^e <= 255, as small as possible */
int grid_exponent(float extent)
{
    if (!(extent > 0.0f))
        return -100;
    int e = (int)std::ceil(std::log2((double)extent / 255.0));
    while (std::ldexp((double)extent, -e) > 255.0) e++;
    while (e > -120 && std::ldexp((double)extent, -(e - 1)) <= 255.0) e--;
    return std::max(-120, std::min(120, e));
}
} // namespace

bool b2_build_cw8(const int* indices, size_t n_tri, const float* positions, size_t n_vert, B2BvhHost& out, std::string& error, int max_leaf)
{
    if (max_leaf <= 0)
        max_leaf = n_tri >= kSmallLeafTriangleCount ? 1 : 2;
    max_leaf = std::min(max_leaf, 3);
    out.nodes.clear();
    out.tris.clear();
    out.stats = B2BvhStats();
    if (n_tri == 0)
        return true;
    if (n_tri > 0x7FFFFFFFull / 3)
    {
        error = "too many triangles for 32-bit indexing";
        return false;
    }
    for (size_t i = 0; i < 3 * n_tri; i++)
        if (indices[i] < 0 || (size_t)indices[i] >= n_vert)
        {
            error = "triangle index out of range";
            return false;
        }

    std::vector<Box> prim_box(n_tri);
    std::vector<float> centroid(3 * n_tri);
    std::vector<int> order(n_tri);
    std::iota(order.begin(), order.end(), 0);
#pragma omp parallel for schedule(static)
    for (long long t = 0; t < (long long)n_tri; t++)
    {
        Box b;
        b.reset();
        for (int k = 0; k < 3; k++)
            b.grow(positions + 3 * (size_t)indices[3 * t + k]);
        prim_box[t] = b;
        /* same centroid notion as the reference octree uses for insertion (bbox centre, Triangle.cpp:8-11) */
        for (int a = 0; a < 3; a++)
            centroid[3 * t + a] = 0.5f * (b.lo[a] + b.hi[a]);
    }

    Builder builder(prim_box, centroid, order, max_leaf);
    int root = builder.alloc();
#pragma omp parallel
#pragma omp single
    builder.build(root, 0, (int)n_tri, 0);
    const std::vector<Bin2Node>& b2 = builder.nodes;

    /* leaves below every binary node (children are allocated after their parent, so a reverse sweep sees them first) */
    const int n_bin = builder.next_node.load();
    std::vector<int> leaf_count((size_t)n_bin);
    for (int i = n_bin - 1; i >= 0; i--)
        leaf_count[i] = b2[i].left < 0 ? 1 : leaf_count[b2[i].left] + leaf_count[b2[i].right];

    /* ---- collapse to 8-wide + emit ---- */
    struct Pending { int bin_node; uint32_t wide_index; int depth; };
    std::deque<Pending> queue;
    out.nodes.reserve(n_tri / 4 + 16);
    out.tris.reserve(n_tri);
    out.nodes.emplace_back();
    queue.push_back({ root, 0u, 1 });
    double sah = 0.0;
    const double root_area = std::max(1.0e-30, (double)b2[root].box.half_area());

    while (!queue.empty())
    {
        Pending job = queue.front();
        queue.pop_front();
        out.stats.max_depth = std::max(out.stats.max_depth, job.depth);

        /* gather up to 8 children */
        int kids[8];
        int n_kids = 0;
        const Bin2Node& top = b2[job.bin_node];
        if (top.left < 0)
            kids[n_kids++] = job.bin_node;
        else
        {
            kids[n_kids++] = top.left;
            kids[n_kids++] = top.right;
            while (n_kids < 8)
            {
                /* First choice: an inner child whose WHOLE subtree fits into the free slots (it disappears as a node: its leaves
                 * become slots of this one). Opening by area alone leaves the bottom of the tree full of 2- and 3-child nodes
                 * (10M triangles: 3.03 M nodes = 4.3 children per node); each of them costs a ray a dependent 80-byte fetch and
                 * the same 8-slot slab test as a full node. Otherwise: the child with the largest surface area. */
                int pick = -1;
                float pick_area = -1.0f;
                bool pick_fits = false;
                for (int i = 0; i < n_kids; i++)
                {
                    if (b2[kids[i]].left < 0)
                        continue;
                    const bool fits = B2_COLLAPSE_FILL && leaf_count[kids[i]] - 1 <= 8 - n_kids;
                    const float area = b2[kids[i]].box.half_area();
                    if ((fits && !pick_fits) || (fits == pick_fits && area > pick_area))
                    {
                        pick = i;
                        pick_area = area;
                        pick_fits = fits;
                    }
                }
                if (pick < 0)
                    break;
                int opened = kids[pick];
                kids[pick] = b2[opened].left;
                kids[n_kids++] = b2[opened].right;
            }
        }

        Box nbox;
        nbox.reset();
        for (int i = 0; i < n_kids; i++)
            nbox.grow(b2[kids[i]].box);
        sah += (double)nbox.half_area() / root_area;

        /* slot s must hold the child that is nearest for rays of octant s (x<0 -> 4, y<0 -> 2, z<0 -> 1) */
        int slot_of[8], kid_in_slot[8];
        for (int i = 0; i < 8; i++) { slot_of[i] = -1; kid_in_slot[i] = -1; }
        float cost[8][8];
        float ncx = 0.5f * (nbox.lo[0] + nbox.hi[0]), ncy = 0.5f * (nbox.lo[1] + nbox.hi[1]), ncz = 0.5f * (nbox.lo[2] + nbox.hi[2]);
        for (int i = 0; i < n_kids; i++)
        {
            const Box& kb = b2[kids[i]].box;
            float dx = 0.5f * (kb.lo[0] + kb.hi[0]) - ncx, dy = 0.5f * (kb.lo[1] + kb.hi[1]) - ncy, dz = 0.5f * (kb.lo[2] + kb.hi[2]) - ncz;
            for (int s = 0; s < 8; s++)
            {
                float sx = (s & 4) ? -1.0f : 1.0f, sy = (s & 2) ? -1.0f : 1.0f, sz = (s & 1) ? -1.0f : 1.0f;
                cost[i][s] = -(dx * sx + dy * sy + dz * sz); /* larger = nearer along that octant's diagonal */
            }
        }
        for (int round = 0; round < n_kids; round++)
        {
            int bi = -1, bs = -1;
            float bc = -FLT_MAX;
            for (int i = 0; i < n_kids; i++)
            {
                if (slot_of[i] >= 0) continue;
                for (int s = 0; s < 8; s++)
                    if (kid_in_slot[s] < 0 && cost[i][s] > bc) { bc = cost[i][s]; bi = i; bs = s; }
            }
            slot_of[bi] = bs;
            kid_in_slot[bs] = bi;
        }

        B2Node node;
        std::memset(&node, 0, sizeof(node));
        node.ox = nbox.lo[0];
        node.oy = nbox.lo[1];
        node.oz = nbox.lo[2];
        int e[3];
        for (int a = 0; a < 3; a++)
            e[a] = grid_exponent(nbox.hi[a] - nbox.lo[a]);
        uint32_t imask = 0;
        node.child_base = (uint32_t)out.nodes.size();
        node.tri_base = (uint32_t)out.tris.size();
        uint32_t tri_offset = 0;

        for (int s = 0; s < 8; s++)
        {
            int i = kid_in_slot[s];
            if (i < 0)
                continue;
            const Bin2Node& kid = b2[kids[i]];
            const float* origin = nbox.lo;
            uint8_t* qlo[3] = { node.qlo_x, node.qlo_y, node.qlo_z };
            uint8_t* qhi[3] = { node.qhi_x, node.qhi_y, node.qhi_z };
            for (int a = 0; a < 3; a++)
            {
                double lo = std::floor(std::ldexp((double)kid.box.lo[a] - (double)origin[a], -e[a]));
                double hi = std::ceil(std::ldexp((double)kid.box.hi[a] - (double)origin[a], -e[a]));
                qlo[a][s] = (uint8_t)std::max(0.0, std::min(255.0, lo));
                qhi[a][s] = (uint8_t)std::max(0.0, std::min(255.0, hi));
            }
            if (kid.left >= 0)
            {
                imask |= 1u << s;
                node.meta[s] = (uint8_t)((1u << 5) | (24u + (uint32_t)s));
                uint32_t child_index = (uint32_t)out.nodes.size();
                out.nodes.emplace_back();
                queue.push_back({ kids[i], child_index, job.depth + 1 });
            }
            else
            {
                uint32_t unary = kid.count == 1 ? 1u : (kid.count == 2 ? 3u : 7u);
                node.meta[s] = (uint8_t)((unary << 5) | tri_offset);
                for (int k = 0; k < kid.count; k++)
                {
                    int prim = order[kid.first + k];
                    const float* A = positions + 3 * (size_t)indices[3 * (size_t)prim + 0];
                    const float* B = positions + 3 * (size_t)indices[3 * (size_t)prim + 1];
                    const float* C = positions + 3 * (size_t)indices[3 * (size_t)prim + 2];
                    B2Tri t;
                    t.ax = A[0]; t.ay = A[1]; t.az = A[2];
                    t.prim = prim;
                    t.e1x = B[0] - A[0]; t.e1y = B[1] - A[1]; t.e1z = B[2] - A[2]; t.pad1 = 0.0f;
                    t.e2x = C[0] - A[0]; t.e2y = C[1] - A[1]; t.e2z = C[2] - A[2]; t.pad2 = 0.0f;
                    out.tris.push_back(t);
                }
                tri_offset += (uint32_t)kid.count;
            }
        }
        node.exyz_imask = (uint32_t)(e[0] + 127) | ((uint32_t)(e[1] + 127) << 8) | ((uint32_t)(e[2] + 127) << 16) | (imask << 24);
        out.nodes[job.wide_index] = node;
    }

    out.stats.node_count = out.nodes.size();
    out.stats.triangle_count = out.tris.size();
    out.stats.bytes = out.nodes.size() * sizeof(B2Node) + out.tris.size() * sizeof(B2Tri);
    out.stats.sah_cost = sah;
    if (out.stats.max_depth >= B2_STACK_ENTRIES)
    {
        error = "CW8 tree deeper than the traversal stack";
        return false;
    }
    return true;
}
