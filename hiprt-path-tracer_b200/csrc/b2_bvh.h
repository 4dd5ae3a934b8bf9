/*
 * b2_bvh.h -- the compressed 8-wide BVH ("CW8") the B200 path traverses, and its traversal.
 *
 * This replaces HIPRT's geometry object + hiprtGeomTraversalClosest/AnyHit
 * (/root/reference/src/Device/includes/Intersect.h:38-69,212-215,284-293) and, for parity, must return what the
 * reference's CPU octree returns (src/Renderer/BVH.h:132-223): the hit with the smallest t among triangles with
 * t > 1e-7, t >= ray.minT, not rejected by the filter (self-hit, src/Device/functions/FilterFunction.h:22-35),
 * computed with the reference's exact Moeller-Trumbore operation order (src/Renderer/Triangle.h:20-63).
 *
 * Layout in HBM (built on the host by b2_bvh_build.cpp, SAH binned BVH2 collapsed to 8-wide):
 *   nodes      : 80-byte records = 5 x 16 B, fetched as five 128-bit loads
 *     word 0-2 : float3 origin of the node's quantisation grid
 *     word 3   : ex | ey<<8 | ez<<16 (biased exponents of the per-axis grid step) | imask<<24 (slots holding inner nodes)
 *     word 4   : index of the first inner child (inner children are contiguous, in slot order)
 *     word 5   : index of the first triangle referenced by this node's leaf slots
 *     word 6-7 : meta[8]   per slot: 0 = empty; inner = 0b001_11sss (sss = slot); leaf = unary count (1,3,7)<<5 | offset
 *     word 8-19: qlo_x[8] qlo_y[8] qlo_z[8] qhi_x[8] qhi_y[8] qhi_z[8]   child boxes, 8-bit, conservative
 *   triangles  : 48-byte records = 3 x 16 B in leaf order: (A.xyz, prim id), (B-A, 0), (C-A, 0)
 * Children are assigned to slots so that slot s is the nearest child for rays of octant s; the traversal order of
 * a node's hit children is then a pure bit trick (slot ^ (7 - octant)), no per-node sorting.
 * (Scheme after Ylitie, Karras, Laine: "Efficient Incoherent Ray Traversal on GPUs Through Compressed Wide BVHs",
 * HPG 2017 -- restated from the paper; no code of the reference or of HIPRT is involved.)
 *
 * Algorithmic bytes per traversal step (DESIGN.md section "Roofline"): S_node = 80 B, S_tri = 48 B.
 */
#ifndef B2_BVH_H
#define B2_BVH_H

#include "b2_math.h"

struct B2Node
{
    float ox, oy, oz;
    uint32_t exyz_imask;
    uint32_t child_base;
    uint32_t tri_base;
    uint8_t meta[8];
    uint8_t qlo_x[8], qlo_y[8], qlo_z[8];
    uint8_t qhi_x[8], qhi_y[8], qhi_z[8];
};
static_assert(sizeof(B2Node) == 80, "CW8 node must be 80 bytes");

struct B2Tri
{
    float ax, ay, az;
    int32_t prim;
    float e1x, e1y, e1z, pad1;
    float e2x, e2y, e2z, pad2;
};
static_assert(sizeof(B2Tri) == 48, "CW8 triangle must be 48 bytes");

/* Device-side descriptor; `b2pt_bvh_device_handle()` returns a device pointer to one of these and that pointer
 * is what goes into HIPRTRenderData::GPU_BVH. */
struct B2BvhDevice
{
    const B2Node* nodes;
    const B2Tri* tris;
    uint32_t node_count;
    uint32_t tri_count;
};

struct B2Hit
{
    float t;      /* -1 when no hit (reference: hiprtHit::t default) */
    float u, v;   /* barycentric weights of vertices B and C */
    int prim;     /* -1 when no hit */
    /* un-normalised geometric normal = cross(B-A, C-A); the caller normalises (Triangle.h:53) */
    V3 ng;
};

/* traversal statistics (instrumented variant only): "traversal steps per ray" of the north star */
struct B2TraversalCounters
{
    unsigned long long rays, nodes, tris;
};

/* Defensive bounds: a corrupted BVH (or a bug) must never hang the GPU. A traversal that exceeds B2_TRAVERSAL_GUARD
 * steps is abandoned and counted in the fault words (b2pt_counters_read reports them; tests assert they stay 0). */
#define B2_TRAVERSAL_GUARD (1u << 20)
#define B2_BOUNDARY_SKIP_GUARD 64
enum { B2_FAULT_TRAVERSAL_GUARD = 0, B2_FAULT_BOUNDARY_GUARD = 1, B2_FAULT_STACK_OVERFLOW = 2, B2_FAULT_COUNT = 3 };

B2_HD void b2_report_fault(unsigned int* faults, int which)
{
    if (!faults)
        return;
#if defined(__CUDA_ARCH__)
    atomicAdd(faults + which, 1u);
#else
    __atomic_fetch_add(faults + which, 1u, __ATOMIC_RELAXED);
#endif
}

#ifndef B2_STACK_ENTRIES
#define B2_STACK_ENTRIES 24
#endif

/* Exact restatement of Triangle::intersect (Triangle.h:20-63) on precomputed edges: edge1 = B-A and edge2 = C-A are
 * the same fp32 subtractions the reference performs per test, done once at build time. */
B2_HD bool b2_intersect_triangle(V3 org, V3 dir, V3 a, V3 edge1, V3 edge2, float& out_t, float& out_u, float& out_v)
{
    const float EPSILON = 0.0000001f;
    V3 h = cross(dir, edge2);
    float det = dot(edge1, h);
    if (det > -EPSILON && det < EPSILON)
        return false;
    float f = 1.0f / det;
    V3 s = org - a;
    float u = f * dot(s, h);
    if (u < 0.0f || u > 1.0f)
        return false;
    V3 q = cross(s, edge1);
    float v = f * dot(dir, q);
    if (v < 0.0f || u + v > 1.0f)
        return false;
    float t = f * dot(edge2, q);
    if (t > EPSILON)
    {
        out_t = t;
        out_u = u;
        out_v = v;
        return true;
    }
    return false;
}

/* The slab test only culls: it is conservative by construction (boxes widened by 8-bit quantisation and by the 0.5e-6 margins
 * below), so its arithmetic need not be the reference's -- one FMA per plane and an approximate reciprocal direction in both
 * codegen flavours. The triangle test, which decides t / u / v / prim, stays exact. */
B2_HD float b2_slab_fma(float a, float b, float c)
{
#if defined(__CUDA_ARCH__)
    return __fmaf_rn(a, b, c);
#else
    return a * b + c;
#endif
}
B2_HD float b2_slab_rcp(float x)
{
#if defined(__CUDA_ARCH__)
    float y;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
#else
    return 1.0f / x;
#endif
}

B2_HD uint32_t b2_bfind(uint32_t x)
{
#if defined(__CUDA_ARCH__)
    return 31u - (uint32_t)__clz((int)x);
#else
    return 31u - (uint32_t)__builtin_clz(x);
#endif
}
B2_HD uint32_t b2_popc(uint32_t x)
{
#if defined(__CUDA_ARCH__)
    return (uint32_t)__popc(x);
#else
    return (uint32_t)__builtin_popcount(x);
#endif
}

#if defined(__CUDA_ARCH__)
struct B2NodeRaw { uint4 q0, q1, q2, q3, q4; };
__device__ __forceinline__ B2NodeRaw b2_load_node(const B2Node* nodes, uint32_t index)
{
    const uint4* p = reinterpret_cast<const uint4*>(nodes + index);
    B2NodeRaw r;
    r.q0 = __ldg(p + 0);
    r.q1 = __ldg(p + 1);
    r.q2 = __ldg(p + 2);
    r.q3 = __ldg(p + 3);
    r.q4 = __ldg(p + 4);
    return r;
}
#endif

/*
 * One traversal. ANY_HIT = false: closest hit (ties: first found wins, as the reference's strict `<`).
 * ANY_HIT = true: returns at the first triangle hit with t < t_max (shadow rays; for opaque scenes this is what the
 * reference's closest-hit-then-compare loop decides, Intersect.h:295-322).
 * `skip_prim` is the reference's last_hit_primitive_index self-intersection filter.
 * `stack` points to this thread's first of B2_STACK_ENTRIES 8-byte slots, `stack_stride` slots apart (on the GPU the
 * slots live in shared memory, entry-major, so the 32 lanes of a warp touch 32 consecutive 8-byte words).
 */
/* The traversal as explicit state + one step, so that a kernel can interleave the steps of many rays in one converged
 * loop (b2_kernels.cu: a lane that finishes its ray fetches the next one from the block's pool while the other lanes
 * keep stepping). b2_traverse_rt below = begin; while (step); -- the same operations in the same order as before. */
/* ---- alpha testing inside the traversal (the reference's any-hit filter, src/Device/functions/FilterFunction.h:47-74) --------------
 * A candidate hit on a material that is not fully opaque is accepted with probability alpha_opacity x base-colour-texture alpha.
 * The reference draws that random number from the PATH's generator, once per candidate in the order its acceleration structure
 * meets them -- an order its own two back-ends (CPU octree, HIPRT) do not share, so a path's later draws already differ between
 * them. Here every RAY owns a small generator seeded from the bits of its origin and direction: which candidates are accepted
 * is a per-ray matter and the path's own stream is not touched. Scenes that use it are therefore held to the converged-image bar
 * (tests/test_gpu_converged.py), not to the per-pixel one; scenes without a non-opaque material never enter this code. */
struct b2pt_texture
{
    const uint8_t* texels;
    int width, height, channels;
    /* decode of the 256 possible texel values, computed ON THE HOST with the C library's powf (sample_texture_rgba's
     * `intrin_pow(rgba, 2.2f)`, Texture.h:62-65, is powf in the reference's host build) so that the device needs no pow at all */
    const float* srgb_to_linear;
};
struct B2AlphaTest
{
    const int* material_indices;            /* per triangle */
    const unsigned char* material_opaque;   /* per material: 1 = never filtered */
    const int* triangle_indices;
    const float* texcoords;                 /* 2 floats per vertex */
    const uint32_t* alpha_words;            /* materials_buffer.alpha_thin_film_hue_dielectric_priority: byte 0 = alpha_opacity */
    const uint32_t* base_color_words;       /* materials_buffer.base_color_roughness_metallic_index: low 16 bits = base colour texture */
    const b2pt_texture* textures;
    int enabled;                            /* the kernel was asked for alpha testing and the render settings allow it */
    int bounce_limit;                       /* rays of a vertex at bounce >= this are not filtered (alpha_testing_indirect_bounce) */
};
B2_HD uint32_t b2_alpha_seed(V3 org, V3 dir)
{
    uint32_t h = b2_float_as_uint(org.x) * 0x9E3779B1u ^ b2_float_as_uint(org.y) * 0x85EBCA77u ^ b2_float_as_uint(org.z) * 0xC2B2AE3Du;
    h ^= b2_float_as_uint(dir.x) * 0x27D4EB2Fu ^ b2_float_as_uint(dir.y) * 0x165667B1u ^ b2_float_as_uint(dir.z) * 0xD3A2646Cu;
    h = (h ^ 61u) ^ (h >> 16); /* wang_hash */
    h *= 9u;
    h ^= h >> 4;
    h *= 0x27D4EB2Du;
    h ^= h >> 15;
    return h ? h : 0x9E3779B9u; /* xorshift must not start at 0 */
}
/* Which of the reference's trace functions a ray stands for. Its CPU path -- the parity oracle -- does not treat them alike:
 * trace_main_path_ray goes through the filter alone (Intersect.h:200-257), while evaluate_shadow_ray_occluded and
 * evaluate_bsdf_light_sample_ray (Intersect.h:295-322, 587-640) re-trace from every accepted hit whose base-colour TEXTURE alpha is
 * below 1, i.e. for next-event rays a texel with any transparency never blocks. The same CPU path leaves FilterFunctionPayload::bounce
 * at its default 0 (intersect_scene_cpu, Intersect.h:180-194), so alpha_testing_indirect_bounce only switches the test on or off. */
enum { B2_ALPHA_RAY_NONE = 0, B2_ALPHA_RAY_MAIN = 1, B2_ALPHA_RAY_NEE = 2 };

/* true = the hit stands (FilterFunction.h returns false for it) */
B2_HD bool b2_alpha_test_accept(const B2AlphaTest& a, int prim, float u, float v, uint32_t& rng_state, bool nee_ray)
{
    const int material_index = a.material_indices[prim];
    if (a.material_opaque[material_index])
        return true;
    float base_color_alpha = 1.0f;
    const uint32_t texture_index = a.base_color_words[material_index] & 0xFFFFu;
    if (texture_index != 0xFFFFu && a.textures != nullptr)
    {
        /* get_hit_base_color_alpha (Material.h:24-37): uv_interpolate, then the alpha channel of sample_texture_rgba(is_srgb = true)
         * (Image8Bit::sample_rgba32f, Image.cpp:161-195: nearest texel, repeat, v flipped; the sRGB decode is applied to alpha too) */
        const int i0 = a.triangle_indices[prim * 3 + 0], i1 = a.triangle_indices[prim * 3 + 1], i2 = a.triangle_indices[prim * 3 + 2];
        const float w = 1.0f - u - v;
        const float tu = a.texcoords[2 * i1] * u + a.texcoords[2 * i2] * v + a.texcoords[2 * i0] * w;
        const float tv = a.texcoords[2 * i1 + 1] * u + a.texcoords[2 * i2 + 1] * v + a.texcoords[2 * i0 + 1] * w;
        const b2pt_texture& t = a.textures[texture_index];
        if (t.channels == 4)
        {
            float fu = tu, fv = tv;
            if (fu != 1.0f)
                fu -= (float)(int)tu;
            if (fv != 1.0f)
                fv -= (float)(int)tv;
            fu = fu < 0.0f ? 1.0f + fu : fu;
            fv = fv < 0.0f ? 1.0f + fv : fv;
            fv = 1.0f - fv;
            const int x = (int)(fu * (float)(t.width - 1)), y = (int)(fv * (float)(t.height - 1));
            base_color_alpha = t.srgb_to_linear[t.texels[((size_t)x + (size_t)y * (size_t)t.width) * 4 + 3]];
        }
    }
    if (nee_ray && base_color_alpha < 1.0f)
        return false;
    const float alpha_opacity = (float)(a.alpha_words[material_index] & 0xFFu) * 0.00392156862745098039f;
    const float composited_alpha = alpha_opacity * base_color_alpha;
    uint32_t x = rng_state;
    x ^= x << 13;
    x ^= x >> 17;
    x ^= x << 5;
    rng_state = x;
    float draw = (float)x / 4294967296.0f;
    draw = draw < 1.0f - 1.0e-7f ? draw : 1.0f - 1.0e-7f;
    return draw < composited_alpha;
}

struct B2Traversal
{
    V3 org, dir;
    float idx, idy, idz;
    uint32_t oct_inv;
    float t_max, t_far;
    int skip_prim;
    bool any_hit;
    B2Hit best;
    uint64_t* top; /* moving stack pointer: see the ptxas note below */
    int sp;
    uint32_t group_base, group_bits;
    unsigned int guard;
    unsigned int n_nodes, n_tris;
    uint32_t alpha_rng; /* 0 = this ray's candidate hits are not alpha-tested */
    bool alpha_nee;     /* B2_ALPHA_RAY_NEE */
};

B2_HD void b2_traverse_begin(B2Traversal& t, bool any_hit, V3 org, V3 dir, float t_max, int skip_prim, uint64_t* stack, int alpha_ray = B2_ALPHA_RAY_NONE)
{
    t.alpha_rng = alpha_ray != B2_ALPHA_RAY_NONE ? b2_alpha_seed(org, dir) : 0u;
    t.alpha_nee = alpha_ray == B2_ALPHA_RAY_NEE;
    t.org = org;
    t.dir = dir;
    t.any_hit = any_hit;
    t.t_max = t_max;
    t.t_far = t_max;
    t.skip_prim = skip_prim;
    t.best.t = -1.0f;
    t.best.u = t.best.v = 0.0f;
    t.best.prim = -1;
    t.best.ng = v3(0.0f);
    /* direction reciprocal, guarded against 0 (the slab test may be conservative; the triangle test is exact) */
    const float tiny = 1.0e-30f;
    t.idx = b2_slab_rcp(fabsf(dir.x) > tiny ? dir.x : (dir.x < 0.0f ? -tiny : tiny));
    t.idy = b2_slab_rcp(fabsf(dir.y) > tiny ? dir.y : (dir.y < 0.0f ? -tiny : tiny));
    t.idz = b2_slab_rcp(fabsf(dir.z) > tiny ? dir.z : (dir.z < 0.0f ? -tiny : tiny));
    const uint32_t octant = (dir.x < 0.0f ? 4u : 0u) | (dir.y < 0.0f ? 2u : 0u) | (dir.z < 0.0f ? 1u : 0u);
    t.oct_inv = 7u - octant;
    /* The stack is addressed through a moving pointer (`top`) plus a depth counter rather than stack[sp * stride]:
     * ptxas 12.9 (sm_100a) mis-folds the `sp - 1` of the pop into the LDS immediate AFTER already decrementing the
     * register (observed: "VIADD R22, R22, -1 ; IMAD R5, R22, 0x200, R37 ; LDS.64 [R5 + -0x200]"), which pops the entry
     * below the top; tools/traverse_selftest.cu is the regression test for this. */
    t.sp = 0;
    t.top = stack;
    t.group_base = 0;            /* node group: index of first child node */
    t.group_bits = 0x80000000u;  /* bits 24..31: pending inner children (by traversal priority); bits 0..7: imask */
    t.guard = 0;
    t.n_nodes = t.n_tris = 0;
}

/* One step = node part, triangle part, pop part, in this order (b2_traverse_step). The three parts are separate functions so that a
 * kernel can run them as warp-wide phases (wf_trace: lanes that found leaf triangles wait until enough lanes have some).
 *
 * Node part: at most one node (its 8 quantised child boxes); the triangles of its hit leaf slots come back in tri_base / tri_bits.
 * Returns false when the step guard tripped (the traversal is abandoned). */
B2_HD bool b2_traverse_node(B2Traversal& t, const B2BvhDevice& bvh, int stack_stride, unsigned int* faults, int stack_entries, uint32_t& tri_base, uint32_t& tri_bits)
{
    const V3 org = t.org, dir = t.dir;
    const uint32_t oct_inv = t.oct_inv;
    tri_base = 0;
    tri_bits = 0;
    if (++t.guard > B2_TRAVERSAL_GUARD)
    {
        b2_report_fault(faults, B2_FAULT_TRAVERSAL_GUARD);
        return false;
    }

    if (t.group_bits > 0x00FFFFFFu)
    {
        const uint32_t child_bit = b2_bfind(t.group_bits);
        const uint32_t slot = (child_bit - 24u) ^ oct_inv;
        const uint32_t imask = t.group_bits & 0xFFu;
        const uint32_t node_index = t.group_base + b2_popc(imask & ~(0xFFFFFFFFu << slot));
        t.group_bits &= ~(1u << child_bit);
        if (t.group_bits > 0x00FFFFFFu)
        {
            if (t.sp < stack_entries)
            {
                *t.top = ((uint64_t)t.group_base << 32) | t.group_bits;
                t.top += stack_stride;
            }
            else
                b2_report_fault(faults, B2_FAULT_STACK_OVERFLOW);
            t.sp++; /* overflow is counted but dropped; the builder bounds the depth well below the stack size */
        }

        t.n_nodes++;
#ifdef B2_TRAVERSE_TRACE
        B2_TRAVERSE_TRACE(t.guard, node_index, t.group_base, t.group_bits, t.sp);
#endif
#if defined(__CUDA_ARCH__)
        const B2NodeRaw raw = b2_load_node(bvh.nodes, node_index);
        const float nox = __uint_as_float(raw.q0.x), noy = __uint_as_float(raw.q0.y), noz = __uint_as_float(raw.q0.z);
        const uint32_t exyz = raw.q0.w;
        const uint32_t n_child_base = raw.q1.x, n_tri_base = raw.q1.y;
        const uint32_t meta_lo = raw.q1.z, meta_hi = raw.q1.w;
        const uint32_t qw[12] = { raw.q2.x, raw.q2.y, raw.q2.z, raw.q2.w, raw.q3.x, raw.q3.y, raw.q3.z, raw.q3.w, raw.q4.x, raw.q4.y, raw.q4.z, raw.q4.w };
#else
        const B2Node& nd = bvh.nodes[node_index];
        const float nox = nd.ox, noy = nd.oy, noz = nd.oz;
        const uint32_t exyz = nd.exyz_imask;
        const uint32_t n_child_base = nd.child_base, n_tri_base = nd.tri_base;
        uint32_t meta_lo, meta_hi, qw[12];
        {
            const uint32_t* w = reinterpret_cast<const uint32_t*>(nd.meta);
            meta_lo = w[0];
            meta_hi = w[1];
            for (int k = 0; k < 12; k++)
                qw[k] = w[2 + k];
        }
#endif
        /* grid step 2^(e-127) per axis, folded into the reciprocal direction */
        const float sx = b2_uint_as_float((exyz & 0xFFu) << 23) * t.idx;
        const float sy = b2_uint_as_float(((exyz >> 8) & 0xFFu) << 23) * t.idy;
        const float sz = b2_uint_as_float(((exyz >> 16) & 0xFFu) << 23) * t.idz;
        const float bx = (nox - org.x) * t.idx;
        const float by = (noy - org.y) * t.idy;
        const float bz = (noz - org.z) * t.idz;
        const float t_far = t.t_far;

        /* near / far planes per axis are picked once per node on the packed words (12 selects) instead of per child (48) */
        const bool neg_x = dir.x < 0.0f, neg_y = dir.y < 0.0f, neg_z = dir.z < 0.0f;
        const uint32_t near_x[2] = { neg_x ? qw[6] : qw[0], neg_x ? qw[7] : qw[1] }, far_x[2] = { neg_x ? qw[0] : qw[6], neg_x ? qw[1] : qw[7] };
        const uint32_t near_y[2] = { neg_y ? qw[8] : qw[2], neg_y ? qw[9] : qw[3] }, far_y[2] = { neg_y ? qw[2] : qw[8], neg_y ? qw[3] : qw[9] };
        const uint32_t near_z[2] = { neg_z ? qw[10] : qw[4], neg_z ? qw[11] : qw[5] }, far_z[2] = { neg_z ? qw[4] : qw[10], neg_z ? qw[5] : qw[11] };

        uint32_t hit_bits = 0;
#pragma unroll
        for (int c = 0; c < 8; c++)
        {
            /* branch-free over the 8 slots: an empty slot has meta = 0 and contributes (0 >> 5) << x = no bit */
            const uint32_t meta = ((c < 4 ? meta_lo : meta_hi) >> ((c & 3) * 8)) & 0xFFu;
            const int w = c >> 2, sh = (c & 3) * 8;
            const float tx0 = b2_slab_fma((float)((near_x[w] >> sh) & 0xFFu), sx, bx), tx1 = b2_slab_fma((float)((far_x[w] >> sh) & 0xFFu), sx, bx);
            const float ty0 = b2_slab_fma((float)((near_y[w] >> sh) & 0xFFu), sy, by), ty1 = b2_slab_fma((float)((far_y[w] >> sh) & 0xFFu), sy, by);
            const float tz0 = b2_slab_fma((float)((near_z[w] >> sh) & 0xFFu), sz, bz), tz1 = b2_slab_fma((float)((far_z[w] >> sh) & 0xFFu), sz, bz);
            /* widened by a few ulps so fp32 rounding in the slab arithmetic can never cull a box the exact
             * triangle test would hit */
            const float t_enter = fmaxf(fmaxf(tx0, ty0), fmaxf(tz0, 0.0f)) * 0.9999995f;
            const float t_exit = fminf(fminf(tx1, ty1), fminf(tz1, t_far)) * 1.0000005f;
            const bool inner = (meta & 0x18u) == 0x18u && (meta >> 5) == 1u;
            const uint32_t bit_index = inner ? (24u + ((meta & 7u) ^ oct_inv)) : (meta & 0x1Fu);
            hit_bits |= (t_enter <= t_exit) ? ((meta >> 5) << bit_index) : 0u;
        }
        t.group_base = n_child_base;
        t.group_bits = (hit_bits & 0xFF000000u) | (exyz >> 24);
        tri_base = n_tri_base;
        tri_bits = hit_bits & 0x00FFFFFFu;
#if defined(__CUDA_ARCH__) && defined(B2_TRAVERSE_PREFETCH)
        /* what the next steps will fetch is known now: the first hit child node and the first leaf triangle. Their lines are asked
         * for here, one triangle phase / loop turn (100 - 300 cycles) before the dependent loads that wait for them. */
        if (t.group_bits > 0x00FFFFFFu)
        {
            const uint32_t next_bit = b2_bfind(t.group_bits);
            const uint32_t next_slot = (next_bit - 24u) ^ oct_inv;
            const uint32_t next_index = n_child_base + b2_popc((t.group_bits & 0xFFu) & ~(0xFFFFFFFFu << next_slot));
            asm volatile("prefetch.global.L1 [%0];" ::"l"(bvh.nodes + next_index));
        }
        if (tri_bits != 0u)
            asm volatile("prefetch.global.L1 [%0];" ::"l"(bvh.tris + n_tri_base + b2_bfind(tri_bits)));
#endif
    }
    else
    {
        t.group_bits = 0;
    }
    return true;
}

/* Triangle part: the exact tests of the leaf triangles found by the node part. Returns true when an any-hit ray is done. */
B2_HD bool b2_traverse_triangles(B2Traversal& t, const B2BvhDevice& bvh, uint32_t tri_base, uint32_t tri_bits, const B2AlphaTest* alpha = nullptr)
{
    const V3 org = t.org, dir = t.dir;
    while (tri_bits != 0)
    {
        const uint32_t k = b2_bfind(tri_bits);
        tri_bits &= ~(1u << k);
        t.n_tris++;
#if defined(__CUDA_ARCH__)
        const float4* tp = reinterpret_cast<const float4*>(bvh.tris + tri_base + k);
        const float4 ta = __ldg(tp), t1 = __ldg(tp + 1), t2 = __ldg(tp + 2);
        const int prim = __float_as_int(ta.w);
        const V3 a = v3(ta.x, ta.y, ta.z), e1 = v3(t1.x, t1.y, t1.z), e2 = v3(t2.x, t2.y, t2.z);
#else
        const B2Tri& tr = bvh.tris[tri_base + k];
        const int prim = tr.prim;
        const V3 a = v3(tr.ax, tr.ay, tr.az), e1 = v3(tr.e1x, tr.e1y, tr.e1z), e2 = v3(tr.e2x, tr.e2y, tr.e2z);
#endif
        if (prim == t.skip_prim)
            continue;
        float th, u, v;
        if (!b2_intersect_triangle(org, dir, a, e1, e2, th, u, v))
            continue;
        if (alpha != nullptr && t.alpha_rng != 0u && !b2_alpha_test_accept(*alpha, prim, u, v, t.alpha_rng, t.alpha_nee))
            continue; /* the ray passes through (FilterFunction.h:66-73) */
        if (t.any_hit)
        {
            if (th < t.t_max)
            {
                t.best.t = th; t.best.u = u; t.best.v = v; t.best.prim = prim; t.best.ng = cross(e1, e2);
                return true;
            }
        }
        else if (th < t.t_far)
        {
            /* strictly closer. Equal-t candidates keep the first one found, like the reference's `<`; which
             * one is "first" depends on the acceleration structure (the documented "ties excepted"). */
            t.best.t = th; t.best.u = u; t.best.v = v; t.best.prim = prim; t.best.ng = cross(e1, e2);
            t.t_far = th;
        }
    }
    return false;
}

/* Pop part: when the current group is exhausted, the next one comes off the stack. Returns false when the traversal is over. */
B2_HD bool b2_traverse_pop(B2Traversal& t, int stack_stride, int stack_entries)
{
    if (t.group_bits <= 0x00FFFFFFu)
    {
        if (t.sp == 0)
            return false;
        t.sp--;
        if (t.sp < stack_entries)
        {
            t.top -= stack_stride;
            const uint64_t e = *t.top;
            t.group_base = (uint32_t)(e >> 32);
            t.group_bits = (uint32_t)e;
        }
        else
            t.group_bits = 0;
    }
    return true;
}

/* Returns false when the traversal is over (t.best holds the result). */
B2_HD bool b2_traverse_step(B2Traversal& t, const B2BvhDevice& bvh, int stack_stride, unsigned int* faults, int stack_entries = B2_STACK_ENTRIES, const B2AlphaTest* alpha = nullptr)
{
    uint32_t tri_base, tri_bits;
    if (!b2_traverse_node(t, bvh, stack_stride, faults, stack_entries, tri_base, tri_bits))
        return false;
    if (b2_traverse_triangles(t, bvh, tri_base, tri_bits, alpha))
        return false;
    return b2_traverse_pop(t, stack_stride, stack_entries);
}

B2_HD B2Hit b2_traverse_rt(const bool ANY_HIT, const B2BvhDevice& bvh, V3 org, V3 dir, float t_max, int skip_prim, uint64_t* stack, int stack_stride,
                           B2TraversalCounters* counters, unsigned int* faults, const B2AlphaTest* alpha = nullptr, int alpha_ray = B2_ALPHA_RAY_NONE)
{
    B2Traversal t;
    b2_traverse_begin(t, ANY_HIT, org, dir, t_max, skip_prim, stack, alpha != nullptr ? alpha_ray : B2_ALPHA_RAY_NONE);
    if (bvh.node_count != 0)
    {
        while (b2_traverse_step(t, bvh, stack_stride, faults, B2_STACK_ENTRIES, alpha))
        {
        }
        if (counters) { counters->rays += 1; counters->nodes += t.n_nodes; counters->tris += t.n_tris; }
    }
    return t.best;
}

template <bool ANY_HIT>
B2_HD B2Hit b2_traverse(const B2BvhDevice& bvh, V3 org, V3 dir, float t_max, int skip_prim, uint64_t* stack, int stack_stride, B2TraversalCounters* counters,
                        unsigned int* faults = nullptr)
{
    return b2_traverse_rt(ANY_HIT, bvh, org, dir, t_max, skip_prim, stack, stack_stride, counters, faults);
}

#if defined(__CUDACC__)
/* ONE out-of-line copy of the traversal per kernel (closest-hit and any-hit share it through a run-time flag): the
 * per-pixel code calls the traversal from 5-7 places (main path, shadow rays of every NEE strategy, BSDF-sample rays)
 * and inlining every one of them made the C2 MegaKernel 13.7k SASS instructions = 220 KB, far beyond the 32 KB L1.5
 * instruction cache (ncu: `no_instruction` was the dominant stall, profiles/r01_ncu_c2_megakernel_first.md).
 * Members are passed by value so the caller's context stays in registers. */
static __device__ __noinline__ B2Hit b2_traverse_shared(bool any_hit, const B2Node* nodes, const B2Tri* tris, uint32_t node_count, float ox, float oy, float oz, float dx,
                                                        float dy, float dz, float t_max, int skip_prim, uint64_t* stack, int stack_stride,
                                                        B2TraversalCounters* counters, unsigned int* faults, const B2AlphaTest* alpha, int alpha_ray)
{
    B2BvhDevice bvh;
    bvh.nodes = nodes;
    bvh.tris = tris;
    bvh.node_count = node_count;
    bvh.tri_count = 0;
    return b2_traverse_rt(any_hit, bvh, v3(ox, oy, oz), v3(dx, dy, dz), t_max, skip_prim, stack, stack_stride, counters, faults, alpha, alpha_ray);
}
#endif

#endif
