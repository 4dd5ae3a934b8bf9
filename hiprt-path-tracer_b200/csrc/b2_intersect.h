/*
 * b2_intersect.h -- the three ray queries of the hot path and the hit -> surface reconstruction.
 *
 * Restates the CLIENT side of the acceleration structure, i.e. /root/reference/src/Device/includes/Intersect.h:
 *   trace_main_path_ray             :200-270   closest hit + HitInfo + material + interior-stack push + normal fix-up
 *   evaluate_shadow_ray_occluded    :276-324   any hit before t_max - 1e-4
 *   evaluate_bsdf_light_sample_ray  :544-643   closest hit returning the emission of what was hit
 * plus NestedDielectricsInteriorStack (NestedDielectrics.h:64-255), RayVolumeState (RayVolumeState.h:13-80),
 * RayPayload (RayPayload.h:22-92) and fix_backfacing_normals (Intersect.h:142-176).
 * Textured materials (8-bit textures, sRGB decode, normal maps, emissive textures: Texture.h, Material.h, Intersect.h:77-134) included.
 * Alpha-tested materials: the any-hit filter's test (FilterFunction.h:47-74) runs inside the traversal (b2_bvh.h) with a per-ray
 * generator -- statistically equivalent to the reference, not bit-exact: the reference draws from the path's generator in the order its
 * own acceleration structure meets candidate triangles (BVH.h:153-176 calls the filter for every triangle of every visited leaf).
 * Spectral dispersion (Intersect.h:259-267, Dispersion.h) included.
 */
#ifndef B2_INTERSECT_H
#define B2_INTERSECT_H

#include "b2_bvh.h"
#include "b2_material.h"

/* strategy constants: values of the reference's option macros (KernelOptions.h:40-57, DirectLightSamplingOptions.h:11-20) */
enum { B2_BSDF_NONE = 0, B2_BSDF_LAMBERTIAN = 1, B2_BSDF_OREN_NAYAR = 2, B2_BSDF_PRINCIPLED = 3 };
enum { B2_LSS_NONE = 0, B2_LSS_ONE_LIGHT = 1, B2_LSS_BSDF = 2, B2_LSS_MIS = 3, B2_LSS_RIS = 4, B2_LSS_RESTIR_DI = 5 };
enum { B2_BASE_UNIFORM = 0, B2_BASE_POWER = 1 };
enum { B2_ESS_NONE = 0, B2_ESS_BINARY_SEARCH = 1, B2_ESS_ALIAS_TABLE = 2 };
enum { B2_LATER_ONE_LIGHT = 0, B2_LATER_BSDF = 1, B2_LATER_MIS = 2, B2_LATER_RIS = 3 };

/* A compile-time variant = one value per strategy macro the reference would pass with -D at JIT time. */
template <int BSDF_OVERRIDE_, int DIRECT_, int BASE_, int LATER_ = B2_LATER_RIS, int ENVMAP_ = B2_ESS_ALIAS_TABLE, bool COUNTERS_ = false>
struct B2Variant
{
    static constexpr int BSDF_OVERRIDE = BSDF_OVERRIDE_;
    static constexpr int DIRECT = DIRECT_;
    static constexpr int BASE = BASE_;
    static constexpr int LATER = LATER_;
    static constexpr int ENVMAP = ENVMAP_;
    static constexpr bool COUNTERS = COUNTERS_;
};

/* BSDFIncidentLightInfo (BSDFs/BSDFIncidentLightInfo.h:9-27) */
enum
{
    B2_LOBE_NO_INFO = 0,
    B2_LOBE_COAT = 1 << 1,
    B2_LOBE_FIRST_METAL = 1 << 2,
    B2_LOBE_SECOND_METAL = 1 << 3,
    B2_LOBE_SPECULAR = 1 << 4,
    B2_LOBE_GLASS_REFLECT = 1 << 5,
    B2_LOBE_GLASS_REFRACT = 1 << 6,
    B2_LOBE_DIFFUSE = 1 << 7,
    B2_LOBE_DIFFUSE_TRANSMISSION = 1 << 8,
    B2_LOBE_NOT_FROM_BSDF = (1 << 8) + 1
};

#define B2_DIELECTRIC_STACK_SIZE 4
#define B2_MAX_MATERIAL_INDEX 0x03FFFFFFu /* 26 bits, StackPriorityEntry::MATERIAL_INDEX_MAXIMUM */

/* entry bits: priority[0..3] topmost[4] odd_parity[5] material_index[6..31] */
struct InteriorStack
{
    uint32_t entries[B2_DIELECTRIC_STACK_SIZE];
    int position;

    B2_HD static int e_priority(uint32_t e) { return (int)(e & 0xFu); }
    B2_HD static bool e_topmost(uint32_t e) { return (e >> 4) & 1u; }
    B2_HD static bool e_odd(uint32_t e) { return (e >> 5) & 1u; }
    B2_HD static int e_material(uint32_t e) { return (int)((e >> 6) & B2_MAX_MATERIAL_INDEX); }
    B2_HD static uint32_t e_make(int material, bool odd, bool topmost, int priority)
    {
        return ((uint32_t)priority & 0xFu) | ((topmost ? 1u : 0u) << 4) | ((odd ? 1u : 0u) << 5) | (((uint32_t)material & B2_MAX_MATERIAL_INDEX) << 6);
    }

    B2_HD void init()
    {
        for (int i = 0; i < B2_DIELECTRIC_STACK_SIZE; i++)
            entries[i] = e_make((int)B2_MAX_MATERIAL_INDEX, true, true, 0);
        position = 0;
    }

    /* returns true when the boundary must be skipped (lower priority than the enclosing medium) */
    B2_HD bool push(int& out_incident_material, int& out_outgoing_material, bool& out_inside_material, int material_index, int material_priority)
    {
        if (position == B2_DIELECTRIC_STACK_SIZE - 1)
            return false;

        int last_entered = 0;
        for (last_entered = position; last_entered >= 0; last_entered--)
        {
            uint32_t e = entries[last_entered];
            if (e_material(e) != material_index && e_topmost(e) && e_odd(e))
                break;
        }

        bool odd_parity = true;
        for (int same = position; same >= 0; same--)
        {
            if (e_material(entries[same]) == material_index)
            {
                entries[same] &= ~(1u << 4);
                odd_parity = !e_odd(entries[same]);
                break;
            }
        }

        out_inside_material = !odd_parity;
        if (position < B2_DIELECTRIC_STACK_SIZE - 1)
            position++;
        entries[position] = e_make(material_index, odd_parity, true, material_priority);

        /* entry 0 is the "vacuum" sentinel (max material index, topmost, odd), so the search above always stops at >= 0 */
        uint32_t enclosing = entries[last_entered < 0 ? 0 : last_entered];
        if (material_priority < e_priority(enclosing))
            return true;
        if (odd_parity)
        {
            out_incident_material = e_material(enclosing);
            out_outgoing_material = material_index;
        }
        else
        {
            out_incident_material = material_index;
            out_outgoing_material = e_material(enclosing);
        }
        return false;
    }

    B2_HD void pop(bool inside_material)
    {
        int top_material = e_material(entries[position]);
        if (position > 0)
            position--;

        if (inside_material)
        {
            int same;
            for (same = position; same >= 0; same--)
                if (e_material(entries[same]) == top_material)
                    break;
            if (same >= 0)
                for (int i = same + 1; i <= position; i++)
                    entries[i - 1] = entries[i];
            if (position > 0)
                position--;
        }

        for (int i = position; i >= 0; i--)
        {
            if (e_material(entries[i]) == top_material)
            {
                entries[i] |= (1u << 4);
                break;
            }
        }
    }
};

struct VolumeState
{
    float distance_in_volume;
    InteriorStack interior_stack;
    int incident_mat_index, outgoing_mat_index;
    bool inside_material;
    float sampled_wavelength;

    B2_HD void init()
    {
        distance_in_volume = 0.0f;
        interior_stack.init();
        incident_mat_index = -1;
        outgoing_mat_index = -1;
        inside_material = false;
        sampled_wavelength = 0.0f;
    }
};

struct HitInfo
{
    V3 inter_point;
    V3 shading_normal;
    V3 geometric_normal;
    V2 texcoords;
    float t;
    int primitive_index;
};
B2_HD HitInfo hit_info_defaults()
{
    HitInfo h;
    h.inter_point = v3(0.0f);
    h.shading_normal = v3(0.0f);
    h.geometric_normal = v3(0.0f);
    h.texcoords = v2(0.0f, 0.0f);
    h.t = -1.0f;
    h.primitive_index = -1;
    return h;
}

enum { B2_RAY_BOUNCE = 0, B2_RAY_MISSED = 1 };

struct Payload
{
    Rgb throughput;
    Rgb ray_color;
    int next_ray_state;
    int bounce;
    float accumulated_roughness;
    Mat material;
    VolumeState volume_state;

    B2_HD void init()
    {
        throughput = rgb(1.0f);
        ray_color = rgb(0.0f);
        next_ray_state = B2_RAY_BOUNCE;
        bounce = 0;
        accumulated_roughness = 0.0f;
        material = mat_defaults();
        volume_state.init();
    }

    /* RayPayload::accumulate_roughness, RayPayload.h:36-90 */
    B2_HD void accumulate_roughness(int sampled_lobe)
    {
        switch (sampled_lobe)
        {
        case B2_LOBE_DIFFUSE:
        case B2_LOBE_DIFFUSE_TRANSMISSION: accumulated_roughness = 1.0f; break;
        case B2_LOBE_COAT: accumulated_roughness = b2_max(material.coat_roughness, accumulated_roughness); break;
        case B2_LOBE_SECOND_METAL: accumulated_roughness = b2_max(material.second_roughness, accumulated_roughness); break;
        case B2_LOBE_FIRST_METAL:
        case B2_LOBE_SPECULAR:
        case B2_LOBE_GLASS_REFLECT:
        case B2_LOBE_GLASS_REFRACT: accumulated_roughness = b2_max(material.roughness, accumulated_roughness); break;
        default: break;
        }
    }
};

/* Everything a per-pixel function needs besides its own state. */
struct Ctx
{
    const b2pt_render_data* rd;
    B2BvhDevice bvh;
    uint64_t* stack;               /* first of B2_STACK_ENTRIES slots private to this thread */
    int stack_stride;              /* distance between this thread's consecutive slots */
    B2TraversalCounters* counters; /* null unless the instrumented variant runs */
    unsigned int* faults;          /* B2_FAULT_COUNT words (device global / host array) or null */
    int light_presampling;         /* ReSTIR_DI_DoLightPresampling: initial candidates come from the presampled light buffer (1: presampled with the
                                    * power alias table, 2: uniformly -- ReSTIR_DI_LightPresamplingStrategy) */
    int bias_correction;           /* ReSTIR_DI_BiasCorrectionWeights (B2_BIAS_*, b2_restir_di.h): a launch-time switch of the reuse passes */
    int restir_options;  /* B2_DI_OPT_* (b2_restir_di.h): the four ReSTIR DI visibility macros of ReSTIRDIOptions.h:50-84 */
    bool alpha_test;               /* the kernel was asked for alpha testing (B2PT_ALPHA_TEST: the scene has materials that are not fully opaque) */
#if defined(B2_ALPHA_TEST_GATED_BY_BOUNCE)
    mutable int ray_bounce;        /* bounce of the vertex whose rays are being traced in place (FilterFunctionPayload::bounce): set by the pass / bounce code */
#endif
};
#if defined(B2_ALPHA_TEST_GATED_BY_BOUNCE)
#define B2_SET_RAY_BOUNCE(ctx, bounce) ((ctx).ray_bounce = (bounce))
#define B2_RAY_BOUNCE(ctx) ((ctx).ray_bounce)
#else
#define B2_SET_RAY_BOUNCE(ctx, bounce) ((void)0)
#define B2_RAY_BOUNCE(ctx) 0
#endif

/* the B2AlphaTest of a render data (RenderBuffers arrays + the two render settings of FilterFunction.h:47-53) */
B2_HD B2AlphaTest b2_alpha_test_of(const b2pt_render_data& rd, bool kernel_wants_it)
{
    B2AlphaTest a;
    a.material_indices = rd.buffers.material_indices;
    a.material_opaque = rd.buffers.material_opaque;
    a.triangle_indices = rd.buffers.triangles_indices;
    a.texcoords = reinterpret_cast<const float*>(rd.buffers.texcoords);
    a.alpha_words = rd.buffers.materials_buffer.alpha_thin_film_hue_dielectric_priority;
    a.base_color_words = rd.buffers.materials_buffer.base_color_roughness_metallic_index;
    a.textures = reinterpret_cast<const b2pt_texture*>(rd.buffers.material_textures);
    a.bounce_limit = rd.render_settings.alpha_testing_indirect_bounce;
    a.enabled = kernel_wants_it && rd.render_settings.do_alpha_testing && a.bounce_limit > 0 && a.material_opaque != nullptr;
    return a;
}
/* Does a ray of a vertex at `bounce` get its candidate hits alpha-tested? The reference's CPU path (the oracle) never hands the bounce
 * to its filter -- FilterFunctionPayload::bounce stays 0 (b2_bvh.h) -- so the limit only switches the test on or off; `bounce` is
 * what its GPU path would compare (kept in the interface, B2_ALPHA_TEST_GATED_BY_BOUNCE selects that behaviour at build time). */
B2_HD bool b2_alpha_tested_at(const B2AlphaTest& a, int bounce)
{
#if defined(B2_ALPHA_TEST_GATED_BY_BOUNCE)
    return a.enabled && bounce < a.bounce_limit;
#else
    (void)bounce;
    return a.enabled != 0;
#endif
}

/* are the queued rays of a vertex at `bounce` alpha-tested? (the wavefront forms decide when they queue a ray) */
B2_HD bool b2_ctx_alpha_tested_at(const Ctx& ctx, int bounce)
{
    if (!ctx.alpha_test)
        return false;
    return b2_alpha_tested_at(b2_alpha_test_of(*ctx.rd, true), bounce);
}

/* HIPRTRenderSettings::do_render_low_resolution, RenderSettings.h */
/* every ray of the per-pixel code goes through here: on the device one shared out-of-line traversal (b2_bvh.h) */
B2_HD B2Hit b2_trace(const Ctx& ctx, bool any_hit, V3 origin, V3 direction, float t_max, int skip_prim, bool main_path_ray = false)
{
    const int alpha_ray = main_path_ray ? B2_ALPHA_RAY_MAIN : B2_ALPHA_RAY_NEE;
    if (ctx.alpha_test)
    {
        /* the rare path (scenes with alpha-tested materials): the filter's inputs are gathered from the render data here, so that the
         * context every kernel carries stays one flag */
        const B2AlphaTest alpha = b2_alpha_test_of(*ctx.rd, true);
        if (b2_alpha_tested_at(alpha, B2_RAY_BOUNCE(ctx)))
        {
#if defined(__CUDA_ARCH__) && !defined(B2_INLINE_TRAVERSAL)
            return b2_traverse_shared(any_hit, ctx.bvh.nodes, ctx.bvh.tris, ctx.bvh.node_count, origin.x, origin.y, origin.z, direction.x, direction.y, direction.z, t_max,
                                      skip_prim, ctx.stack, ctx.stack_stride, ctx.counters, ctx.faults, &alpha, alpha_ray);
#else
            return b2_traverse_rt(any_hit, ctx.bvh, origin, direction, t_max, skip_prim, ctx.stack, ctx.stack_stride, ctx.counters, ctx.faults, &alpha, alpha_ray);
#endif
        }
    }
#if defined(__CUDA_ARCH__) && !defined(B2_INLINE_TRAVERSAL)
    return b2_traverse_shared(any_hit, ctx.bvh.nodes, ctx.bvh.tris, ctx.bvh.node_count, origin.x, origin.y, origin.z, direction.x, direction.y, direction.z, t_max,
                              skip_prim, ctx.stack, ctx.stack_stride, ctx.counters, ctx.faults, nullptr, B2_ALPHA_RAY_NONE);
#else
    return b2_traverse_rt(any_hit, ctx.bvh, origin, direction, t_max, skip_prim, ctx.stack, ctx.stack_stride, ctx.counters, ctx.faults);
#endif
}

B2_HD bool do_render_low_resolution(const b2pt_render_settings& s)
{
    return s.wants_render_low_resolution && s.allow_render_low_resolution && s.accumulate;
}

B2_HD V3 load_v3(const b2pt_float3* p, int i) { return v3(p[i].x, p[i].y, p[i].z); }
B2_HD Rgb load_rgb(const b2pt_rgb* p, int i) { return rgb(p[i].r, p[i].g, p[i].b); }

B2_HD V3 reflect_ray(V3 ray_direction, V3 surface_normal)
{
    return 2.0f * dot(ray_direction, surface_normal) * surface_normal - ray_direction;
}

/* Intersect.h:142-176 */
B2_HD void fix_backfacing_normals(HitInfo& hit, V3 view_direction)
{
    if (dot(view_direction, hit.geometric_normal) < 0.0f)
    {
        hit.geometric_normal *= -1.0f;
        hit.shading_normal *= -1.0f;
    }
    if (dot(view_direction, hit.shading_normal) < 0.0f)
        hit.shading_normal *= -1.0f;

    V3 reflected = reflect_ray(view_direction, hit.shading_normal);
    if (dot(reflected, hit.geometric_normal) <= 0.0f)
    {
        const float epsilon = 0.01f;
        reflected -= normalize((dot(reflected, hit.geometric_normal) - epsilon) * hit.geometric_normal);
        hit.shading_normal = normalize(view_direction + reflected);
    }
}

/* local_to_world_frame (ONB.h): T * v.x + B * v.y + N * v.z */
B2_HD V3 b2_local_to_world_frame(V3 T, V3 B, V3 N, V3 v) { return normalize(v.x * T + v.y * B + v.z * N); }

/* normal_mapping (Intersect.h:77-112) */
B2_HD V3 normal_mapping(const b2pt_render_data& rd, int normal_map_texture_index, int i0, int i1, int i2, V2 interpolated_texcoords, V3 surface_normal)
{
    const b2pt_float2 tA = rd.buffers.texcoords[i0], tB = rd.buffers.texcoords[i1], tC = rd.buffers.texcoords[i2];
    V2 delta_P1P0 = v2(tB.x, tB.y) - v2(tA.x, tA.y);
    V2 delta_P2P0 = v2(tC.x, tC.y) - v2(tA.x, tA.y);
    V3 P0 = load_v3(rd.buffers.vertices_positions, i0), P1 = load_v3(rd.buffers.vertices_positions, i1), P2 = load_v3(rd.buffers.vertices_positions, i2);
    V3 edge_P0P1 = P1 - P0;
    V3 edge_P0P2 = P2 - P0;

    const float det_bias = 1.0e-6f;
    float det = delta_P1P0.x * delta_P2P0.y - delta_P1P0.y * delta_P2P0.x + det_bias;
    float det_inverse = 1.0f / det;
    V3 T = (edge_P0P1 * delta_P2P0.y - edge_P0P2 * delta_P1P0.y) * det_inverse;
    V3 B = (edge_P0P2 * delta_P1P0.x - edge_P0P1 * delta_P2P0.x) * det_inverse;
    if (length2(T) < 1.0e-6f || length2(B) < 1.0e-6f)
        return surface_normal;

    float rgba[4];
    sample_texture_rgba(rd.buffers.material_textures, normal_map_texture_index, false, interpolated_texcoords, rgba);
    V3 normal_tangent_space = normalize(v3(rgba[0] - 0.5f, rgba[1] - 0.5f, rgba[2] - 0.5f));
    return b2_local_to_world_frame(normalize(T), normalize(B), surface_normal, normal_tangent_space);
}

/* get_shading_normal (Intersect.h:114-134): smooth normal when the first vertex has one, then the material's normal map */
B2_HD V3 shading_normal_at(const b2pt_render_data& rd, V3 geometric_normal, int i0, int i1, int i2, float u, float v, int primitive_index, V2 interpolated_texcoords)
{
    if (!rd.render_settings.do_normal_mapping)
        return geometric_normal;
    V3 surface_normal = geometric_normal;
    if (rd.buffers.has_vertex_normals[i0])
    {
        /* uv_interpolate: A * (1 - u - v) + B * u + C * v  (Texture.h) */
        V3 nA = load_v3(rd.buffers.vertex_normals, i0), nB = load_v3(rd.buffers.vertex_normals, i1), nC = load_v3(rd.buffers.vertex_normals, i2);
        surface_normal = normalize(nB * u + nC * v + nA * (1.0f - u - v));
    }
    if (rd.buffers.material_textures != nullptr)
    {
        const int material_index = rd.buffers.material_indices[primitive_index];
        const int normal_map_texture_index = (int)(rd.buffers.materials_buffer.normal_map_emission_index[material_index] & 0xFFFFu);
        if (normal_map_texture_index != B2_NO_TEXTURE)
            surface_normal = normal_mapping(rd, normal_map_texture_index, i0, i1, i2, interpolated_texcoords, surface_normal);
    }
    return surface_normal;
}

/* sample_wavelength_uniformly (Dispersion.h:16-17,469-474): integer constants, so (830 - 360) is an int */
B2_HD float sample_wavelength_uniformly(Rng& rng)
{
    float r = rng.next_float();
    return r * (830 - 360) + 360;
}

template <class V>
B2_HD int effective_dielectric_priority(const Mat& m)
{
    return (V::BSDF_OVERRIDE == B2_BSDF_LAMBERTIAN || V::BSDF_OVERRIDE == B2_BSDF_OREN_NAYAR) ? 0 : (int)m.dielectric_priority;
}

/* `first`: the closest hit of (origin, direction) when the caller already traced it (the pooled tracer of the persistent
 * MegaKernel); null = trace here. Re-traces after a skipped nested-dielectric boundary are always traced here. */
template <class V>
B2_HD bool trace_main_path_ray(const Ctx& ctx, V3 origin, V3 direction, Payload& payload, HitInfo& out_hit, int last_hit_primitive_index, Rng& rng,
                               const B2Hit* first = nullptr)
{
    const b2pt_render_data& rd = *ctx.rd;
    B2Hit hit;
    bool skipping_volume_boundary = false;
    int boundary_skips = 0;
    do
    {
        if (++boundary_skips > B2_BOUNDARY_SKIP_GUARD)
        {
            b2_report_fault(ctx.faults, B2_FAULT_BOUNDARY_GUARD);
            return false;
        }
        if (first != nullptr && boundary_skips == 1)
            hit = *first;
        else
            hit = b2_trace(ctx, false, origin, direction, 3.0e38f, last_hit_primitive_index, /* main_path_ray = */ true);
        if (hit.prim < 0)
            return false;

        const int i0 = rd.buffers.triangles_indices[hit.prim * 3 + 0];
        const int i1 = rd.buffers.triangles_indices[hit.prim * 3 + 1];
        const int i2 = rd.buffers.triangles_indices[hit.prim * 3 + 2];

        out_hit.inter_point = origin + hit.t * direction;
        out_hit.primitive_index = hit.prim;
        {
            const b2pt_float2 tA = rd.buffers.texcoords[i0], tB = rd.buffers.texcoords[i1], tC = rd.buffers.texcoords[i2];
            out_hit.texcoords = v2(tB.x, tB.y) * hit.u + v2(tC.x, tC.y) * hit.v + v2(tA.x, tA.y) * (1.0f - hit.u - hit.v);
        }
        /* hiprtHit::normal is already normalize(cross(e1, e2)) in the CPU path (Triangle.h:53) and
         * trace_main_path_ray normalises it once more (Intersect.h:235) */
        out_hit.geometric_normal = normalize(normalize(hit.ng));
        out_hit.shading_normal = shading_normal_at(rd, out_hit.geometric_normal, i0, i1, i2, hit.u, hit.v, hit.prim, out_hit.texcoords);
        out_hit.t = hit.t;

        const int material_index = rd.buffers.material_indices[hit.prim];
        /* scenes without a single texture keep the shorter untextured gather (same values: every texture index is NO_TEXTURE) */
        if (rd.buffers.material_textures != nullptr)
            payload.material = load_scene_material_textured(rd.buffers.materials_buffer, rd.buffers.material_textures, material_index, out_hit.texcoords, rd.bsdfs_data.white_furnace_mode != 0);
        else
            payload.material = load_scene_material(rd.buffers.materials_buffer, material_index, rd.bsdfs_data.white_furnace_mode != 0);

        skipping_volume_boundary = payload.volume_state.interior_stack.push(
            payload.volume_state.incident_mat_index, payload.volume_state.outgoing_mat_index, payload.volume_state.inside_material,
            material_index, effective_dielectric_priority<V>(payload.material));

        if (payload.volume_state.inside_material)
            payload.volume_state.distance_in_volume += hit.t;

        fix_backfacing_normals(out_hit, -direction);

        if (skipping_volume_boundary)
        {
            origin = out_hit.inter_point;
            payload.volume_state.distance_in_volume += hit.t;
            /* NB: like the reference, last_hit_primitive_index is NOT updated when re-tracing (Intersect.h:250-257) */
        }
    } while (skipping_volume_boundary);

    if (payload.material.dispersion_scale > 0.0f && payload.material.specular_transmission > 0.0f && payload.volume_state.sampled_wavelength == 0.0f)
        payload.volume_state.sampled_wavelength = -sample_wavelength_uniformly(rng);

    return true;
}

/* true when something lies strictly before t_max - 1e-4 (opaque scenes) */
B2_HD bool shadow_ray_occluded(const Ctx& ctx, V3 origin, V3 direction, float t_max, int last_hit_primitive_index)
{
    B2Hit hit = b2_trace(ctx, true, origin, direction, t_max - 1.0e-4f, last_hit_primitive_index);
    return hit.prim >= 0;
}

struct LightRayHit
{
    int prim;
    int material_index;
    float distance;
    V3 shading_normal;
    V3 geometric_normal;
    Rgb emission;
};

/* closest hit along a BSDF-sampled direction, returning what (emissive or not) was hit; t_max is 1e35 at every call site */
B2_HD bool light_ray_hit_from(const b2pt_render_data& rd, const B2Hit& hit, float t_max, LightRayHit& out);
B2_HD bool bsdf_light_sample_ray(const Ctx& ctx, V3 origin, V3 direction, float t_max, LightRayHit& out, int last_hit_primitive_index)
{
    B2Hit hit = b2_trace(ctx, false, origin, direction, 3.0e38f, last_hit_primitive_index);
    return light_ray_hit_from(*ctx.rd, hit, t_max, out);
}

/* what bsdf_light_sample_ray derives from the closest hit of its ray */
B2_HD bool light_ray_hit_from(const b2pt_render_data& rd, const B2Hit& hit, float t_max, LightRayHit& out)
{
    if (hit.prim < 0)
        return false;
    if (!(hit.t < t_max - 1.0e-4f))
        return false;

    const int i0 = rd.buffers.triangles_indices[hit.prim * 3 + 0];
    const int i1 = rd.buffers.triangles_indices[hit.prim * 3 + 1];
    const int i2 = rd.buffers.triangles_indices[hit.prim * 3 + 2];
    const int material_index = rd.buffers.material_indices[hit.prim];
    V2 interpolated_texcoords;
    {
        const b2pt_float2 tA = rd.buffers.texcoords[i0], tB = rd.buffers.texcoords[i1], tC = rd.buffers.texcoords[i2];
        interpolated_texcoords = v2(tB.x, tB.y) * hit.u + v2(tC.x, tC.y) * hit.v + v2(tA.x, tA.y) * (1.0f - hit.u - hit.v);
    }
    out.emission = load_rgb(rd.buffers.materials_buffer.emission, material_index);
    if (rd.buffers.material_textures != nullptr)
    {
        /* Intersect.h:620-623: any emission texture index but NO_TEXTURE reads the texture (get_material_property returns black for
         * the CONSTANT_EMISSIVE marker) */
        const int emission_texture_index = (int)(rd.buffers.materials_buffer.normal_map_emission_index[material_index] >> 16);
        if (emission_texture_index != B2_NO_TEXTURE)
        {
            out.emission = rgb(0.0f);
            if (emission_texture_index != B2_CONSTANT_EMISSIVE_TEXTURE)
            {
                float rgba[4];
                sample_texture_rgba(rd.buffers.material_textures, emission_texture_index, false, interpolated_texcoords, rgba);
                out.emission = rgb(rgba[0], rgba[1], rgba[2]);
            }
        }
    }
    out.geometric_normal = normalize(normalize(hit.ng));
    out.shading_normal = shading_normal_at(rd, out.geometric_normal, i0, i1, i2, hit.u, hit.v, hit.prim, interpolated_texcoords);
    out.prim = hit.prim;
    out.material_index = material_index;
    out.distance = hit.t;
    return true;
}

#endif
