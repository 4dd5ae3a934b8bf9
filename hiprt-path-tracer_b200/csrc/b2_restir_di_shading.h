/*
 * b2_restir_di_shading.h -- ReSTIR DI data types, the SoA reservoir storage, the target function and the final
 * shading of a pixel's reservoir (the part of ReSTIR DI the MegaKernel itself executes at bounce 0).
 *
 * Restates /root/reference/src/Device/includes/ReSTIR/DI/Reservoir.h:22-121 (sample, reservoir, WRS update,
 * combine_with, end, end_with_normalization), DI/SampleFlags.h:11-33, ReSTIR/Surface.h:13-80 (pixel surface from the
 * G-buffer), DI/TargetFunction.h:12-98 and DI/FinalShading.h:16-125.
 *
 * Storage. The reference keeps reservoirs as 36-byte AoS records (a21); here a reservoir buffer of the same byte size
 * (width*height*36, allocated by the host exactly as ReSTIRDIRenderPass.h:108-115 does) is read and written as NINE
 * PLANES of 4-byte words -- [M | weight_sum | UCW | point.x | point.y | point.z | triangle | target_function | flags],
 * plane p of pixel i at word p * (width*height) + i -- so that a warp's 32 pixels touch 32 consecutive words per
 * field (coalesced 128-byte transactions) instead of 32 records 36 bytes apart. The host never looks inside these
 * buffers (it only ping-pongs pointers), so the drop-in contract is unaffected.
 */
#ifndef B2_RESTIR_DI_SHADING_H
#define B2_RESTIR_DI_SHADING_H

#include "b2_envmap.h"

enum
{
    B2_DI_FLAGS_NONE = 0,
    B2_DI_FLAGS_ENVMAP_SAMPLE = 1 << 0,
    B2_DI_FLAGS_UNOCCLUDED = 1 << 7
};

struct DiSample
{
    V3 point_on_light_source;
    int emissive_triangle_index;
    float target_function;
    uint32_t flags; /* an unsigned char in the reference: every update is masked to 8 bits */

    B2_HD bool is_envmap_sample() const { return (flags & B2_DI_FLAGS_ENVMAP_SAMPLE) != 0; }
    /* bits 1..6 = coat, first metal, second metal, specular, glass reflect, glass refract (SampleFlags.h / Reservoir.h:46-49) */
    B2_HD int lobe_info() const { return (int)(flags & (0x3Fu << 1)); }
    B2_HD void or_flags(int f) { flags = (flags | (uint32_t)f) & 0xFFu; }
};

B2_HD DiSample di_sample_defaults()
{
    DiSample s;
    s.point_on_light_source = v3(0.0f);
    s.emissive_triangle_index = -1;
    s.target_function = 0.0f;
    s.flags = B2_DI_FLAGS_NONE;
    return s;
}

struct DiReservoir
{
    int M;
    float weight_sum;
    float UCW;
    DiSample sample;

    B2_HD void init()
    {
        M = 0;
        weight_sum = 0.0f;
        UCW = 0.0f;
        sample = di_sample_defaults();
    }

    B2_HD void add_one_candidate(const DiSample& new_sample, float weight, Rng& rng)
    {
        M++;
        weight_sum += weight;
        if (rng.next_float() < weight / weight_sum)
            sample = new_sample;
    }

    B2_HD bool combine_with(const DiReservoir& other, float mis_weight, float target_function, float jacobian_determinant, Rng& rng)
    {
        if (other.UCW <= 0.0f)
        {
            M += other.M;
            return false;
        }
        float reservoir_sample_weight = mis_weight * target_function * other.UCW * jacobian_determinant;
        M += other.M;
        weight_sum += reservoir_sample_weight;
        if (rng.next_float() < reservoir_sample_weight / weight_sum)
        {
            sample = other.sample;
            sample.target_function = target_function;
            return true;
        }
        return false;
    }

    /* combine_with with its random number drawn beforehand (callers draw it exactly when other.UCW > 0, at the stream position of
     * the reference's draw) */
    B2_HD bool combine_with_drawn(const DiReservoir& other, float mis_weight, float target_function, float jacobian_determinant, float draw)
    {
        if (other.UCW <= 0.0f)
        {
            M += other.M;
            return false;
        }
        float reservoir_sample_weight = mis_weight * target_function * other.UCW * jacobian_determinant;
        M += other.M;
        weight_sum += reservoir_sample_weight;
        if (draw < reservoir_sample_weight / weight_sum)
        {
            sample = other.sample;
            sample.target_function = target_function;
            return true;
        }
        return false;
    }

    B2_HD void end()
    {
        if (weight_sum == 0.0f)
            UCW = 0.0f;
        else
            UCW = 1.0f / sample.target_function * weight_sum;
    }

    B2_HD void end_with_normalization(float normalization_numerator, float normalization_denominator)
    {
        if (weight_sum == 0.0f || weight_sum < 1.0e-10f || weight_sum > 1.0e10f || normalization_denominator == 0.0f || normalization_numerator == 0.0f)
            UCW = 0.0f;
        else
            UCW = 1.0f / sample.target_function * weight_sum * normalization_numerator / normalization_denominator;
        M = b2_min(M, 1000000);
    }
};

/* ---- SoA planes ---- */
enum { DI_PLANE_M = 0, DI_PLANE_WEIGHT_SUM, DI_PLANE_UCW, DI_PLANE_PX, DI_PLANE_PY, DI_PLANE_PZ, DI_PLANE_TRIANGLE, DI_PLANE_TARGET, DI_PLANE_FLAGS, DI_PLANE_COUNT };

B2_HD size_t di_plane_stride(const b2pt_render_data& rd) { return (size_t)rd.render_settings.render_resolution.x * (size_t)rd.render_settings.render_resolution.y; }

B2_HD DiReservoir di_load(const b2pt_render_data& rd, const void* buffer, int pixel_index)
{
    const uint32_t* w = reinterpret_cast<const uint32_t*>(buffer) + pixel_index;
    const size_t n = di_plane_stride(rd);
    DiReservoir r;
    r.M = (int)w[DI_PLANE_M * n];
    r.weight_sum = b2_uint_as_float(w[DI_PLANE_WEIGHT_SUM * n]);
    r.UCW = b2_uint_as_float(w[DI_PLANE_UCW * n]);
    r.sample.point_on_light_source = v3(b2_uint_as_float(w[DI_PLANE_PX * n]), b2_uint_as_float(w[DI_PLANE_PY * n]), b2_uint_as_float(w[DI_PLANE_PZ * n]));
    r.sample.emissive_triangle_index = (int)w[DI_PLANE_TRIANGLE * n];
    r.sample.target_function = b2_uint_as_float(w[DI_PLANE_TARGET * n]);
    r.sample.flags = w[DI_PLANE_FLAGS * n] & 0xFFu;
    return r;
}

B2_HD int di_load_M(const b2pt_render_data& rd, const void* buffer, int pixel_index)
{
    return (int)(reinterpret_cast<const uint32_t*>(buffer) + pixel_index)[DI_PLANE_M * di_plane_stride(rd)];
}

B2_HD void di_store(const b2pt_render_data& rd, void* buffer, int pixel_index, const DiReservoir& r)
{
    uint32_t* w = reinterpret_cast<uint32_t*>(buffer) + pixel_index;
    const size_t n = di_plane_stride(rd);
    w[DI_PLANE_M * n] = (uint32_t)r.M;
    w[DI_PLANE_WEIGHT_SUM * n] = b2_float_as_uint(r.weight_sum);
    w[DI_PLANE_UCW * n] = b2_float_as_uint(r.UCW);
    w[DI_PLANE_PX * n] = b2_float_as_uint(r.sample.point_on_light_source.x);
    w[DI_PLANE_PY * n] = b2_float_as_uint(r.sample.point_on_light_source.y);
    w[DI_PLANE_PZ * n] = b2_float_as_uint(r.sample.point_on_light_source.z);
    w[DI_PLANE_TRIANGLE * n] = (uint32_t)r.sample.emissive_triangle_index;
    w[DI_PLANE_TARGET * n] = b2_float_as_uint(r.sample.target_function);
    w[DI_PLANE_FLAGS * n] = r.sample.flags & 0xFFu;
}

/* ---- pixel surface (ReSTIR/Surface.h) ---- */
struct DiSurface
{
    Mat material;
    VolumeState ray_volume_state;
    int primitive_index;
    V3 view_direction;
    V3 shading_normal;
    V3 geometric_normal;
    V3 shading_point;
};

template <class V>
B2_HD void reconstruct_first_hit(VolumeState& vs, const Mat& material, const int* material_indices, int primitive_index, Rng& rng);
B2_HD PackedMat load_packed_material(const void* buffer, uint32_t pixel_index);

template <class V>
B2_HD DiSurface get_pixel_surface(const b2pt_render_data& rd, int pixel_index, bool previous_frame, Rng& rng)
{
    const b2pt_gbuffer& g = previous_frame ? rd.g_buffer_prev_frame : rd.g_buffer;
    DiSurface s;
    s.material = unpack_material(load_packed_material(g.materials, (uint32_t)pixel_index));
    s.primitive_index = g.first_hit_prim_index[pixel_index];
    s.ray_volume_state.init();
    reconstruct_first_hit<V>(s.ray_volume_state, s.material, rd.buffers.material_indices, s.primitive_index, rng);
    /* the reference reads the CURRENT G-buffer position for the view direction in both cases (Surface.h:58) */
    const b2pt_camera& cam = previous_frame ? rd.prev_camera : rd.current_camera;
    s.view_direction = normalize(v3(cam.position.x, cam.position.y, cam.position.z) - load_v3(rd.g_buffer.primary_hit_position, pixel_index));
    s.shading_normal = unpack_normal_oct24(g.shading_normals[pixel_index]);
    s.geometric_normal = unpack_normal_oct24(g.geometric_normals[pixel_index]);
    s.shading_point = load_v3(g.primary_hit_position, pixel_index);
    return s;
}

B2_HD V3 triangle_normal_not_normalized(const b2pt_render_data& rd, int triangle_index)
{
    /* A reservoir whose weights went NaN (0/0 candidate weight, see InitialCandidates.h:405-413 with a back-facing hit)
     * keeps the default sample (triangle -1) but passes the `UCW <= 0` tests; the reference then reads
     * triangles_indices[-3..-1] / material_indices[-1] (undefined, harmless on its host heap) and ends in a NaN ->
     * pink pixel all the same. Never read out of bounds on the device: the NaN UCW still poisons the product. */
    if (triangle_index < 0)
        return v3(0.0f);
    V3 A = load_v3(rd.buffers.vertices_positions, rd.buffers.triangles_indices[triangle_index * 3 + 0]);
    V3 B = load_v3(rd.buffers.vertices_positions, rd.buffers.triangles_indices[triangle_index * 3 + 1]);
    V3 C = load_v3(rd.buffers.vertices_positions, rd.buffers.triangles_indices[triangle_index * 3 + 2]);
    return cross(B - A, C - A);
}

B2_HD V3 di_light_sample_direction(const b2pt_render_data& rd, const DiSample& sample, V3 surface_shading_point, float& out_distance_to_light)
{
    V3 sample_direction;
    if (sample.is_envmap_sample())
    {
        sample_direction = mat3_mul_vec(rd.world_settings.envmap_to_world_matrix, sample.point_on_light_source);
        out_distance_to_light = 1.0e35f;
    }
    else
    {
        sample_direction = sample.point_on_light_source - surface_shading_point;
        out_distance_to_light = length(sample_direction);
        sample_direction = sample_direction / out_distance_to_light;
    }
    return sample_direction;
}

template <class V>
B2_HD Rgb di_light_sample_emission(const b2pt_render_data& rd, const DiSample& sample, V3 sample_direction)
{
    if (sample.is_envmap_sample())
    {
        float envmap_pdf;
        return envmap_eval<V>(rd.world_settings, sample_direction, envmap_pdf);
    }
    if (sample.emissive_triangle_index < 0)
        return rgb(0.0f); /* see triangle_normal_not_normalized */
    int material_index = rd.buffers.material_indices[sample.emissive_triangle_index];
    return load_rgb(rd.buffers.materials_buffer.emission, material_index);
}

/* ReSTIR_DI_evaluate_target_function<withVisibility>, DI/TargetFunction.h:48-98 */
/* known_bsdf_color: the value bsdf_eval returns for exactly this context (same surface, direction, lobe info NO_INFO, bounce 0,
 * accumulated roughness 0, no regularisation), when the caller has just computed it: the evaluation is a pure function of its
 * inputs, so the initial-candidates loop does not run it twice per candidate as the reference does (InitialCandidates.h:226-240) */
template <class V, bool WITH_VISIBILITY>
B2_HD float di_target_function(const Ctx& ctx, const DiSample& sample, DiSurface& surface, const Rgb* known_bsdf_color = nullptr)
{
    const b2pt_render_data& rd = *ctx.rd;
    if (sample.emissive_triangle_index == -1 && !sample.is_envmap_sample())
        return 0.0f;

    float bsdf_pdf_value;
    float distance_to_light;
    V3 sample_direction = di_light_sample_direction(rd, sample, surface.shading_point, distance_to_light);

    float cosine_term = dot(surface.shading_normal, sample_direction);
    if (cosine_term <= 0.0f)
        return 0.0f;

    int incident_light_info = sample.lobe_info();
    BsdfCtx bctx = make_bsdf_ctx(surface.view_direction, surface.shading_normal, surface.geometric_normal, sample_direction, &incident_light_info, &surface.ray_volume_state, false,
                                 &surface.material, 0, 0.0f);
    Rgb bsdf_color = known_bsdf_color != nullptr ? *known_bsdf_color : bsdf_eval<V>(rd, bctx, bsdf_pdf_value);
    Rgb sample_emission = di_light_sample_emission<V>(rd, sample, sample_direction);

    float geometry_term = 1.0f;
    if (!sample.is_envmap_sample())
    {
        V3 emissive_triangle_normal = normalize(triangle_normal_not_normalized(rd, sample.emissive_triangle_index));
        geometry_term = cosine_term_at_light_source(emissive_triangle_normal, -sample_direction);
        geometry_term /= b2_square(distance_to_light);
    }

    float target_function = luminance(bsdf_color * sample_emission * cosine_term * geometry_term);
    if (target_function == 0.0f)
        return 0.0f;

    if (WITH_VISIBILITY)
    {
        bool visible = !shadow_ray_occluded(ctx, surface.shading_point, sample_direction, distance_to_light, surface.primitive_index);
        target_function *= visible ? 1.0f : 0.0f;
    }
    return target_function;
}

/* evaluate_ReSTIR_DI_reservoir once the sample's visibility is known (DI/FinalShading.h:44-99) */
template <class V>
B2_HD Rgb restir_di_shade_sample(const Ctx& ctx, Payload& payload, const HitInfo& hit, V3 view_direction, const DiReservoir& reservoir, V3 shadow_ray_direction, float distance_to_light,
                                 bool in_shadow)
{
    const b2pt_render_data& rd = *ctx.rd;
    const DiSample& sample = reservoir.sample;
    Rgb final_color = rgb(0.0f);
    if (!in_shadow)
    {
        float bsdf_pdf_value;
        int incident_light_info = sample.lobe_info();
        BsdfCtx bctx = make_bsdf_ctx(view_direction, hit.shading_normal, hit.geometric_normal, shadow_ray_direction, &incident_light_info, &payload.volume_state, false, &payload.material,
                                     0, 0.0f, B2_REGULARIZATION_MIS);
        Rgb bsdf_color = bsdf_eval<V>(rd, bctx, bsdf_pdf_value);

        float cosine_at_evaluated_point = dot(hit.shading_normal, shadow_ray_direction);
        if (sample.flags & B2_LOBE_GLASS_REFRACT)
            cosine_at_evaluated_point = fabsf(cosine_at_evaluated_point);

        if (cosine_at_evaluated_point > 0.0f)
        {
            Rgb sample_emission = di_light_sample_emission<V>(rd, sample, shadow_ray_direction);
            float area_measure_to_solid_angle_conversion;
            if (sample.is_envmap_sample())
                area_measure_to_solid_angle_conversion = 1.0f;
            else
            {
                V3 emissive_triangle_normal = normalize(triangle_normal_not_normalized(rd, sample.emissive_triangle_index));
                area_measure_to_solid_angle_conversion = cosine_term_at_light_source(emissive_triangle_normal, -shadow_ray_direction);
                area_measure_to_solid_angle_conversion /= b2_square(distance_to_light);
            }
            final_color = bsdf_color * reservoir.UCW * sample_emission * cosine_at_evaluated_point * area_measure_to_solid_angle_conversion;
        }
    }
    return final_color;
}

/* evaluate_ReSTIR_DI_reservoir, DI/FinalShading.h:16-101 */
template <class V>
B2_HD Rgb evaluate_ReSTIR_DI_reservoir(const Ctx& ctx, Payload& payload, const HitInfo& hit, V3 view_direction, const DiReservoir& reservoir)
{
    const b2pt_render_data& rd = *ctx.rd;
    if (reservoir.UCW <= 0.0f)
        return rgb(0.0f);

    const DiSample& sample = reservoir.sample;
    float distance_to_light;
    V3 shadow_ray_direction = di_light_sample_direction(rd, sample, hit.inter_point, distance_to_light);

    bool in_shadow = false;
    if (sample.flags & B2_DI_FLAGS_UNOCCLUDED)
        in_shadow = false;
    else if (rd.render_settings.restir_di_settings.do_final_shading_visibility)
        in_shadow = shadow_ray_occluded(ctx, hit.inter_point, shadow_ray_direction, distance_to_light, hit.primitive_index);
    return restir_di_shade_sample<V>(ctx, payload, hit, view_direction, reservoir, shadow_ray_direction, distance_to_light, in_shadow);
}

/* sample_light_ReSTIR_DI, DI/FinalShading.h:103-125 (validate_reservoir writes the killed UCW back, like the reference's
 * reference-typed access to the output buffer) */
template <class V>
B2_HD Rgb sample_light_ReSTIR_DI(const Ctx& ctx, Payload& payload, const HitInfo& hit, V3 view_direction, Rng& rng, int x, int y)
{
    (void)rng;
    const b2pt_render_data& rd = *ctx.rd;
    int pixel_index = x + y * rd.render_settings.render_resolution.x;
    void* buffer = rd.render_settings.restir_di_settings.restir_output_reservoirs;
    DiReservoir reservoir = di_load(rd, buffer, pixel_index);
    if (reservoir.sample.is_envmap_sample() && rd.world_settings.ambient_light_type != 2)
    {
        reservoir.UCW = 0.0f;
        reinterpret_cast<uint32_t*>(buffer)[DI_PLANE_UCW * di_plane_stride(rd) + pixel_index] = b2_float_as_uint(0.0f);
    }
    return evaluate_ReSTIR_DI_reservoir<V>(ctx, payload, hit, view_direction, reservoir);
}

#endif
