/*
 * b2_lights.h -- next-event estimation on emissive triangles: light selection (uniform / power alias table),
 * point-on-triangle sampling, pdf conversions and the one-light / BSDF / MIS strategies.
 *
 * Restates /root/reference/src/Device/includes/LightSampling/LightUtils.h:51-135,142-199,248-309,1377-1390,
 * LightSampling/PDFConversion.h:19-46, AliasTable.h:13-22 and LightSampling/Lights.h:25-226. RNG draw order per
 * strategy is the reference's: uniform = index, 2 point draws; power = alias index, alias
 * coin, 2 point draws; MIS = light sample first, then the BSDF sample.
 * RIS lives in b2_ris.h, ReSTIR DI in b2_restir_di.h, the environment map in b2_envmap.h.
 */
#ifndef B2_LIGHTS_H
#define B2_LIGHTS_H

#include "b2_bsdf.h"

struct LightSample
{
    int emissive_triangle_index;
    V3 light_source_normal;
    float light_area;
    Rgb emission;
    V3 point_on_light;
    float area_measure_pdf;
    int incident_light_info;
};

B2_HD LightSample light_sample_none()
{
    LightSample s;
    s.emissive_triangle_index = -1;
    s.light_source_normal = v3(0.0f, 1.0f, 0.0f);
    s.light_area = 1.0f;
    s.emission = rgb(0.0f);
    s.point_on_light = v3(0.0f);
    s.area_measure_pdf = 0.0f;
    s.incident_light_info = B2_LOBE_NO_INFO;
    return s;
}

/* DirectLightSamplingAllowBackfacingLights = FALSE (DirectLightSamplingOptions.h:122) */
B2_HD float cosine_term_at_light_source(V3 light_source_normal, V3 minus_direction_to_light)
{
    return b2_max(0.0f, dot(light_source_normal, minus_direction_to_light));
}

B2_HD float area_to_solid_angle_pdf(float area_pdf, float distance, float cos_theta_at_light_source)
{
    if (cos_theta_at_light_source < 1.0e-8f)
        return 0.0f;
    return area_pdf * b2_square(distance) / cos_theta_at_light_source;
}

B2_HD float solid_angle_to_area_pdf(float solid_angle_pdf, float distance, float cos_theta_at_light_source)
{
    if (cos_theta_at_light_source < 1.0e-8f)
        return 0.0f;
    return solid_angle_pdf / b2_square(distance) * cos_theta_at_light_source;
}

B2_HD int alias_table_sample(const b2pt_alias_table& table, Rng& rng)
{
    int random_index = rng.next_index((int)table.size);
    float probability = table.alias_table_probas[random_index];
    if (rng.next_float() > probability)
        random_index = table.alias_table_alias[random_index];
    return random_index;
}

B2_HD Rgb clamp_light_contribution(Rgb light_contribution, float clamp_max_value, bool clamp_condition)
{
    if (!has_nan(light_contribution) && clamp_max_value > 0.0f && clamp_condition)
        light_contribution = rgb_clamp(light_contribution, -clamp_max_value, clamp_max_value);
    return light_contribution;
}

/* Heitz 2019 low-distortion square -> triangle map (LightUtils.h:142-156) */
B2_HD void square_to_triangle(float& x, float& y)
{
    if (y > x)
    {
        x *= 0.5f;
        y -= x;
    }
    else
    {
        y *= 0.5f;
        x -= y;
    }
}

B2_HD void sample_point_on_triangle(const b2pt_render_data& rd, int triangle_index, Rng& rng, V3& out_point, V3& out_normal, float& out_area)
{
    V3 A = load_v3(rd.buffers.vertices_positions, rd.buffers.triangles_indices[triangle_index * 3 + 0]);
    V3 B = load_v3(rd.buffers.vertices_positions, rd.buffers.triangles_indices[triangle_index * 3 + 1]);
    V3 C = load_v3(rd.buffers.vertices_positions, rd.buffers.triangles_indices[triangle_index * 3 + 2]);

    float u = rng.next_float();
    float v = rng.next_float();
    square_to_triangle(u, v);

    V3 AB = B - A;
    V3 AC = C - A;
    V3 normal = cross(AB, AC);
    float length_normal = length(normal);

    out_point = A + AB * u + AC * v;
    out_normal = normal / length_normal;
    out_area = 0.5f * length_normal;
}

template <class V>
B2_HD LightSample sample_one_emissive_triangle(const b2pt_render_data& rd, Rng& rng)
{
    if (rd.buffers.emissive_triangles_count == 0)
        return light_sample_none();

    int emissive_index;
    if constexpr (V::BASE == B2_BASE_UNIFORM)
        emissive_index = rng.next_index(rd.buffers.emissive_triangles_count);
    else
        emissive_index = alias_table_sample(rd.buffers.emissives_power_alias_table, rng);
    int triangle_index = rd.buffers.emissive_triangles_primitive_indices[emissive_index];

    LightSample s = light_sample_none();
    sample_point_on_triangle(rd, triangle_index, rng, s.point_on_light, s.light_source_normal, s.light_area);
    s.emissive_triangle_index = triangle_index;
    s.emission = load_rgb(rd.buffers.materials_buffer.emission, rd.buffers.material_indices[triangle_index]);

    s.area_measure_pdf = 1.0f / s.light_area;
    if constexpr (V::BASE == B2_BASE_UNIFORM)
        s.area_measure_pdf /= rd.buffers.emissive_triangles_count;
    else
        s.area_measure_pdf *= (luminance(s.emission) * s.light_area) / rd.buffers.emissives_power_alias_table.sum_elements;
    return s;
}

template <class V>
B2_HD float pdf_of_emissive_triangle_hit_area_measure(const b2pt_render_data& rd, float light_area, Rgb light_emission)
{
    float area_measure_pdf = 1.0f / light_area;
    if constexpr (V::BASE == B2_BASE_UNIFORM)
        area_measure_pdf /= rd.buffers.emissive_triangles_count;
    else
        area_measure_pdf *= (luminance(light_emission) * light_area) / rd.buffers.emissives_power_alias_table.sum_elements;
    return area_measure_pdf;
}

template <class V>
B2_HD float pdf_of_emissive_triangle_hit_solid_angle(const b2pt_render_data& rd, const LightRayHit& light_hit, V3 to_light_direction)
{
    float cosine_light_source = cosine_term_at_light_source(light_hit.geometric_normal, -to_light_direction);
    float pdf_area_measure = pdf_of_emissive_triangle_hit_area_measure<V>(rd, rd.buffers.triangles_areas[light_hit.prim], light_hit.emission);
    return area_to_solid_angle_pdf(pdf_area_measure, light_hit.distance, cosine_light_source);
}

/* MaterialUtils::can_do_light_sampling: with the default macros (delta-distribution optimisation ON, microfacet
 * regularisation ON) it is constant true (MaterialUtils.h:62-83) */
B2_HD bool can_do_light_sampling(const Mat&) { return true; }

/* Lights.h:25-87 */
template <class V>
B2_HD Rgb sample_one_light_no_MIS(const Ctx& ctx, Payload& payload, const HitInfo& hit, V3 view_direction, Rng& rng)
{
    const b2pt_render_data& rd = *ctx.rd;
    if (!can_do_light_sampling(payload.material))
        return rgb(0.0f);

    LightSample light_sample = sample_one_emissive_triangle<V>(rd, rng);
    if (light_sample.area_measure_pdf <= 0.0f)
        return rgb(0.0f);

    V3 shadow_ray_origin = hit.inter_point;
    V3 shadow_ray_direction = light_sample.point_on_light - shadow_ray_origin;
    float distance_to_light = length(shadow_ray_direction);
    V3 shadow_dir = shadow_ray_direction / distance_to_light;

    Rgb light_source_radiance = rgb(0.0f);
    float dot_light_source = cosine_term_at_light_source(light_sample.light_source_normal, -shadow_dir);
    if (dot_light_source > 0.0f)
    {
        bool in_shadow = shadow_ray_occluded(ctx, shadow_ray_origin, shadow_dir, distance_to_light, hit.primitive_index);
        if (!in_shadow)
        {
            float bsdf_pdf_value;
            int incident_light_info = light_sample.incident_light_info;
            BsdfCtx bctx = make_bsdf_ctx(view_direction, hit.shading_normal, hit.geometric_normal, shadow_dir, &incident_light_info, &payload.volume_state, false,
                                         &payload.material, payload.bounce, payload.accumulated_roughness, B2_REGULARIZATION_CLASSIC);
            Rgb bsdf_color = bsdf_eval<V>(rd, bctx, bsdf_pdf_value);
            if (bsdf_pdf_value != 0.0f)
            {
                float light_sample_solid_angle_pdf = area_to_solid_angle_pdf(light_sample.area_measure_pdf, distance_to_light, dot_light_source);
                if (light_sample_solid_angle_pdf > 0.0f)
                {
                    float cosine_term = fabsf(dot(hit.shading_normal, shadow_dir));
                    /* "/ unoccluded_probability" of NEE++ is 1 (NEE++ off, Intersect.h:389-391) */
                    light_source_radiance = light_sample.emission * cosine_term * bsdf_color / light_sample_solid_angle_pdf / 1.0f;
                }
            }
        }
    }
    return light_source_radiance;
}

/* Lights.h:89-121 */
template <class V>
B2_HD Rgb sample_one_light_bsdf(const Ctx& ctx, Payload& payload, const HitInfo& hit, V3 view_direction, Rng& rng)
{
    const b2pt_render_data& rd = *ctx.rd;
    float bsdf_sample_pdf;
    V3 sampled_direction;
    int incident_light_info = B2_LOBE_NO_INFO;
    BsdfCtx bctx = make_bsdf_ctx(view_direction, hit.shading_normal, hit.geometric_normal, v3(0.0f), &incident_light_info, &payload.volume_state, false,
                                 &payload.material, payload.bounce, payload.accumulated_roughness, B2_REGULARIZATION_CLASSIC);
    Rgb bsdf_color = bsdf_sample<V>(rd, bctx, sampled_direction, bsdf_sample_pdf, rng);

    Rgb bsdf_radiance = rgb(0.0f);
    if (bsdf_sample_pdf > 0.0f)
    {
        LightRayHit light_hit;
        bool found = bsdf_light_sample_ray(ctx, hit.inter_point, sampled_direction, 1.0e35f, light_hit, hit.primitive_index);
        if (found && !is_black(light_hit.emission) && cosine_term_at_light_source(light_hit.geometric_normal, -sampled_direction) > 0.0f)
        {
            float cosine_term = fabsf(dot(hit.shading_normal, sampled_direction));
            bsdf_radiance = bsdf_color * cosine_term * light_hit.emission / bsdf_sample_pdf;
        }
    }
    return bsdf_radiance;
}

/* Lights.h:123-226 */
template <class V>
B2_HD Rgb sample_one_light_MIS(const Ctx& ctx, Payload& payload, const HitInfo& hit, V3 view_direction, Rng& rng)
{
    const b2pt_render_data& rd = *ctx.rd;
    Rgb light_source_radiance_mis = rgb(0.0f);

    if (can_do_light_sampling(payload.material))
    {
        LightSample light_sample = sample_one_emissive_triangle<V>(rd, rng);
        if (light_sample.area_measure_pdf > 0.0f)
        {
            V3 shadow_ray_direction = light_sample.point_on_light - hit.inter_point;
            float distance_to_light = length(shadow_ray_direction);
            V3 shadow_dir = shadow_ray_direction / distance_to_light;

            bool in_shadow = shadow_ray_occluded(ctx, hit.inter_point, shadow_dir, distance_to_light, hit.primitive_index);
            if (!in_shadow)
            {
                float bsdf_pdf_value;
                int incident_light_info = B2_LOBE_NO_INFO;
                BsdfCtx bctx = make_bsdf_ctx(view_direction, hit.shading_normal, hit.geometric_normal, shadow_dir, &incident_light_info, &payload.volume_state, false,
                                             &payload.material, payload.bounce, payload.accumulated_roughness, B2_REGULARIZATION_MIS);
                Rgb bsdf_color = bsdf_eval<V>(rd, bctx, bsdf_pdf_value);
                if (bsdf_pdf_value > 0.0f)
                {
                    float cos_theta_at_light_source = cosine_term_at_light_source(light_sample.light_source_normal, -shadow_dir);
                    if (cos_theta_at_light_source > 1.0e-5f)
                    {
                        float light_sample_solid_angle_pdf = area_to_solid_angle_pdf(light_sample.area_measure_pdf, distance_to_light, cos_theta_at_light_source);
                        float mis_weight = balance_heuristic(light_sample_solid_angle_pdf, bsdf_pdf_value);
                        float cosine_term = fabsf(dot(hit.shading_normal, shadow_dir));
                        light_source_radiance_mis = bsdf_color * cosine_term * light_sample.emission * mis_weight / light_sample_solid_angle_pdf / 1.0f;
                    }
                }
            }
        }
    }

    float bsdf_sample_pdf;
    V3 sampled_direction;
    int incident_light_info = B2_LOBE_NO_INFO;
    Rgb bsdf_radiance_mis = rgb(0.0f);
    BsdfCtx bctx = make_bsdf_ctx(view_direction, hit.shading_normal, hit.geometric_normal, v3(0.0f), &incident_light_info, &payload.volume_state, false,
                                 &payload.material, payload.bounce, payload.accumulated_roughness, B2_REGULARIZATION_MIS);
    Rgb bsdf_color = bsdf_sample<V>(rd, bctx, sampled_direction, bsdf_sample_pdf, rng);
    if (bsdf_sample_pdf > 0.0f)
    {
        LightRayHit light_hit;
        bool found = bsdf_light_sample_ray(ctx, hit.inter_point, sampled_direction, 1.0e35f, light_hit, hit.primitive_index);
        if (found && !is_black(light_hit.emission) && cosine_term_at_light_source(light_hit.geometric_normal, -sampled_direction) > 0.0f)
        {
            float light_pdf_solid_angle = pdf_of_emissive_triangle_hit_solid_angle<V>(rd, light_hit, sampled_direction);
            float mis_weight = balance_heuristic(bsdf_sample_pdf, light_pdf_solid_angle);
            float cosine_term = fabsf(dot(hit.shading_normal, sampled_direction));
            bsdf_radiance_mis = bsdf_color * cosine_term * light_hit.emission * mis_weight / bsdf_sample_pdf;
        }
    }

    return light_source_radiance_mis + bsdf_radiance_mis;
}

#endif
