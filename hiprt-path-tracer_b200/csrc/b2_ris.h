/*
 * b2_ris.h -- next-event estimation by resampled importance sampling (RIS / weighted reservoir sampling) over light
 * and BSDF candidates with balance-heuristic MIS: the reference's default direct-lighting strategy
 * (LSS_RIS_BSDF_AND_LIGHT) and the later-bounce strategy under ReSTIR DI.
 *
 * Restates /root/reference/src/Device/includes/RIS/RIS.h:18-81 (shading of the surviving sample, 1 shadow ray),
 * :83-250 (candidate generation: number_of_light_candidates light samples, number_of_bsdf_candidates BSDF samples
 * with one closest-hit ray each), :252-262, and RIS/RIS_Reservoir.h:20-50,108-114. RNG draw order (SURVEY.md
 * appendix C): per light candidate = light sample (3 or 4 draws) then the WRS draw, which happens even when the
 * candidate weight is 0; per BSDF candidate = BSDF sample draws, then the WRS draw.
 * RISUseVisiblityTargetFunction = FALSE and NEE++ off (reference defaults).
 */
#ifndef B2_RIS_H
#define B2_RIS_H

#include "b2_envmap.h"

struct RisSample
{
    Rgb emission;
    V3 point_on_light_source;
    float target_function;
    bool is_bsdf_sample;
    Rgb bsdf_sample_contribution;
    float bsdf_sample_cosine_term;
};

B2_HD RisSample ris_sample_defaults()
{
    RisSample s;
    s.emission = rgb(0.0f);
    s.point_on_light_source = v3(0.0f);
    s.target_function = 0.0f;
    s.is_bsdf_sample = false;
    s.bsdf_sample_contribution = rgb(0.0f);
    s.bsdf_sample_cosine_term = 0.0f;
    return s;
}

struct RisReservoir
{
    unsigned int M;
    float weight_sum;
    float UCW;
    RisSample sample;

    B2_HD void init()
    {
        M = 0;
        weight_sum = 0.0f;
        UCW = 0.0f;
        sample = ris_sample_defaults();
    }

    B2_HD void add_one_candidate(const RisSample& new_sample, float weight, Rng& rng)
    {
        M++;
        weight_sum += weight;
        if (rng.next_float() < weight / weight_sum)
            sample = new_sample;
    }

    B2_HD void end()
    {
        if (weight_sum == 0.0f)
            UCW = 0.0f;
        else
            UCW = 1.0f / sample.target_function * weight_sum;
    }
};

B2_HD bool check_minimum_light_contribution(float minimum_contribution, Rgb contribution)
{
    if (minimum_contribution > 0.0f)
        return !(contribution.r < minimum_contribution && contribution.g < minimum_contribution && contribution.b < minimum_contribution);
    return true;
}

B2_HD bool check_minimum_light_contribution(float minimum_contribution, float contribution)
{
    if (minimum_contribution > 0.0f)
        return !(contribution < minimum_contribution);
    return true;
}

/* One light candidate of the RIS reservoir (RIS.h:103-163): sample, target function, MIS weight, stream into the reservoir. */
template <class V>
B2_HD void ris_light_candidate(const Ctx& ctx, Payload& payload, const HitInfo& hit, V3 view_direction, Rng& rng, int nb_light_candidates, int nb_bsdf_candidates,
                               RisReservoir& reservoir)
{
    const b2pt_render_data& rd = *ctx.rd;
    float target_function = 0.0f;
    float candidate_weight = 0.0f;
    Rgb candidate_bsdf_color = rgb(0.0f);
    LightSample light_sample = sample_one_emissive_triangle<V>(rd, rng);

    if (light_sample.area_measure_pdf > 0.0f)
    {
        V3 to_light_direction = light_sample.point_on_light - hit.inter_point;
        float distance_to_light = length(to_light_direction);
        to_light_direction = to_light_direction / distance_to_light;
        float cosine_at_light_source = cosine_term_at_light_source(light_sample.light_source_normal, -to_light_direction);
        float cosine_at_evaluated_point = fabsf(dot(hit.shading_normal, to_light_direction));
        if (cosine_at_evaluated_point > 0.0f && cosine_at_light_source > 1.0e-6f)
        {
            float bsdf_pdf_value = 0.0f;
            bool contributes_enough = check_minimum_light_contribution(rd.render_settings.minimum_light_contribution, light_sample.emission / light_sample.area_measure_pdf);
            if (!contributes_enough)
                target_function = 0.0f;
            else
            {
                int incident_light_info = B2_LOBE_NO_INFO;
                BsdfCtx bctx = make_bsdf_ctx(view_direction, hit.shading_normal, hit.geometric_normal, to_light_direction, &incident_light_info, &payload.volume_state, false,
                                             &payload.material, payload.bounce, payload.accumulated_roughness, B2_REGULARIZATION_MIS);
                Rgb bsdf_color = bsdf_eval<V>(rd, bctx, bsdf_pdf_value);
                candidate_bsdf_color = bsdf_color;
                Rgb light_contribution = bsdf_color * light_sample.emission * cosine_at_evaluated_point;
                contributes_enough = check_minimum_light_contribution(rd.render_settings.minimum_light_contribution, light_contribution / bsdf_pdf_value / light_sample.area_measure_pdf);
                if (!contributes_enough)
                    target_function = 0.0f;
                else
                    target_function = luminance(light_contribution);
            }
            float solid_angle_light_pdf = area_to_solid_angle_pdf(light_sample.area_measure_pdf, distance_to_light, cosine_at_light_source);
            float mis_weight = balance_heuristic(solid_angle_light_pdf, (float)nb_light_candidates, bsdf_pdf_value, (float)nb_bsdf_candidates);
            candidate_weight = mis_weight * target_function / solid_angle_light_pdf;
        }
    }

    RisSample light_RIS_sample = ris_sample_defaults();
    light_RIS_sample.is_bsdf_sample = false;
    light_RIS_sample.point_on_light_source = light_sample.point_on_light;
    light_RIS_sample.target_function = target_function;
    light_RIS_sample.emission = light_sample.emission;
    /* kept for ris_shade_survivor: a light candidate can only be selected with a weight > 0, i.e. after this evaluation */
    light_RIS_sample.bsdf_sample_contribution = candidate_bsdf_color;
    reservoir.add_one_candidate(light_RIS_sample, candidate_weight, rng);
}

/* The part of a BSDF candidate (RIS.h:165-243) that follows its ray: `hit_found` / `light_hit` are what
 * bsdf_light_sample_ray returned for the sampled direction (hit_found is false when the sample's pdf was 0). */
template <class V>
B2_HD void ris_bsdf_candidate_finish(const b2pt_render_data& rd, const HitInfo& hit, Rng& rng, int nb_light_candidates, int nb_bsdf_candidates, Rgb bsdf_color,
                                     V3 sampled_bsdf_direction, float bsdf_sample_pdf, bool hit_found, const LightRayHit& light_hit, RisReservoir& reservoir)
{
    float target_function = 0.0f;
    float candidate_weight = 0.0f;
    RisSample bsdf_RIS_sample = ris_sample_defaults();
    if (bsdf_sample_pdf > 0.0f)
    {
        if (hit_found && !is_black(light_hit.emission) && cosine_term_at_light_source(light_hit.geometric_normal, -sampled_bsdf_direction) > 0.0f)
        {
            float cosine_at_evaluated_point = fabsf(dot(hit.shading_normal, sampled_bsdf_direction));
            Rgb light_contribution = bsdf_color * light_hit.emission * cosine_at_evaluated_point;
            target_function = luminance(light_contribution);

            float light_pdf = pdf_of_emissive_triangle_hit_solid_angle<V>(rd, light_hit, sampled_bsdf_direction);
            bool contributes_enough = bsdf_sample_pdf <= 0.0f || check_minimum_light_contribution(rd.render_settings.minimum_light_contribution, light_contribution / light_pdf / bsdf_sample_pdf);
            if (!contributes_enough)
                target_function = 0.0f;

            float mis_weight = balance_heuristic(bsdf_sample_pdf, (float)nb_bsdf_candidates, light_pdf, (float)nb_light_candidates);
            candidate_weight = mis_weight * target_function / bsdf_sample_pdf;

            bsdf_RIS_sample.emission = light_hit.emission;
            bsdf_RIS_sample.point_on_light_source = hit.inter_point + sampled_bsdf_direction * light_hit.distance;
            bsdf_RIS_sample.is_bsdf_sample = true;
            bsdf_RIS_sample.bsdf_sample_contribution = bsdf_color;
            bsdf_RIS_sample.bsdf_sample_cosine_term = cosine_at_evaluated_point;
            bsdf_RIS_sample.target_function = target_function;
        }
    }
    reservoir.add_one_candidate(bsdf_RIS_sample, candidate_weight, rng);
}

/* One whole BSDF candidate: sample, trace, finish. */
template <class V>
B2_HD void ris_bsdf_candidate(const Ctx& ctx, Payload& payload, const HitInfo& hit, V3 view_direction, Rng& rng, int nb_light_candidates, int nb_bsdf_candidates,
                              RisReservoir& reservoir)
{
    const b2pt_render_data& rd = *ctx.rd;
    float bsdf_sample_pdf = 0.0f;
    V3 sampled_bsdf_direction;
    int incident_light_info = B2_LOBE_NO_INFO;
    BsdfCtx bctx = make_bsdf_ctx(view_direction, hit.shading_normal, hit.geometric_normal, v3(0.0f), &incident_light_info, &payload.volume_state, false, &payload.material,
                                 payload.bounce, payload.accumulated_roughness, B2_REGULARIZATION_MIS);
    Rgb bsdf_color = bsdf_sample<V>(rd, bctx, sampled_bsdf_direction, bsdf_sample_pdf, rng);
    LightRayHit light_hit;
    bool hit_found = false;
    if (bsdf_sample_pdf > 0.0f)
        hit_found = bsdf_light_sample_ray(ctx, hit.inter_point, sampled_bsdf_direction, 1.0e35f, light_hit, hit.primitive_index);
    ris_bsdf_candidate_finish<V>(rd, hit, rng, nb_light_candidates, nb_bsdf_candidates, bsdf_color, sampled_bsdf_direction, bsdf_sample_pdf, hit_found, light_hit, reservoir);
}

template <class V>
B2_HD RisReservoir sample_bsdf_and_lights_RIS_reservoir(const Ctx& ctx, Payload& payload, const HitInfo& hit, V3 view_direction, Rng& rng)
{
    const b2pt_render_data& rd = *ctx.rd;
    const bool low_res = do_render_low_resolution(rd.render_settings);
    int nb_light_candidates = low_res ? 1 : rd.render_settings.ris_settings.number_of_light_candidates;
    int nb_bsdf_candidates = low_res ? 1 : rd.render_settings.ris_settings.number_of_bsdf_candidates;
    if (!can_do_light_sampling(payload.material))
        nb_light_candidates = 0;

    RisReservoir reservoir;
    reservoir.init();
    for (int i = 0; i < nb_light_candidates; i++)
        ris_light_candidate<V>(ctx, payload, hit, view_direction, rng, nb_light_candidates, nb_bsdf_candidates, reservoir);
    for (int i = 0; i < nb_bsdf_candidates; i++)
        ris_bsdf_candidate<V>(ctx, payload, hit, view_direction, rng, nb_light_candidates, nb_bsdf_candidates, reservoir);
    reservoir.end();
    return reservoir;
}

/* The survivor's shading once its visibility is known (RIS.h:38-81); `direction` = normalised direction to the sample. */
template <class V>
B2_HD Rgb ris_shade_survivor(const Ctx& ctx, Payload& payload, const HitInfo& hit, V3 view_direction, const RisReservoir& reservoir, V3 shadow_ray_direction_normalized, bool in_shadow)
{
    const b2pt_render_data& rd = *ctx.rd;
    const RisSample& sample = reservoir.sample;
    Rgb final_color = rgb(0.0f);
    if (!in_shadow)
    {
        float bsdf_pdf_value;
        float cosine_at_evaluated_point;
        Rgb bsdf_color;
        if (sample.is_bsdf_sample)
        {
            bsdf_color = sample.bsdf_sample_contribution;
            cosine_at_evaluated_point = sample.bsdf_sample_cosine_term;
        }
        else
        {
            /* RIS.h:56-63 evaluates the BSDF of the surviving light sample here: same surface, same direction (both normalise
             * point_on_light - inter_point the same way), same lobe info / bounce / roughness / regularisation as the evaluation the
             * candidate went through when it was weighed (ris_light_candidate), which is a pure function of those: its value was kept */
            (void)bsdf_pdf_value;
            bsdf_color = sample.bsdf_sample_contribution;
            cosine_at_evaluated_point = fabsf(dot(hit.shading_normal, shadow_ray_direction_normalized));
        }
        final_color = bsdf_color * reservoir.UCW * sample.emission * cosine_at_evaluated_point;
        if (!sample.is_bsdf_sample)
            final_color /= 1.0f; /* NEE++ unoccluded probability (NEE++ off) */
    }
    return final_color;
}

template <class V>
B2_HD Rgb evaluate_ris_reservoir_sample(const Ctx& ctx, Payload& payload, const HitInfo& hit, V3 view_direction, const RisReservoir& reservoir)
{
    if (reservoir.UCW <= 0.0f)
        return rgb(0.0f);

    const RisSample& sample = reservoir.sample;
    V3 evaluated_point = hit.inter_point;
    V3 shadow_ray_direction = sample.point_on_light_source - evaluated_point;
    float distance_to_light = length(shadow_ray_direction);
    V3 shadow_ray_direction_normalized = shadow_ray_direction / distance_to_light;

    bool in_shadow;
    if (sample.is_bsdf_sample)
        in_shadow = false;
    else
        in_shadow = shadow_ray_occluded(ctx, evaluated_point, shadow_ray_direction_normalized, distance_to_light, hit.primitive_index);
    return ris_shade_survivor<V>(ctx, payload, hit, view_direction, reservoir, shadow_ray_direction_normalized, in_shadow);
}

template <class V>
B2_HD Rgb sample_lights_RIS(const Ctx& ctx, Payload& payload, const HitInfo& hit, V3 view_direction, Rng& rng)
{
    if (ctx.rd->buffers.emissive_triangles_count == 0)
        return rgb(0.0f);
    RisReservoir reservoir = sample_bsdf_and_lights_RIS_reservoir<V>(ctx, payload, hit, view_direction, rng);
    return evaluate_ris_reservoir_sample<V>(ctx, payload, hit, view_direction, reservoir);
}

#endif
