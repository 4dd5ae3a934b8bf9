/*
 * b2_nee.h -- next-event estimation cut at its rays, for the phased MegaKernel of b2_kernels.cu.
 *
 * The reference's NEE strategies trace their rays in the middle of the shading code (sample_one_light_no_MIS /
 * sample_one_light_MIS, /root/reference/src/Device/includes/LightSampling/Lights.h:25-226; sample_lights_RIS,
 * RIS/RIS.h:18-262; evaluate_ReSTIR_DI_reservoir, ReSTIR/DI/FinalShading.h:16-101). Here each strategy is the same
 * arithmetic and the same RNG draws in the same order, reorganised around two kinds of rays:
 *   - a SEQUENTIAL ray: its result decides what the strategy does next (RIS: the closest hit along BSDF candidate 0 is a
 *     reservoir candidate; which sample survives decides where the shadow ray goes). It is traced between `begin` and
 *     `finish_sampling`.
 *   - DEFERRED rays: their results only switch already computed radiance on or off (shadow rays), or are looked up for an
 *     emission that is then weighted by already computed factors (the BSDF-sample ray of MIS). No RNG draw and no later
 *     decision of the path depends on them, so they are traced TOGETHER WITH THE NEXT MAIN RAY and `resolve` runs one trace
 *     later. For that, the BSDF value of a light sample is evaluated before its visibility is known (the reference evaluates
 *     it only when the sample is visible; bsdf_eval is pure, so the visible case computes the same bits and the occluded case
 *     is discarded).
 * The kernel can then trace the rays of a whole thread block together (one converged traversal loop fed from a shared ray
 * pool of up to three rays per thread) instead of every lane walking the BVH alone inside divergent shading code. The
 * un-split functions (b2_lights.h, b2_ris.h, b2_restir_di_shading.h) stay the reference statement of the math: the host
 * port and the other kernel forms use them, and tests/ compare all forms against the same golden fixtures.
 *
 *   strategy            sequential ray                        deferred X                          deferred Y
 *   ONE_LIGHT           -                                     shadow ray to the light sample      -
 *   BSDF                -                                     -                                   closest hit along the BSDF sample
 *   MIS                 -                                     shadow ray to the light sample      closest hit along the BSDF sample
 *   RIS                 closest hit along BSDF candidate 0    shadow ray to the survivor          -
 *   ReSTIR DI, bounce 0 -                                     shadow ray to the reservoir sample  -
 *   (ReSTIR DI at later bounces = its later-bounces strategy, RIS by default)
 * Envmap NEE (b2_envmap.h) and RIS BSDF candidates beyond the first keep tracing in place: rare, not pooled.
 */
#ifndef B2_NEE_H
#define B2_NEE_H

#include "b2_envmap.h"
#include "b2_lights.h"
#include "b2_restir_di_shading.h"
#include "b2_ris.h"

enum { B2_RAY_NONE = 0, B2_RAY_ANY_HIT = 1, B2_RAY_CLOSEST = 2 };

struct RayRequest
{
    V3 origin, direction;
    float t_max; /* as passed to the traversal: shadow rays carry (distance - 1e-4), closest-hit rays 3e38 */
    int skip_prim;
    int kind;
    int main_path; /* 1: the path's main ray (trace_main_path_ray); 0: a next-event ray -- the alpha test tells them apart (b2_bvh.h) */
};

B2_HD void request_none(RayRequest& r) { r.kind = B2_RAY_NONE; }
B2_HD void request_shadow(RayRequest& r, V3 origin, V3 direction, float distance, int skip_prim)
{
    /* shadow_ray_occluded, b2_intersect.h */
    r.origin = origin;
    r.direction = direction;
    r.t_max = distance - 1.0e-4f;
    r.skip_prim = skip_prim;
    r.kind = B2_RAY_ANY_HIT;
    r.main_path = 0;
}
B2_HD void request_closest(RayRequest& r, V3 origin, V3 direction, int skip_prim, bool main_path_ray = false)
{
    r.main_path = main_path_ray ? 1 : 0;
    /* bsdf_light_sample_ray, b2_intersect.h */
    r.origin = origin;
    r.direction = direction;
    r.t_max = 3.0e38f;
    r.skip_prim = skip_prim;
    r.kind = B2_RAY_CLOSEST;
}

/* strategy used at this vertex (Lights.h:228-352) */
template <class V>
B2_HD int nee_strategy_at(int bounce)
{
    if (V::DIRECT != B2_LSS_RESTIR_DI || bounce == 0)
        return V::DIRECT;
    return V::LATER == B2_LATER_ONE_LIGHT ? B2_LSS_ONE_LIGHT : V::LATER == B2_LATER_BSDF ? B2_LSS_BSDF : V::LATER == B2_LATER_MIS ? B2_LSS_MIS : B2_LSS_RIS;
}

template <class V>
struct Nee
{
    static constexpr bool HAS_ONE_LIGHT = V::DIRECT == B2_LSS_ONE_LIGHT || (V::DIRECT == B2_LSS_RESTIR_DI && V::LATER == B2_LATER_ONE_LIGHT);
    static constexpr bool HAS_BSDF = V::DIRECT == B2_LSS_BSDF || (V::DIRECT == B2_LSS_RESTIR_DI && V::LATER == B2_LATER_BSDF);
    static constexpr bool HAS_MIS = V::DIRECT == B2_LSS_MIS || (V::DIRECT == B2_LSS_RESTIR_DI && V::LATER == B2_LATER_MIS);
    static constexpr bool HAS_RIS = V::DIRECT == B2_LSS_RIS || (V::DIRECT == B2_LSS_RESTIR_DI && V::LATER == B2_LATER_RIS);
    static constexpr bool HAS_RESTIR = V::DIRECT == B2_LSS_RESTIR_DI;
    /* compile-time shape of the trace schedule (see the table above) */
    static constexpr bool HAS_SEQUENTIAL_RAY = HAS_RIS;
    static constexpr bool HAS_DEFERRED_Y = HAS_MIS || HAS_BSDF;

    int strategy;
    bool active;              /* false: sample_emissive_geometry's early-outs (no emissives / white furnace) */
    /* carried from finish_sampling to resolve (one trace later) */
    bool have_x;              /* deferred X was requested: light_radiance counts only if it came back unoccluded */
    Rgb light_radiance;       /* radiance of the light-sampled part IF visible */
    Rgb bsdf_color;           /* BSDF-sampled part (MIS / BSDF): weights of the emission the deferred Y ray finds */
    V3 sampled_direction;
    float bsdf_sample_pdf;
    /* RIS only, between begin and finish_sampling */
    RisReservoir reservoir;
    int nb_light_candidates, nb_bsdf_candidates;
    bool have_candidate_ray;

    /* ---- after the main hit: up to the sequential ray ------------------------------------------------------------------ */
    B2_HD void begin(const Ctx& ctx, Payload& payload, const HitInfo& hit, V3 view_direction, Rng& rng, RayRequest& sequential_ray)
    {
        const b2pt_render_data& rd = *ctx.rd;
        request_none(sequential_ray);
        strategy = nee_strategy_at<V>(payload.bounce);
        active = true;
        have_x = false;
        light_radiance = rgb(0.0f);
        bsdf_sample_pdf = 0.0f;
        /* sample_emissive_geometry, Lights.h:313-352 */
        if (rd.buffers.emissive_triangles_count == 0 && !(rd.world_settings.ambient_light_type == 2 && V::DIRECT == B2_LSS_RESTIR_DI))
            active = false;
        if (rd.bsdfs_data.white_furnace_mode && rd.bsdfs_data.white_furnace_mode_turn_off_emissives)
            active = false;
        if (!active)
            return;
        if (HAS_RIS && strategy == B2_LSS_RIS)
            ris_begin(ctx, payload, hit, view_direction, rng, sequential_ray);
    }

    /* ---- result of the sequential ray -> rest of the sampling; ends with the deferred rays -------------------------------------- */
    B2_HD void finish_sampling(const Ctx& ctx, Payload& payload, const HitInfo& hit, V3 view_direction, Rng& rng, int x, int y, const B2Hit& sequential_result,
                               RayRequest& deferred_x, RayRequest& deferred_y)
    {
        const b2pt_render_data& rd = *ctx.rd;
        request_none(deferred_x);
        request_none(deferred_y);
        if (!active)
            return;
        if (HAS_ONE_LIGHT && strategy == B2_LSS_ONE_LIGHT)
        {
            /* sample_one_light_no_MIS, Lights.h:25-87 */
            LightSample light_sample = sample_one_emissive_triangle<V>(rd, rng);
            if (light_sample.area_measure_pdf <= 0.0f)
                return;
            V3 shadow_ray_direction = light_sample.point_on_light - hit.inter_point;
            float distance_to_light = length(shadow_ray_direction);
            V3 shadow_dir = shadow_ray_direction / distance_to_light;
            float dot_light_source = cosine_term_at_light_source(light_sample.light_source_normal, -shadow_dir);
            if (dot_light_source > 0.0f)
            {
                float bsdf_pdf_value;
                int incident_light_info = light_sample.incident_light_info;
                BsdfCtx bctx = make_bsdf_ctx(view_direction, hit.shading_normal, hit.geometric_normal, shadow_dir, &incident_light_info, &payload.volume_state, false,
                                             &payload.material, payload.bounce, payload.accumulated_roughness, B2_REGULARIZATION_CLASSIC);
                Rgb color = bsdf_eval<V>(rd, bctx, bsdf_pdf_value);
                if (bsdf_pdf_value != 0.0f)
                {
                    float light_sample_solid_angle_pdf = area_to_solid_angle_pdf(light_sample.area_measure_pdf, distance_to_light, dot_light_source);
                    if (light_sample_solid_angle_pdf > 0.0f)
                    {
                        float cosine_term = fabsf(dot(hit.shading_normal, shadow_dir));
                        light_radiance = light_sample.emission * cosine_term * color / light_sample_solid_angle_pdf / 1.0f;
                    }
                }
                request_visibility(deferred_x, hit, shadow_dir, distance_to_light);
            }
        }
        else if (HAS_MIS && strategy == B2_LSS_MIS)
        {
            /* sample_one_light_MIS, Lights.h:123-226 */
            LightSample light_sample = sample_one_emissive_triangle<V>(rd, rng);
            if (light_sample.area_measure_pdf > 0.0f)
            {
                V3 shadow_ray_direction = light_sample.point_on_light - hit.inter_point;
                float distance_to_light = length(shadow_ray_direction);
                V3 shadow_dir = shadow_ray_direction / distance_to_light;
                float bsdf_pdf_value;
                int incident_light_info = B2_LOBE_NO_INFO;
                BsdfCtx bctx = make_bsdf_ctx(view_direction, hit.shading_normal, hit.geometric_normal, shadow_dir, &incident_light_info, &payload.volume_state, false,
                                             &payload.material, payload.bounce, payload.accumulated_roughness, B2_REGULARIZATION_MIS);
                Rgb color = bsdf_eval<V>(rd, bctx, bsdf_pdf_value);
                if (bsdf_pdf_value > 0.0f)
                {
                    float cos_theta_at_light_source = cosine_term_at_light_source(light_sample.light_source_normal, -shadow_dir);
                    if (cos_theta_at_light_source > 1.0e-5f)
                    {
                        float light_sample_solid_angle_pdf = area_to_solid_angle_pdf(light_sample.area_measure_pdf, distance_to_light, cos_theta_at_light_source);
                        float mis_weight = balance_heuristic(light_sample_solid_angle_pdf, bsdf_pdf_value);
                        float cosine_term = fabsf(dot(hit.shading_normal, shadow_dir));
                        light_radiance = color * cosine_term * light_sample.emission * mis_weight / light_sample_solid_angle_pdf / 1.0f;
                    }
                }
                request_visibility(deferred_x, hit, shadow_dir, distance_to_light);
            }
            bsdf_part_sample(rd, payload, hit, view_direction, rng, B2_REGULARIZATION_MIS, deferred_y);
        }
        else if (HAS_BSDF && strategy == B2_LSS_BSDF)
        {
            bsdf_part_sample(rd, payload, hit, view_direction, rng, B2_REGULARIZATION_CLASSIC, deferred_y);
        }
        else if (HAS_RIS && strategy == B2_LSS_RIS)
        {
            ris_finish(ctx, payload, hit, view_direction, rng, sequential_result, deferred_x);
        }
        else if (HAS_RESTIR && strategy == B2_LSS_RESTIR_DI)
        {
            /* sample_light_ReSTIR_DI + evaluate_ReSTIR_DI_reservoir, FinalShading.h:16-125 */
            int pixel_index = x + y * rd.render_settings.render_resolution.x;
            void* buffer = rd.render_settings.restir_di_settings.restir_output_reservoirs;
            DiReservoir di_reservoir = di_load(rd, buffer, pixel_index);
            if (di_reservoir.sample.is_envmap_sample() && rd.world_settings.ambient_light_type != 2)
            {
                di_reservoir.UCW = 0.0f;
                reinterpret_cast<uint32_t*>(buffer)[DI_PLANE_UCW * di_plane_stride(rd) + pixel_index] = b2_float_as_uint(0.0f);
            }
            if (di_reservoir.UCW <= 0.0f)
                return;
            float distance_to_light;
            V3 shadow_dir = di_light_sample_direction(rd, di_reservoir.sample, hit.inter_point, distance_to_light);
            light_radiance = restir_di_shade_sample<V>(ctx, payload, hit, view_direction, di_reservoir, shadow_dir, distance_to_light, false);
            if (!(di_reservoir.sample.flags & B2_DI_FLAGS_UNOCCLUDED) && rd.render_settings.restir_di_settings.do_final_shading_visibility)
                request_visibility(deferred_x, hit, shadow_dir, distance_to_light);
        }
    }

    /* ---- one trace later: the strategy's radiance (what sample_one_light_* / sample_lights_RIS / sample_light_ReSTIR_DI return).
     * `shading_normal` = the vertex's shading normal (the caller still holds the vertex's HitInfo at this point). */
    B2_HD Rgb resolve(const b2pt_render_data& rd, V3 shading_normal, const B2Hit& result_x, const B2Hit& result_y) const
    {
        if (!active)
            return rgb(0.0f);
        Rgb radiance = rgb(0.0f);
        bool in_shadow = have_x && result_x.prim >= 0;
        if (!in_shadow)
            radiance = light_radiance;
        if (HAS_MIS && strategy == B2_LSS_MIS)
            return radiance + bsdf_part_radiance(rd, shading_normal, result_y, true);
        if (HAS_BSDF && strategy == B2_LSS_BSDF)
            return bsdf_part_radiance(rd, shading_normal, result_y, false);
        return radiance;
    }

    /* a shadow ray is only worth tracing when something would be switched off by it */
    B2_HD void request_visibility(RayRequest& ray, const HitInfo& hit, V3 shadow_dir, float distance_to_light)
    {
        /* exactly zero only: a NaN or negative value must still be discarded by an occluded ray, as in the reference */
        if (light_radiance.r == 0.0f && light_radiance.g == 0.0f && light_radiance.b == 0.0f)
            return;
        have_x = true;
        request_shadow(ray, hit.inter_point, shadow_dir, distance_to_light, hit.primitive_index);
    }

    /* ---- BSDF-sample half of sample_one_light_bsdf / sample_one_light_MIS ------------------------------------------------ */
    B2_HD void bsdf_part_sample(const b2pt_render_data& rd, Payload& payload, const HitInfo& hit, V3 view_direction, Rng& rng, int regularization, RayRequest& ray)
    {
        int incident_light_info = B2_LOBE_NO_INFO;
        BsdfCtx bctx = make_bsdf_ctx(view_direction, hit.shading_normal, hit.geometric_normal, v3(0.0f), &incident_light_info, &payload.volume_state, false,
                                     &payload.material, payload.bounce, payload.accumulated_roughness, regularization);
        bsdf_color = bsdf_sample<V>(rd, bctx, sampled_direction, bsdf_sample_pdf, rng);
        if (bsdf_sample_pdf > 0.0f)
            request_closest(ray, hit.inter_point, sampled_direction, hit.primitive_index);
    }

    B2_HD Rgb bsdf_part_radiance(const b2pt_render_data& rd, V3 shading_normal, const B2Hit& result, bool with_mis) const
    {
        Rgb radiance = rgb(0.0f);
        if (bsdf_sample_pdf > 0.0f)
        {
            LightRayHit light_hit;
            bool found = light_ray_hit_from(rd, result, 1.0e35f, light_hit);
            if (found && !is_black(light_hit.emission) && cosine_term_at_light_source(light_hit.geometric_normal, -sampled_direction) > 0.0f)
            {
                if (with_mis)
                {
                    float light_pdf_solid_angle = pdf_of_emissive_triangle_hit_solid_angle<V>(rd, light_hit, sampled_direction);
                    float mis_weight = balance_heuristic(bsdf_sample_pdf, light_pdf_solid_angle);
                    float cosine_term = fabsf(dot(shading_normal, sampled_direction));
                    radiance = bsdf_color * cosine_term * light_hit.emission * mis_weight / bsdf_sample_pdf;
                }
                else
                {
                    float cosine_term = fabsf(dot(shading_normal, sampled_direction));
                    radiance = bsdf_color * cosine_term * light_hit.emission / bsdf_sample_pdf;
                }
            }
        }
        return radiance;
    }

    /* ---- RIS (RIS.h:83-262) ------------------------------------------------------------------------------------ */
    /* the light-candidate loop, then the sampling half of BSDF candidate 0 */
    B2_HD void ris_begin(const Ctx& ctx, Payload& payload, const HitInfo& hit, V3 view_direction, Rng& rng, RayRequest& ray)
    {
        const b2pt_render_data& rd = *ctx.rd;
        const bool low_res = do_render_low_resolution(rd.render_settings);
        nb_light_candidates = low_res ? 1 : rd.render_settings.ris_settings.number_of_light_candidates;
        nb_bsdf_candidates = low_res ? 1 : rd.render_settings.ris_settings.number_of_bsdf_candidates;
        if (!can_do_light_sampling(payload.material))
            nb_light_candidates = 0;
        have_candidate_ray = false;
        reservoir.init();
        /* sample_lights_RIS returns black before anything else when there is no emissive triangle: no candidates at all,
         * reservoir.end() then leaves UCW = 0 and nothing is traced or drawn from the RNG */
        if (rd.buffers.emissive_triangles_count == 0)
        {
            nb_light_candidates = 0;
            nb_bsdf_candidates = 0;
            return;
        }

        for (int i = 0; i < nb_light_candidates; i++)
            ris_light_candidate<V>(ctx, payload, hit, view_direction, rng, nb_light_candidates, nb_bsdf_candidates, reservoir);

        if (nb_bsdf_candidates > 0)
        {
            int incident_light_info = B2_LOBE_NO_INFO;
            BsdfCtx bctx = make_bsdf_ctx(view_direction, hit.shading_normal, hit.geometric_normal, v3(0.0f), &incident_light_info, &payload.volume_state, false, &payload.material,
                                         payload.bounce, payload.accumulated_roughness, B2_REGULARIZATION_MIS);
            bsdf_sample_pdf = 0.0f;
            bsdf_color = bsdf_sample<V>(rd, bctx, sampled_direction, bsdf_sample_pdf, rng);
            if (bsdf_sample_pdf > 0.0f)
            {
                have_candidate_ray = true;
                request_closest(ray, hit.inter_point, sampled_direction, hit.primitive_index);
            }
        }
    }

    /* the rest of BSDF candidate 0, BSDF candidates 1.. (traced in place), reservoir.end(), the survivor's shading if visible */
    B2_HD void ris_finish(const Ctx& ctx, Payload& payload, const HitInfo& hit, V3 view_direction, Rng& rng, const B2Hit& candidate_result, RayRequest& deferred_x)
    {
        const b2pt_render_data& rd = *ctx.rd;
        if (nb_bsdf_candidates > 0)
        {
            LightRayHit light_hit;
            bool hit_found = have_candidate_ray && light_ray_hit_from(rd, candidate_result, 1.0e35f, light_hit);
            ris_bsdf_candidate_finish<V>(rd, hit, rng, nb_light_candidates, nb_bsdf_candidates, bsdf_color, sampled_direction, bsdf_sample_pdf, hit_found, light_hit, reservoir);
            for (int i = 1; i < nb_bsdf_candidates; i++)
                ris_bsdf_candidate<V>(ctx, payload, hit, view_direction, rng, nb_light_candidates, nb_bsdf_candidates, reservoir);
        }
        reservoir.end();
        bsdf_sample_pdf = 0.0f; /* not a MIS / BSDF strategy: resolve() looks at light_radiance only */

        /* evaluate_ris_reservoir_sample (RIS.h:18-81) */
        if (reservoir.UCW <= 0.0f)
            return;
        V3 shadow_ray_direction = reservoir.sample.point_on_light_source - hit.inter_point;
        float distance_to_light = length(shadow_ray_direction);
        V3 shadow_dir = shadow_ray_direction / distance_to_light;
        light_radiance = ris_shade_survivor<V>(ctx, payload, hit, view_direction, reservoir, shadow_dir, false);
        if (!reservoir.sample.is_bsdf_sample)
            request_visibility(deferred_x, hit, shadow_dir, distance_to_light);
    }
};

/* ---- envmap NEE cut at its rays (sample_environment_map, b2_envmap.h / Envmap.h:170-274) -------------------------------------
 * The reference traces two shadow rays in the middle of the envmap MIS estimator: one towards the envmap sample, one along a
 * BSDF sample; each only decides whether an already determined contribution counts. Here both contributions are computed as if
 * visible (bsdf_eval is pure; the reference skips it when the sample is occluded, and the occluded value is discarded here),
 * the two rays are handed back as deferred any-hit rays that ride with the next main ray in the block's ray pool, and
 * EnvmapNee::resolve adds what turned out visible one trace later -- same arithmetic, same RNG draws in the same order.
 * (In round 1 these two rays were traced in place by every lane on its own: on the 10M-triangle ReSTIR DI + envmap workload
 * that was the larger part of the MegaKernel, profiles/r02_ncu_c4s.md.) */
struct EnvmapNee
{
    Rgb light_part, bsdf_part; /* envmap_mis_contribution / bsdf_mis_contribution IF their rays come back unoccluded */
    bool have_light_ray, have_bsdf_ray;

    B2_HD void reset()
    {
        light_part = bsdf_part = rgb(0.0f);
        have_light_ray = have_bsdf_ray = false;
    }

    template <class V>
    B2_HD void sample(const Ctx& ctx, Payload& payload, const HitInfo& hit, V3 view_direction, Rng& rng, RayRequest& light_ray, RayRequest& bsdf_ray)
    {
        reset();
        request_none(light_ray);
        request_none(bsdf_ray);
        const b2pt_render_data& rd = *ctx.rd;
        const b2pt_world_settings& ws = rd.world_settings;
        /* sample_environment_map's early-outs */
        if (ws.ambient_light_type != 2 || rd.bsdfs_data.white_furnace_mode)
            return;
        if (ws.envmap_intensity <= 0.0f)
            return;
        if (payload.bounce == 0 && V::DIRECT == B2_LSS_RESTIR_DI)
            return;
        if constexpr (V::ENVMAP != B2_ESS_NONE)
        {
            /* sample_environment_map_with_mis */
            float envmap_pdf;
            V3 sampled_direction;
            Rgb envmap_color = envmap_sample<V>(ws, sampled_direction, envmap_pdf, rng);
            if (can_do_light_sampling(payload.material))
            {
                float cosine_term = dot(hit.shading_normal, sampled_direction);
                if (envmap_pdf > 0.0f && cosine_term > 0.0f)
                {
                    float bsdf_pdf_value;
                    int incident_light_info = B2_LOBE_NO_INFO;
                    BsdfCtx bctx = make_bsdf_ctx(view_direction, hit.shading_normal, hit.geometric_normal, sampled_direction, &incident_light_info, &payload.volume_state,
                                                 false, &payload.material, payload.bounce, payload.accumulated_roughness, B2_REGULARIZATION_MIS);
                    Rgb bsdf_color = bsdf_eval<V>(rd, bctx, bsdf_pdf_value);
                    float mis_weight = balance_heuristic(envmap_pdf, bsdf_pdf_value);
                    light_part = bsdf_color * cosine_term * mis_weight * envmap_color / envmap_pdf / 1.0f;
                    /* exactly zero only: anything else (NaN included) must still be switched off by an occluded ray */
                    if (!(light_part.r == 0.0f && light_part.g == 0.0f && light_part.b == 0.0f))
                    {
                        have_light_ray = true;
                        request_shadow(light_ray, hit.inter_point, sampled_direction, 1.0e35f, hit.primitive_index);
                    }
                }
            }

            float bsdf_sample_pdf;
            V3 bsdf_sampled_dir;
            int incident_light_info = B2_LOBE_NO_INFO;
            BsdfCtx bctx = make_bsdf_ctx(view_direction, hit.shading_normal, hit.geometric_normal, v3(0.0f), &incident_light_info, &payload.volume_state, false,
                                         &payload.material, payload.bounce, payload.accumulated_roughness, B2_REGULARIZATION_MIS);
            Rgb bsdf_color = bsdf_sample<V>(rd, bctx, bsdf_sampled_dir, bsdf_sample_pdf, rng);
            float cosine_term = fabsf(dot(hit.shading_normal, bsdf_sampled_dir));
            if (bsdf_sample_pdf > 0.0f)
            {
                float envmap_eval_pdf;
                Rgb envmap_radiance = envmap_eval<V>(ws, bsdf_sampled_dir, envmap_eval_pdf);
                if (envmap_eval_pdf > 0.0f)
                {
                    float mis_weight = balance_heuristic(bsdf_sample_pdf, envmap_eval_pdf);
                    bsdf_part = envmap_radiance * mis_weight * cosine_term * bsdf_color / bsdf_sample_pdf;
                    if (!(bsdf_part.r == 0.0f && bsdf_part.g == 0.0f && bsdf_part.b == 0.0f))
                    {
                        have_bsdf_ray = true;
                        request_shadow(bsdf_ray, hit.inter_point, bsdf_sampled_dir, 1.0e35f, hit.primitive_index);
                    }
                }
            }
        }
    }

    /* sample_environment_map's value: bsdf_mis_contribution + envmap_mis_contribution */
    B2_HD Rgb resolve(const B2Hit& light_result, const B2Hit& bsdf_result) const
    {
        Rgb envmap_mis_contribution = (have_light_ray && light_result.prim >= 0) ? rgb(0.0f) : light_part;
        Rgb bsdf_mis_contribution = (have_bsdf_ray && bsdf_result.prim >= 0) ? rgb(0.0f) : bsdf_part;
        return bsdf_mis_contribution + envmap_mis_contribution;
    }
};

#endif
