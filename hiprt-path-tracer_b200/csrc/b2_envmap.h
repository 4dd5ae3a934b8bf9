/*
 * b2_envmap.h -- lat-long environment map: lookup, alias-table importance sampling, envmap/BSDF MIS pair.
 * Restates /root/reference/src/Device/includes/LightSampling/Envmap.h:29-38,77-133,139-168,170-249,251-274 and
 * the nearest-texel fetch of Texture.h:209-241 from RGBE 9-9-9-5 texels (Packing.h:441-485).
 */
#ifndef B2_ENVMAP_H
#define B2_ENVMAP_H

#include "b2_lights.h"

B2_HD Rgb unpack_rgbe9995(uint32_t packed)
{
    float r = (float)(packed & 0x1FFu), g = (float)((packed >> 9) & 0x1FFu), b = (float)((packed >> 18) & 0x1FFu);
    int e = (int)(packed >> 27) - 24;
    return rgb(ldexpf(r, e), ldexpf(g, e), ldexpf(b, e));
}

B2_HD Rgb envmap_texel(const b2pt_world_settings& ws, V2 uv)
{
    int x = (int)(uv.x * (float)(ws.envmap_width - 1u));
    int y = (int)(uv.y * (float)(ws.envmap_height - 1u));
    int index = x + y * (int)ws.envmap_width;
    return unpack_rgbe9995(ws.envmap[index]) * ws.envmap_intensity;
}

B2_HD Rgb eval_envmap_no_pdf(const b2pt_world_settings& ws, V3 direction)
{
    V3 rotated = mat3_mul_vec(ws.world_to_envmap_matrix, direction);
    float u = (float)((b2w)0.5f + b2w_atan2(rotated.z, rotated.x) * (b2w)B2_INV_2_PI);
    float v = (float)((b2w)0.5f + b2w_asin(rotated.y) * (b2w)B2_INV_PI);
    return envmap_texel(ws, v2(u, v));
}

template <class V>
B2_HD Rgb envmap_sample(const b2pt_world_settings& ws, V3& sampled_direction, float& envmap_pdf, Rng& rng)
{
    if constexpr (V::ENVMAP == B2_ESS_NONE)
    {
        envmap_pdf = 0.0f;
        return rgb(0.0f);
    }
    float total_sum = ws.envmap_total_sum;
    int random_index = alias_table_sample(ws.envmap_alias_table, rng);
    int y = (int)floorf((float)random_index / (float)ws.envmap_width);
    int x = (int)floorf((float)random_index - (float)y * (float)ws.envmap_width);

    float u = (float)x / (float)ws.envmap_width;
    float v = (float)y / (float)ws.envmap_height;
    float phi = u * B2_TWO_PI;
    float theta = b2_max(1.0e-5f, v * B2_PI);
    b2w sin_theta_w, cos_theta_w, sin_phi_w, cos_phi_w;
    b2w_sincos(theta, sin_theta_w, cos_theta_w);
    b2w_sincos(phi, sin_phi_w, cos_phi_w);
    float cos_theta = (float)cos_theta_w;
    float sin_theta = (float)sin_theta_w;
    sampled_direction = v3((float)((b2w)(-sin_theta) * cos_phi_w), -cos_theta, (float)((b2w)(-sin_theta) * sin_phi_w));
    sampled_direction = mat3_mul_vec(ws.envmap_to_world_matrix, sampled_direction);

    Rgb radiance = envmap_texel(ws, v2(u, v));
    envmap_pdf = luminance(radiance) / (total_sum * ws.envmap_intensity);
    envmap_pdf *= (float)(ws.envmap_width * ws.envmap_height);
    envmap_pdf /= (B2_TWO_PI_SQUARED * sin_theta);
    return radiance;
}

template <class V>
B2_HD Rgb envmap_eval(const b2pt_world_settings& ws, V3 direction, float& pdf)
{
    if constexpr (V::ENVMAP == B2_ESS_NONE)
    {
        pdf = 0.0f;
        return rgb(0.0f);
    }
    Rgb radiance = eval_envmap_no_pdf(ws, direction);
    float theta_bsdf_dir = (float)b2w_acos(-direction.y);
    float sin_theta = (float)b2w_sin(theta_bsdf_dir);
    pdf = luminance(radiance) / (ws.envmap_total_sum * ws.envmap_intensity);
    pdf *= (float)(ws.envmap_width * ws.envmap_height);
    pdf /= (B2_TWO_PI_SQUARED * sin_theta);
    return radiance;
}

/* Envmap.h:170-249 (EnvmapSamplingDoBSDFMIS = TRUE) */
template <class V>
B2_HD Rgb sample_environment_map_with_mis(const Ctx& ctx, Payload& payload, const HitInfo& hit, V3 view_direction, Rng& rng)
{
    const b2pt_render_data& rd = *ctx.rd;
    float envmap_pdf;
    V3 sampled_direction;
    Rgb envmap_color = envmap_sample<V>(rd.world_settings, sampled_direction, envmap_pdf, rng);
    Rgb envmap_mis_contribution = rgb(0.0f);

    if (can_do_light_sampling(payload.material))
    {
        float cosine_term = dot(hit.shading_normal, sampled_direction);
        if (envmap_pdf > 0.0f && cosine_term > 0.0f)
        {
            bool in_shadow = shadow_ray_occluded(ctx, hit.inter_point, sampled_direction, 1.0e35f, hit.primitive_index);
            if (!in_shadow)
            {
                float bsdf_pdf_value;
                int incident_light_info = B2_LOBE_NO_INFO;
                BsdfCtx bctx = make_bsdf_ctx(view_direction, hit.shading_normal, hit.geometric_normal, sampled_direction, &incident_light_info, &payload.volume_state,
                                             false, &payload.material, payload.bounce, payload.accumulated_roughness, B2_REGULARIZATION_MIS);
                Rgb bsdf_color = bsdf_eval<V>(rd, bctx, bsdf_pdf_value);
                float mis_weight = balance_heuristic(envmap_pdf, bsdf_pdf_value);
                envmap_mis_contribution = bsdf_color * cosine_term * mis_weight * envmap_color / envmap_pdf / 1.0f;
            }
        }
    }

    float bsdf_sample_pdf;
    V3 bsdf_sampled_dir;
    Rgb bsdf_mis_contribution = rgb(0.0f);
    int incident_light_info = B2_LOBE_NO_INFO;
    BsdfCtx bctx = make_bsdf_ctx(view_direction, hit.shading_normal, hit.geometric_normal, v3(0.0f), &incident_light_info, &payload.volume_state, false,
                                 &payload.material, payload.bounce, payload.accumulated_roughness, B2_REGULARIZATION_MIS);
    Rgb bsdf_color = bsdf_sample<V>(rd, bctx, bsdf_sampled_dir, bsdf_sample_pdf, rng);
    float cosine_term = fabsf(dot(hit.shading_normal, bsdf_sampled_dir));
    if (bsdf_sample_pdf > 0.0f)
    {
        bool in_shadow = shadow_ray_occluded(ctx, hit.inter_point, bsdf_sampled_dir, 1.0e35f, hit.primitive_index);
        if (!in_shadow)
        {
            float envmap_eval_pdf;
            Rgb envmap_radiance = envmap_eval<V>(rd.world_settings, bsdf_sampled_dir, envmap_eval_pdf);
            if (envmap_eval_pdf > 0.0f)
            {
                float mis_weight = balance_heuristic(bsdf_sample_pdf, envmap_eval_pdf);
                bsdf_mis_contribution = envmap_radiance * mis_weight * cosine_term * bsdf_color / bsdf_sample_pdf;
            }
        }
    }
    return bsdf_mis_contribution + envmap_mis_contribution;
}

/* Envmap.h:251-274 */
template <class V>
B2_HD Rgb sample_environment_map(const Ctx& ctx, Payload& payload, const HitInfo& hit, V3 view_direction, Rng& rng)
{
    const b2pt_world_settings& ws = ctx.rd->world_settings;
    if (ws.ambient_light_type != 2 || ctx.rd->bsdfs_data.white_furnace_mode)
        return rgb(0.0f);
    if (ws.envmap_intensity <= 0.0f)
        return rgb(0.0f);
    if (payload.bounce == 0 && V::DIRECT == B2_LSS_RESTIR_DI)
        return rgb(0.0f);
    if constexpr (V::ENVMAP == B2_ESS_NONE)
        return rgb(0.0f);
    else
        return sample_environment_map_with_mis<V>(ctx, payload, hit, view_direction, rng);
}

#endif
