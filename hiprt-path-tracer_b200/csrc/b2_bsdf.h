/*
 * b2_bsdf.h -- BSDF evaluation context, the Lambertian BRDF and the compile-time BSDF dispatcher.
 * Restates /root/reference/src/Device/includes/BSDFs/BSDFContext.h:13-53, BSDFs/Lambertian.h:14-59 and
 * Dispatcher.h:14-82. The layered principled BSDF lives in b2_principled.h.
 */
#ifndef B2_BSDF_H
#define B2_BSDF_H

#include "b2_intersect.h"
#include "b2_sampling.h"

/* MicrofacetRegularization::RegularizationMode (BSDFs/MicrofacetRegularization.h:15-20) */
enum { B2_REGULARIZATION_NONE = 0, B2_REGULARIZATION_CLASSIC = 1, B2_REGULARIZATION_MIS = 2 };

struct BsdfCtx
{
    const Mat* material;
    VolumeState* volume_state;
    V3 view_direction;
    V3 shading_normal;
    V3 geometric_normal;
    V3 to_light_direction;
    int* incident_light_info;
    int current_bounce;
    float accumulated_path_roughness;
    bool update_ray_volume_state;
    int regularization_mode;
};

B2_HD BsdfCtx make_bsdf_ctx(V3 view_direction, V3 shading_normal, V3 geometric_normal, V3 to_light_direction, int* incident_light_info,
                            VolumeState* volume_state, bool update_ray_volume_state, const Mat* material, int current_bounce,
                            float accumulated_path_roughness, int regularization_mode = B2_REGULARIZATION_NONE)
{
    BsdfCtx c;
    c.material = material;
    c.volume_state = volume_state;
    c.view_direction = view_direction;
    c.shading_normal = shading_normal;
    c.geometric_normal = geometric_normal;
    c.to_light_direction = to_light_direction;
    c.incident_light_info = incident_light_info;
    c.current_bounce = current_bounce;
    c.accumulated_path_roughness = accumulated_path_roughness;
    c.update_ray_volume_state = update_ray_volume_state;
    c.regularization_mode = regularization_mode;
    return c;
}

B2_HD Rgb lambertian_eval(const Mat& material, float NoL, float& pdf)
{
    pdf = 0.0f;
    if (NoL <= 0.0f)
        return rgb(0.0f);
    pdf = NoL * B2_INV_PI;
    return material.base_color * B2_INV_PI;
}

B2_HD float lambertian_pdf(float NoL)
{
    if (NoL <= 0.0f)
        return 0.0f;
    return NoL * B2_INV_PI;
}

B2_HD Rgb lambertian_sample(const Mat& material, V3 shading_normal, V3& sampled_direction, float& pdf, Rng& rng, int& out_sampled_lobe)
{
    sampled_direction = cosine_weighted_sample_around_normal_world_space(shading_normal, rng);
    out_sampled_lobe = B2_LOBE_DIFFUSE;
    return lambertian_eval(material, dot(shading_normal, sampled_direction), pdf);
}

/* defined in b2_principled.h */
B2_HD_NOINLINE Rgb principled_bsdf_eval(const b2pt_render_data& rd, BsdfCtx& ctx, float& pdf);
B2_HD_NOINLINE float principled_bsdf_pdf(const b2pt_render_data& rd, BsdfCtx& ctx);
B2_HD_NOINLINE Rgb principled_bsdf_sample(const b2pt_render_data& rd, BsdfCtx& ctx, V3& sampled_direction, float& pdf, Rng& rng);

template <class V>
B2_HD Rgb bsdf_eval(const b2pt_render_data& rd, BsdfCtx& ctx, float& pdf)
{
    if constexpr (V::BSDF_OVERRIDE == B2_BSDF_LAMBERTIAN)
        return lambertian_eval(*ctx.material, dot(ctx.to_light_direction, ctx.shading_normal), pdf);
    else
        return principled_bsdf_eval(rd, ctx, pdf);
}

template <class V>
B2_HD float bsdf_pdf(const b2pt_render_data& rd, BsdfCtx& ctx)
{
    if constexpr (V::BSDF_OVERRIDE == B2_BSDF_LAMBERTIAN)
        return lambertian_pdf(dot(ctx.to_light_direction, ctx.shading_normal));
    else
        return principled_bsdf_pdf(rd, ctx);
}

template <class V>
B2_HD Rgb bsdf_sample(const b2pt_render_data& rd, BsdfCtx& ctx, V3& sampled_direction, float& pdf, Rng& rng)
{
    if constexpr (V::BSDF_OVERRIDE == B2_BSDF_LAMBERTIAN)
        return lambertian_sample(*ctx.material, ctx.shading_normal, sampled_direction, pdf, rng, *ctx.incident_light_info);
    else
        return principled_bsdf_sample(rd, ctx, sampled_direction, pdf, rng);
}

#endif
