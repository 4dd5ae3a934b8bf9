/*
 * b2_principled.h -- the layered principled BSDF of the hot path: coat / sheen / two metal lobes / glass / glossy
 * base (specular GGX over Lambertian) / diffuse transmission, with GGX multiple-scattering energy compensation
 * (LUT-driven), microfacet regularization, thin-film interference Fresnel and the sheen LTC lobe.
 *
 * Restates, with the reference's operation order (parity with its host build depends on lobe-selection CDFs and
 * sampled directions rounding identically), the default-macro configuration of
 *   /root/reference/src/Device/includes/BSDFs/Principled.h          :49-1939 (lobes, layering, weights, sample/eval/pdf)
 *   BSDFs/Microfacet.h                                               :25-500  (GGX D / Lambda / G, VNDF sampling, reflect / refract)
 *   BSDFs/MicrofacetEnergyCompensation.h                             :27-64, 101-703 (conductor + glass compensation)
 *   BSDFs/PrincipledEnergyCompensation.h                             :15-129  (glossy base + clearcoat compensation)
 *   BSDFs/MicrofacetRegularization.h                                 :22-161
 *   BSDFs/ThinFilm.h                                                 :13-228
 *   BSDFs/SheenLTC.h                                                 :24-174
 *   Fresnel.h                                                        :12-172
 *   HostDeviceCommon/Material/MaterialUtils.h                        :16-178
 *   Dispersion.h                                                     :485-499 (dispersive scenes are refused at upload, so the IOR is the base IOR)
 * LUT fetches are NEAREST texel with the CPU reference's addressing (src/Image/Image.cpp:564-600, 761-803): that is
 * what the oracle computes, so that is what parity needs (the reference's GPU build uses point / manual-bilinear
 * texture fetches, Texture.h:143-207).
 * Compile-time options fixed at the reference defaults: Lambertian diffuse lobe, GGX VNDF sampling, all energy
 * compensations ON, delta-distribution evaluation optimisation ON, microfacet regularization ON with consistent
 * parameterisation and diffusion heuristic, no Fresnel-based lobe-probability adjustment.
 */
#ifndef B2_PRINCIPLED_H
#define B2_PRINCIPLED_H

#include "b2_bsdf.h"
#include "b2_glass_exponent_table.h"

#define B2_GGX_DOT_PRODUCTS_CLAMP 1.0e-3f

/* MaterialUtils::SpecularDeltaReflectionSampled */
enum { B2_NOT_SPECULAR = -1, B2_SPECULAR_PEAK_NOT_SAMPLED = 0, B2_SPECULAR_PEAK_SAMPLED = 1 };

/* ------------------------------------------------------------------------------------------------------------
 * LUT fetches (nearest texel, repeat addressing, bottom-left origin)
 * ------------------------------------------------------------------------------------------------------------ */
B2_HD float lut_wrap(float c)
{
    float r = c;
    if (r != 1.0f)
        r -= (float)(int)c;
    return r < 0.0f ? 1.0f + r : r;
}

/* Image32Bit::sample_rgba32f: returns channel `ch` texels of a 2D table */
B2_HD V3 lut_fetch_2d(const void* handle, float u_in, float v_in)
{
    const b2pt_lut* lut = reinterpret_cast<const b2pt_lut*>(handle);
    if (lut == nullptr)
        return v3(0.0f);
    float u = lut_wrap(u_in);
    float v = lut_wrap(v_in);
    v = 1.0f - v;
    int x = (int)(u * (float)(lut->width - 1));
    int y = (int)(v * (float)(lut->height - 1));
    /* a NaN or astronomically large coordinate (the reference then reads out of bounds, and crashes on its CPU path) lands on an edge texel */
    x = x < 0 ? 0 : (x > lut->width - 1 ? lut->width - 1 : x);
    y = y < 0 ? 0 : (y > lut->height - 1 ? lut->height - 1 : y);
    const float* t = lut->texels + (size_t)(x + y * lut->width) * lut->channels;
    V3 r = v3(t[0], 0.0f, 0.0f);
    if (lut->channels >= 3)
    {
        r.y = t[1];
        r.z = t[2];
    }
    return r;
}

/* Image32Bit3D::sample_rgba32f, first channel */
B2_HD float lut_fetch_3d(const void* handle, float u_in, float v_in, float w_in)
{
    const b2pt_lut* lut = reinterpret_cast<const b2pt_lut*>(handle);
    if (lut == nullptr)
        return 0.0f;
    float u = lut_wrap(u_in);
    float v = lut_wrap(v_in);
    float w = lut_wrap(w_in);
    v = 1.0f - v;
    int x = (int)(u * (float)(lut->width - 1));
    int y = (int)(v * (float)(lut->height - 1));
    int z = (int)(w * (float)(lut->depth - 1));
    x = x < 0 ? 0 : (x > lut->width - 1 ? lut->width - 1 : x);
    y = y < 0 ? 0 : (y > lut->height - 1 ? lut->height - 1 : y);
    z = z < 0 ? 0 : (z > lut->depth - 1 ? lut->depth - 1 : z);
    return lut->texels[((size_t)z * lut->height * lut->width + (size_t)(x + y * lut->width)) * lut->channels];
}

/* ------------------------------------------------------------------------------------------------------------
 * Fresnel (Fresnel.h)
 * ------------------------------------------------------------------------------------------------------------ */
B2_HD float F0_from_eta(float eta_t, float eta_i)
{
    float nume_F0 = (eta_t - eta_i);
    float denom_F0 = (eta_t + eta_i);
    return (nume_F0 * nume_F0) / (denom_F0 * denom_F0);
}
B2_HD float F0_from_eta_t_and_relative_ior(float eta_t, float relative_eta) { return F0_from_eta(eta_t, eta_t / relative_eta); }

/* relative_eta = eta_t / eta_i */
B2_HD float full_fresnel_dielectric(float cos_theta_i, float relative_eta)
{
    if (fabsf(1.0f - relative_eta) < 1.0e-4f)
        return 0.0f;
    float sin_theta_i2 = 1.0f - cos_theta_i * cos_theta_i;
    float sin_theta_t2 = sin_theta_i2 / (relative_eta * relative_eta);
    if (sin_theta_t2 >= 1.0f)
        return 1.0f;
    float cos_theta_t = sqrtf(1.0f - sin_theta_t2);
    float r_parallel = (relative_eta * cos_theta_i - cos_theta_t) / (relative_eta * cos_theta_i + cos_theta_t);
    float r_perpendicular = (cos_theta_i - relative_eta * cos_theta_t) / (cos_theta_i + relative_eta * cos_theta_t);
    return (r_parallel * r_parallel + r_perpendicular * r_perpendicular) / 2;
}
B2_HD float full_fresnel_dielectric(float cos_theta_i, float eta_i, float eta_t) { return full_fresnel_dielectric(cos_theta_i, eta_t / eta_i); }

/* Hoffman 2023 / Adobe Standard Material F82-tint conductor Fresnel */
B2_HD Rgb adobe_f82_tint_fresnel(Rgb F0, Rgb F82, Rgb F90, float F90_falloff_exponent, float cos_theta)
{
    Rgb base_term = F0 + (F90 - F0) * (float)b2w_pow(1.0f - cos_theta, F90_falloff_exponent);
    if (max_component(base_term) < 1.0e-8f)
        return rgb(0.0f);
    float lazanyi_correction = cos_theta * b2_pow_6(1.0f - cos_theta);
    const float cos_theta_max = 1.0f / 7.0f;
    const float denom_a = cos_theta_max * b2_pow_6(1.0f - cos_theta_max);
    Rgb nume_a = (F0 + (F90 - F0) * (float)b2w_pow(1.0f - cos_theta_max, F90_falloff_exponent)) * (rgb(1.0f) - F82);
    Rgb a = nume_a / denom_a;
    Rgb F = base_term - a * lazanyi_correction;
    return rgb_clamp(F, 0.0f, 1.0f);
}

/* d'Eon, "A Hitchhiker's Guide to Multiple Scattering", eq. 11.15 fit */
B2_HD float fresnel_hemispherical_albedo_fit(float relative_eta)
{
    return b2f_log((10893.0f * relative_eta - 1438.2f) / (-774.4f * b2_square(relative_eta) + 10212.0f * relative_eta + 1.0f));
}

/* ------------------------------------------------------------------------------------------------------------
 * Thin-film interference Fresnel (Belcour & Barla 2017; ThinFilm.h)
 * ------------------------------------------------------------------------------------------------------------ */
B2_HD Rgb thin_film_eval_sensitivity(float opd, float shift)
{
    float phase = 2.0f * B2_PI * opd * 1.0e-6f;
    V3 val = v3(5.4856e-13f, 4.4201e-13f, 5.2481e-13f);
    V3 pos = v3(1.6810e+06f, 1.7953e+06f, 2.2084e+06f);
    V3 var = v3(4.3278e+09f, 9.3046e+09f, 6.6121e+09f);
    V3 sq = v3(sqrtf((2.0f * B2_PI * var).x), sqrtf((2.0f * B2_PI * var).y), sqrtf((2.0f * B2_PI * var).z));
    V3 arg = pos * phase + v3(shift);
    V3 cs = v3(b2f_cos(arg.x), b2f_cos(arg.y), b2f_cos(arg.z));
    V3 ea = -1.0f * var * phase * phase;
    V3 ex = v3(b2f_exp(ea.x), b2f_exp(ea.y), b2f_exp(ea.z));
    V3 xyz = val * sq * cs * ex;
    xyz.x += 9.7470e-14f * sqrtf(2.0f * B2_PI * 4.5282e+09f) * b2f_cos(2.2399e+06f * phase + shift) * b2f_exp(-4.5282e+09f * phase * phase);
    V3 r = xyz / 1.0685e-7f;
    return rgb(r.x, r.y, r.z);
}

B2_HD void thin_film_fresnel_phase(float cos_theta_i, float eta1, float eta2, float kappa2, float& phi_par, float& phi_perp)
{
    float sinThetaSqr = 1.0f - b2_square(cos_theta_i);
    float A = b2_square(eta2) * (1.0f - b2_square(kappa2)) - b2_square(eta1) * sinThetaSqr;
    float B = sqrtf(b2_square(A) + b2_square(2 * b2_square(eta2) * kappa2));
    float U = sqrtf((A + B) * 0.5f);
    float V = sqrtf((B - A) * 0.5f);
    float phi_perp_y = 2.0f * eta1 * V * cos_theta_i;
    float phi_perp_x = b2_square(U) + b2_square(V) - b2_square(eta1 * cos_theta_i);
    phi_perp = b2f_atan2(phi_perp_y, phi_perp_x);
    float phi_par_y = 2.0f * eta1 * b2_square(eta2) * cos_theta_i * (2.0f * kappa2 * U - (1.0f - b2_square(kappa2)) * V);
    float phi_par_x = b2_square(b2_square(eta2) * (1.0f + b2_square(kappa2)) * cos_theta_i) - b2_square(eta1) * (b2_square(U) + b2_square(V));
    phi_par = b2f_atan2(phi_par_y, phi_par_x);
}

B2_HD void thin_film_fresnel_conductor(float cos_theta_i, float eta, float k, float& Rp2, float& Rs2)
{
    float cos_theta_i_2 = cos_theta_i * cos_theta_i;
    float sin_theta_i_2 = 1.0f - cos_theta_i_2;
    float temp1 = eta * eta - k * k - sin_theta_i_2;
    float a2pb2 = sqrtf(temp1 * temp1 + 4.0f * k * k * eta * eta);
    float a = sqrtf(0.5f * (a2pb2 + temp1));
    float term1 = a2pb2 + cos_theta_i_2;
    float term2 = 2.0f * a * cos_theta_i;
    Rs2 = (term1 - term2) / (term1 + term2);
    Rs2 = b2_clamp(0.0f, 1.0f, Rs2);
    float term3 = a2pb2 * cos_theta_i_2 + sin_theta_i_2 * sin_theta_i_2;
    float term4 = term2 * sin_theta_i_2;
    Rp2 = Rs2 * (term3 - term4) / (term3 + term4);
    Rp2 = b2_clamp(0.0f, 1.0f, Rp2);
}

B2_HD Rgb rgb_hue_shift(Rgb color, float hue_shift_degrees)
{
    if (hue_shift_degrees == 0.0f)
        return color;
    float cosA = b2f_cos(hue_shift_degrees / 180.0f * B2_PI);
    float sinA = b2f_sin(hue_shift_degrees / 180.0f * B2_PI);
    const float sqrt_1_3 = 0.57735026918962576451f;
    const float one_over_3 = 1.0f / 3.0f;
    float m00 = cosA + (1.0f - cosA) / 3.0f;
    float m01 = one_over_3 * (1.0f - cosA) - sqrt_1_3 * sinA;
    float m02 = one_over_3 * (1.0f - cosA) + sqrt_1_3 * sinA;
    float m10 = one_over_3 * (1.0f - cosA) + sqrt_1_3 * sinA;
    float m11 = cosA + one_over_3 * (1.0f - cosA);
    float m12 = one_over_3 * (1.0f - cosA) - sqrt_1_3 * sinA;
    float m20 = one_over_3 * (1.0f - cosA) - sqrt_1_3 * sinA;
    float m21 = one_over_3 * (1.0f - cosA) + sqrt_1_3 * sinA;
    float m22 = cosA + one_over_3 * (1.0f - cosA);
    Rgb s;
    s.r = color.r * m00 + color.g * m01 + color.b * m02;
    s.g = color.r * m10 + color.g * m11 + color.b * m12;
    s.b = color.r * m20 + color.g * m21 + color.b * m22;
    return rgb_clamp(s, 0.0f, 1.0f);
}

B2_HD_NOINLINE Rgb thin_film_fresnel(const Mat& material, float ambient_IOR, float HoL)
{
    if (material.thin_film == 0.0f)
        return rgb(0.0f);

    float eta1 = ambient_IOR;
    float eta2 = material.thin_film_ior;
    float eta3 = material.thin_film_do_ior_override ? material.thin_film_base_ior_override : material.ior;
    float kappa3 = material.thin_film_do_ior_override ? material.thin_film_kappa_3 : 0.0f;

    float R12p = 0.0f, R12s = 0.0f, T121p = 0.0f, T121s = 0.0f, R23p = 0.0f, R23s = 0.0f, cos_theta_2 = 0.0f;
    float cos_theta_transmission_2 = 1.0f - (1.0f - b2_square(HoL)) * b2_square(eta1 / eta2);
    if (cos_theta_transmission_2 <= 0.0f)
    {
        R12s = 1.0f;
        R12p = 1.0f;
        T121p = 0.0f;
        T121s = 0.0f;
    }
    else
    {
        cos_theta_2 = sqrtf(cos_theta_transmission_2);
        thin_film_fresnel_conductor(HoL, eta2 / eta1, 0.0f, R12p, R12s);
        thin_film_fresnel_conductor(cos_theta_2, eta3 / eta2, kappa3, R23p, R23s);
        T121p = 1.0f - R12p;
        T121s = 1.0f - R12s;
    }

    float D = material.thin_film_thickness / 1000.0f * cos_theta_2;
    float phi21p, phi21s, phi23p, phi23s;
    thin_film_fresnel_phase(HoL, eta1, eta2, 0.0f, phi21p, phi21s);
    thin_film_fresnel_phase(cos_theta_2, eta2, eta3, kappa3, phi23p, phi23s);
    phi21p = B2_PI - phi21p;
    phi21s = B2_PI - phi21s;

    float r123p = sqrtf(R12p * R23p);
    float r123s = sqrtf(R12s * R23s);

    float Rs = (b2_square(T121p) * R23p) / (1.0f - R12p * R23p);
    float C0 = R12p + Rs;
    Rgb I = rgb(C0);
    Rgb Sm;
    float Cm = Rs - T121p;
    for (int m = 1; m <= 2; ++m)
    {
        Cm *= r123p;
        Sm = 2.0f * thin_film_eval_sensitivity(m * D, m * (phi23p + phi21p));
        I += Cm * Sm;
    }
    Rs = (b2_square(T121s) * R23s) / (1.0f - R12s * R23s);
    C0 = R12s + Rs;
    I += rgb(C0);
    Cm = Rs - T121s;
    for (int m = 1; m <= 2; ++m)
    {
        Cm *= r123s;
        Sm = 2.0f * thin_film_eval_sensitivity(m * D, m * (phi23s + phi21s));
        I += Cm * Sm;
    }
    I *= 0.5f;

    float r = 2.3646381f * I.r - 0.8965361f * I.g - 0.4680737f * I.b;
    float g = -0.5151664f * I.r + 1.4264000f * I.g + 0.0887608f * I.b;
    float b = 0.0052037f * I.r - 0.0144081f * I.g + 1.0092106f * I.b;
    I = rgb_clamp(rgb(r, g, b), 0.0f, 1.0f);
    return rgb_hue_shift(I, material.thin_film_hue_shift_degrees * 360.0f);
}

/* ------------------------------------------------------------------------------------------------------------
 * Microfacet regularization (Jendersie & Grosch 2019; MicrofacetRegularization.h)
 * ------------------------------------------------------------------------------------------------------------ */
B2_HD float regularization_consistent_tau_compute(float tau_0, int sample_number)
{
    /* powf(...) * sqrt(...) / (...): the unqualified sqrt makes the product and the quotient double (MicrofacetRegularization.h:160) */
    return 1.0f / (2.0f * B2_PI * (1.0f - b2f_cos(b2f_atan((float)((b2w)b2f_pow((float)(sample_number + 1), -1.0f / 6.0f) * b2w_sqrt(B2_FOUR_PI * tau_0 - 1.0f) / (b2w)(B2_TWO_PI * tau_0 - 1.0f))))));
}

/* The value depends on two launch-uniform inputs only (tau_0 and the sample number; the call sites ask for sample_number and
 * sample_number + 1), yet the reference evaluates its powf / sqrt / atanf / cosf chain at every regularised lobe: ~300
 * instructions per call and, inlined at ten sites, 12 % of the precise MegaKernel's code. On the device every block computes the two
 * values once (b2_block_constants_init, first thing in every pixel kernel) and the call sites read them from shared memory. */
#if defined(__CUDACC__)
__shared__ float b2_s_consistent_tau[2];
__device__ __forceinline__ void b2_block_constants_init(const b2pt_render_data& rd)
{
    if (threadIdx.x < 2)
        b2_s_consistent_tau[threadIdx.x] = regularization_consistent_tau_compute(rd.bsdfs_data.microfacet_regularization_tau_0, (int)rd.render_settings.sample_number + (int)threadIdx.x);
    __syncthreads();
}
#endif
B2_HD float regularization_consistent_tau(const b2pt_render_data& rd, int sample_offset)
{
#if defined(__CUDA_ARCH__)
    return b2_s_consistent_tau[sample_offset];
#else
    return regularization_consistent_tau_compute(rd.bsdfs_data.microfacet_regularization_tau_0, (int)rd.render_settings.sample_number + sample_offset);
#endif
}

B2_HD float regularize_reflection(const b2pt_render_data& rd, int mode, float initial_roughness, float accumulated_path_roughness)
{
    if (mode == B2_REGULARIZATION_NONE)
        return initial_roughness;
    float consistent_tau = regularization_consistent_tau(rd, 0);
    float final_tau = consistent_tau / b2_max(b2_square(accumulated_path_roughness), 1.0e-8f);
    float regularized_roughness = sqrtf(sqrtf(1.0f / (final_tau * B2_PI)));
    return b2_max(rd.bsdfs_data.microfacet_regularization_min_roughness, b2_max(initial_roughness, regularized_roughness));
}

B2_HD float regularize_refraction(const b2pt_render_data& rd, int mode, float initial_roughness, float accumulated_path_roughness, float eta_i, float eta_t)
{
    if (mode == B2_REGULARIZATION_NONE)
        return initial_roughness;
    /* the reference passes sample_number + 1 here and only here (MicrofacetRegularization.h:70) */
    float consistent_tau = regularization_consistent_tau(rd, 1);
    float final_tau = consistent_tau / b2_max(b2_square(accumulated_path_roughness), 1.0e-8f);
    float regularized_roughness = sqrtf(sqrtf(1.0f / (final_tau * B2_PI * b2_square(eta_i - eta_t) / (4.0f * b2_square(b2_max(eta_i, eta_t))))));
    return b2_max(rd.bsdfs_data.microfacet_regularization_min_roughness, b2_max(initial_roughness, regularized_roughness));
}

B2_HD float regularize_mix_reflection_refraction(const b2pt_render_data& rd, int mode, float initial_roughness, float accumulated_path_roughness, float eta_i, float eta_t)
{
    if (mode == B2_REGULARIZATION_NONE)
        return initial_roughness;
    float consistent_tau = regularization_consistent_tau(rd, 0);
    float final_tau = consistent_tau / b2_max(b2_square(accumulated_path_roughness), 1.0e-8f);
    float regularized_roughness_reflection = sqrtf(sqrtf(1.0f / (final_tau * B2_PI)));
    if (eta_i == eta_t)
        return regularized_roughness_reflection;
    float regularized_roughness_refraction = sqrtf(sqrtf(1.0f / (final_tau * B2_PI * b2_square(eta_i - eta_t) / (4.0f * b2_square(b2_max(eta_i, eta_t))))));
    return b2_max(rd.bsdfs_data.microfacet_regularization_min_roughness,
                  b2_max(initial_roughness, regularized_roughness_refraction * 0.75f + regularized_roughness_reflection * 0.25f));
}

/* ------------------------------------------------------------------------------------------------------------
 * MaterialUtils
 * ------------------------------------------------------------------------------------------------------------ */
B2_HD float thin_walled_roughness(bool thin_walled, float base_roughness, float relative_eta)
{
    if (!thin_walled)
        return base_roughness;
    float remapped = (float)((b2w)base_roughness * b2w_sqrt(3.7f * (relative_eta - 1.0f) * b2_square(relative_eta - 0.5f) / b2_pow_3(relative_eta)));
    return b2_clamp(0.0f, 1.0f, remapped / 1.39f);
}

B2_HD int is_specular_delta_reflection_sampled(const Mat& material, float delta_distribution_roughness, float delta_distribution_anisotropy, int incident_light_info)
{
    if (!material_is_perfectly_smooth(delta_distribution_roughness))
        return B2_NOT_SPECULAR;

    bool matching_base_substrate_anisotropy = fabsf(delta_distribution_anisotropy - material.anisotropy) < 1.0e-3f;
    bool sampled_from_glass = incident_light_info == B2_LOBE_GLASS_REFLECT && material_is_perfectly_smooth(material.roughness) && matching_base_substrate_anisotropy;
    if (sampled_from_glass)
        return B2_SPECULAR_PEAK_SAMPLED;

    bool sampled_from_first_metal = incident_light_info == B2_LOBE_FIRST_METAL && material_is_perfectly_smooth(material.roughness) && matching_base_substrate_anisotropy;
    bool sampled_from_second_metal = incident_light_info == B2_LOBE_SECOND_METAL && material_is_perfectly_smooth(material.second_roughness) && matching_base_substrate_anisotropy;
    if (sampled_from_first_metal || sampled_from_second_metal)
        return B2_SPECULAR_PEAK_SAMPLED;

    bool matching_coat_anisotropy = fabsf(delta_distribution_anisotropy - material.coat_anisotropy) < 1.0e-3f;
    bool sampled_from_coat = incident_light_info == B2_LOBE_COAT && matching_coat_anisotropy && material_is_perfectly_smooth(material.coat_roughness);
    if (sampled_from_coat)
        return B2_SPECULAR_PEAK_SAMPLED;

    bool sampled_from_specular = incident_light_info == B2_LOBE_SPECULAR && material_is_perfectly_smooth(material.roughness) && matching_base_substrate_anisotropy;
    if (sampled_from_specular)
        return B2_SPECULAR_PEAK_SAMPLED;

    return B2_SPECULAR_PEAK_NOT_SAMPLED;
}

/* ------------------------------------------------------------------------------------------------------------
 * GGX microfacet model (Microfacet.h)
 * ------------------------------------------------------------------------------------------------------------ */
B2_HD float GGX_anisotropic(float alpha_x, float alpha_y, V3 m)
{
    float denom = (m.x * m.x) / (alpha_x * alpha_x) + (m.y * m.y) / (alpha_y * alpha_y) + (m.z * m.z);
    return 1.0f / (B2_PI * alpha_x * alpha_y * denom * denom);
}

B2_HD float GGX_anisotropic_vndf(float D, float G1V, V3 local_view_direction, V3 local_microfacet_normal)
{
    float HoL = b2_max(B2_GGX_DOT_PRODUCTS_CLAMP, dot(local_view_direction, local_microfacet_normal));
    return G1V * D * HoL / local_view_direction.z;
}

B2_HD float G1_Smith_lambda(float alpha_x, float alpha_y, V3 d)
{
    float ax = d.x * alpha_x;
    float ay = d.y * alpha_y;
    return (float)(((b2w)-1.0f + b2w_sqrt(1.0f + (ax * ax + ay * ay) / (d.z * d.z))) * (b2w)0.5f);
}

B2_HD float G1_Smith(float alpha_x, float alpha_y, V3 d) { return 1.0f / (1.0f + G1_Smith_lambda(alpha_x, alpha_y, d)); }

/* torrance_sparrow_GGX_eval_reflect<0>: single scattering FDG / (4 NoL NoV), pdf = VNDF / (4 VoH) */
B2_HD Rgb ggx_eval_reflect(const b2pt_render_data& rd, float material_roughness, float material_anisotropy, Rgb F, V3 local_view_direction, V3 local_to_light_direction,
                           V3 local_halfway_vector, float& out_pdf, int delta_sampled)
{
    out_pdf = 0.0f;
    if (material_is_perfectly_smooth(material_roughness))
    {
        if (delta_sampled == B2_SPECULAR_PEAK_NOT_SAMPLED)
            return rgb(0.0f);
        if (dot(reflect_ray(local_view_direction, v3(0.0f, 0.0f, 1.0f)), local_to_light_direction) < B2_DELTA_DISTRIBUTION_ALIGNEMENT_THRESHOLD)
        {
            out_pdf = 0.0f;
            return rgb(0.0f);
        }
        out_pdf = B2_DELTA_DISTRIBUTION_HIGH_VALUE;
        return rgb(B2_DELTA_DISTRIBUTION_HIGH_VALUE) * F / fabsf(local_to_light_direction.z);
    }

    if (local_to_light_direction.z < 0.0f)
        return rgb(0.0f);

    float alpha_x, alpha_y;
    material_alphas(material_roughness, material_anisotropy, alpha_x, alpha_y);

    float D = GGX_anisotropic(alpha_x, alpha_y, local_halfway_vector);
    float lambda_V = G1_Smith_lambda(alpha_x, alpha_y, local_view_direction);
    float G1V = 1.0f / (1.0f + lambda_V);
    float Dvisible = GGX_anisotropic_vndf(D, G1V, local_view_direction, local_halfway_vector);

    float NoV = b2_max(B2_GGX_DOT_PRODUCTS_CLAMP, fabsf(local_view_direction.z));
    float NoL = b2_max(B2_GGX_DOT_PRODUCTS_CLAMP, fabsf(local_to_light_direction.z));

    out_pdf = Dvisible / (4.0f * dot(local_view_direction, local_halfway_vector));
    if (out_pdf == 0.0f)
        return rgb(0.0f);

    float lambda_L = G1_Smith_lambda(alpha_x, alpha_y, local_to_light_direction);
    if (rd.bsdfs_data.GGX_masking_shadowing == 1) /* HeightUncorrelated */
    {
        float G1L = 1.0f / (1.0f + lambda_L);
        float G2 = G1V * G1L;
        return F * D * G2 / (4.0f * NoL * NoV);
    }
    float G2HeightCorrelated = 1.0f / (1.0f + lambda_V + lambda_L);
    return F * D * G2HeightCorrelated / (4.0f * NoL * NoV);
}

B2_HD float ggx_pdf_reflect(float material_roughness, float material_anisotropy, V3 local_view_direction, V3 local_to_light_direction, V3 local_halfway_vector, int delta_sampled)
{
    if (material_is_perfectly_smooth(material_roughness))
    {
        if (delta_sampled == B2_SPECULAR_PEAK_NOT_SAMPLED)
            return 0.0f;
        if (dot(reflect_ray(local_view_direction, v3(0.0f, 0.0f, 1.0f)), local_to_light_direction) < B2_DELTA_DISTRIBUTION_ALIGNEMENT_THRESHOLD)
            return 0.0f;
        return B2_DELTA_DISTRIBUTION_HIGH_VALUE;
    }
    if (local_to_light_direction.z < 0.0f)
        return 0.0f;

    float alpha_x, alpha_y;
    material_alphas(material_roughness, material_anisotropy, alpha_x, alpha_y);
    float D = GGX_anisotropic(alpha_x, alpha_y, local_halfway_vector);
    float lambda_V = G1_Smith_lambda(alpha_x, alpha_y, local_view_direction);
    float G1V = 1.0f / (1.0f + lambda_V);
    float Dvisible = GGX_anisotropic_vndf(D, G1V, local_view_direction, local_halfway_vector);
    return Dvisible / (4.0f * dot(local_view_direction, local_halfway_vector));
}

/* get_GGX_energy_compensation_conductors (Turquin 2019, eq. 9-15) */
B2_HD Rgb ggx_energy_compensation_conductors(const b2pt_render_data& rd, Rgb F0, float material_roughness, bool material_do_energy_compensation, V3 local_view_direction, int current_bounce)
{
    bool max_bounce_reached = current_bounce > rd.bsdfs_data.metal_energy_compensation_max_bounce && rd.bsdfs_data.metal_energy_compensation_max_bounce > -1;
    bool smooth_enough = material_roughness <= rd.bsdfs_data.energy_compensation_roughness_threshold;
    bool invalid_view_direction = local_view_direction.z < 0.0f;
    if (!material_do_energy_compensation || smooth_enough || max_bounce_reached || invalid_view_direction)
        return rgb(1.0f);

    float u = b2_max(0.0f, local_view_direction.z);
    float v = 1.0f - material_roughness;
    float Ess = lut_fetch_2d(rd.bsdfs_data.GGX_conductor_directional_albedo, u, v).x;
    float kms = (1.0f - Ess) / Ess;
    return rgb(1.0f) + kms * F0;
}

/* torrance_sparrow_GGX_eval_reflect<1> */
B2_HD Rgb ggx_eval_reflect_compensated(const b2pt_render_data& rd, float material_roughness, float material_anisotropy, bool material_do_energy_compensation, Rgb F,
                                       V3 local_view_direction, V3 local_to_light_direction, V3 local_halfway_vector, float& out_pdf, int delta_sampled, int current_bounce)
{
    Rgb ms_compensation_term = ggx_energy_compensation_conductors(rd, F, material_roughness, material_do_energy_compensation, local_view_direction, current_bounce);
    Rgb single_scattering = ggx_eval_reflect(rd, material_roughness, material_anisotropy, F, local_view_direction, local_to_light_direction, local_halfway_vector, out_pdf, delta_sampled);
    return single_scattering * ms_compensation_term;
}

B2_HD Rgb ggx_eval_refract(const Mat& material, float roughness, float relative_eta, Rgb fresnel_reflectance, V3 local_view_direction, V3 local_to_light_direction,
                           V3 local_halfway_vector, float& out_pdf, int incident_light_info)
{
    float NoL = local_to_light_direction.z;
    float NoV = local_view_direction.z;
    float HoL = dot(local_to_light_direction, local_halfway_vector);
    float HoV = dot(local_view_direction, local_halfway_vector);

    Rgb color;
    if (material_is_perfectly_smooth(roughness))
    {
        bool incident_direction_is_perfect_refraction = dot(refract_ray(local_view_direction, v3(0.0f, 0.0f, 1.0f), relative_eta), local_to_light_direction) > B2_DELTA_DISTRIBUTION_ALIGNEMENT_THRESHOLD;
        if (incident_light_info == B2_LOBE_GLASS_REFRACT && incident_direction_is_perfect_refraction)
        {
            color = rgb(B2_DELTA_DISTRIBUTION_HIGH_VALUE) * (rgb(1.0f) - fresnel_reflectance) * material.base_color;
            color /= fabsf(NoL);
            out_pdf = B2_DELTA_DISTRIBUTION_HIGH_VALUE;
        }
        else
        {
            color = rgb(0.0f);
            out_pdf = 0.0f;
        }
    }
    else
    {
        float dot_prod = HoL + HoV / relative_eta;
        float dot_prod2 = dot_prod * dot_prod;
        float denom = dot_prod2 * NoL * NoV;

        float alpha_x, alpha_y;
        material_alphas(roughness, material.anisotropy, alpha_x, alpha_y);
        float D = GGX_anisotropic(alpha_x, alpha_y, local_halfway_vector);
        float G1_V = G1_Smith(alpha_x, alpha_y, local_view_direction);
        float G1_L = G1_Smith(alpha_x, alpha_y, local_to_light_direction);
        float G2 = G1_V * G1_L;

        float dwm_dwi = fabsf(HoL) / dot_prod2;
        float D_pdf = G1_V / fabsf(NoV) * D * fabsf(HoV);
        out_pdf = dwm_dwi * D_pdf;
        color = material.base_color * D * (rgb(1.0f) - fresnel_reflectance) * G2 * fabsf(HoL * HoV / denom);
    }
    color /= b2_square(relative_eta);
    return color;
}

B2_HD float ggx_pdf_refract(const Mat& material, float roughness, float relative_eta, V3 local_view_direction, V3 local_to_light_direction, V3 local_halfway_vector, int incident_light_info)
{
    float NoV = local_view_direction.z;
    float HoL = dot(local_to_light_direction, local_halfway_vector);
    float HoV = dot(local_view_direction, local_halfway_vector);
    if (material_is_perfectly_smooth(roughness))
    {
        bool incident_direction_is_perfect_refraction = dot(refract_ray(local_view_direction, v3(0.0f, 0.0f, 1.0f), relative_eta), local_to_light_direction) > B2_DELTA_DISTRIBUTION_ALIGNEMENT_THRESHOLD;
        if (incident_light_info == B2_LOBE_GLASS_REFRACT && incident_direction_is_perfect_refraction)
            return B2_DELTA_DISTRIBUTION_HIGH_VALUE;
        return 0.0f;
    }
    float dot_prod = HoL + HoV / relative_eta;
    float dot_prod2 = dot_prod * dot_prod;
    float alpha_x, alpha_y;
    material_alphas(roughness, material.anisotropy, alpha_x, alpha_y);
    float D = GGX_anisotropic(alpha_x, alpha_y, local_halfway_vector);
    float G1_V = G1_Smith(alpha_x, alpha_y, local_view_direction);
    float dwm_dwi = fabsf(HoL) / dot_prod2;
    float D_pdf = G1_V / fabsf(NoV) * D * fabsf(HoV);
    return dwm_dwi * D_pdf;
}

/* Heitz 2018, "Sampling the GGX Distribution of Visible Normals" */
B2_HD V3 GGX_VNDF_sample(V3 local_view_direction, float alpha_x, float alpha_y, Rng& rng)
{
    float r1 = rng.next_float();
    float r2 = rng.next_float();

    V3 Vh = normalize(v3(alpha_x * local_view_direction.x, alpha_y * local_view_direction.y, local_view_direction.z));
    float lensq = Vh.x * Vh.x + Vh.y * Vh.y;
    V3 T1 = lensq > 0.0f ? v3(-Vh.y, Vh.x, 0.0f) / sqrtf(lensq) : v3(1.0f, 0.0f, 0.0f);
    V3 T2 = cross(Vh, T1);

    float r = sqrtf(r1);
    float phi = B2_TWO_PI * r2;
    b2w sin_phi_w, cos_phi_w;
    b2w_sincos(phi, sin_phi_w, cos_phi_w);
    float t1 = (float)((b2w)r * cos_phi_w);
    float t2 = (float)((b2w)r * sin_phi_w);
    float s = 0.5f * (1.0f + Vh.z);
    t2 = (float)((b2w)(1.0f - s) * b2w_sqrt(1.0f - t1 * t1) + (b2w)(s * t2));

    V3 Nh = t1 * T1 + t2 * T2 + sqrtf(b2_max(0.0f, 1.0f - t1 * t1 - t2 * t2)) * Vh;
    return normalize(v3(alpha_x * Nh.x, alpha_y * Nh.y, b2_max(0.0f, Nh.z)));
}

B2_HD V3 GGX_anisotropic_sample_microfacet(V3 local_view_direction, float alpha_x, float alpha_y, Rng& rng)
{
    if (alpha_x <= B2_ROUGHNESS_CLAMP && alpha_y <= B2_ROUGHNESS_CLAMP)
        return v3(0.0f, 0.0f, 1.0f);
    return GGX_VNDF_sample(local_view_direction, alpha_x, alpha_y, rng);
}

B2_HD V3 microfacet_GGX_sample_reflection(float roughness, float anisotropy, V3 local_view_direction, Rng& rng)
{
    int below_normal = (local_view_direction.z < 0) ? -1 : 1;
    float alpha_x, alpha_y;
    material_alphas(roughness, anisotropy, alpha_x, alpha_y);
    V3 microfacet_normal = GGX_anisotropic_sample_microfacet(local_view_direction * (float)below_normal, alpha_x, alpha_y, rng);
    V3 sampled_direction = reflect_ray(local_view_direction, microfacet_normal * (float)below_normal);
    return normalize(sampled_direction);
}

/* ------------------------------------------------------------------------------------------------------------
 * Glass energy compensation (MicrofacetEnergyCompensation.h:101-703)
 * ------------------------------------------------------------------------------------------------------------ */
#if defined(__CUDACC__)
static __device__ const B2ExponentCurve b2_glass_exponent_lower_device[B2_GLASS_EXPONENT_LOWER_COUNT] = B2_GLASS_EXPONENT_LOWER_INIT;
static __device__ const B2ExponentCurve b2_glass_exponent_higher_device[B2_GLASS_EXPONENT_HIGHER_COUNT] = B2_GLASS_EXPONENT_HIGHER_INIT;
#endif
static const B2ExponentCurve b2_glass_exponent_lower_host[B2_GLASS_EXPONENT_LOWER_COUNT] = B2_GLASS_EXPONENT_LOWER_INIT;
static const B2ExponentCurve b2_glass_exponent_higher_host[B2_GLASS_EXPONENT_HIGHER_COUNT] = B2_GLASS_EXPONENT_HIGHER_INIT;

/* compile-time copies of the curves' eta bounds (b2_glass_exponent_table.h; tests/test_abi.py keeps them in step): the bracket
 * searches below unroll into compare chains on immediates; only the one segment a thread ends up with is fetched from the table */
#define B2_GLASS_LOWER_BOUND(i) ((i) == 0 ? 1.01f : (i) == 1 ? 1.02f : (i) == 2 ? 1.03f : (i) == 3 ? 1.1f : (i) == 4 ? 1.2f : (i) == 5 ? 1.4f : (i) == 6 ? 1.5f : (i) == 7 ? 2.0f : 2.4f)
#define B2_GLASS_HIGHER_BOUND(i) ((i) == 0 ? 1.01f : (i) == 1 ? 1.02f : (i) == 2 ? 1.03f : (i) == 3 ? 1.1f : (i) == 4 ? 1.2f : (i) == 5 ? 1.4f : (i) == 6 ? 1.5f : (i) == 7 ? 2.0f : (i) == 8 ? 2.4f : 3.0f)

/* index of the roughness segment: the first k with roughness <= k * 0.1f (the reference's thresholds are the fp32 literals
 * 0.0f, 0.1f, ... 1.0f); 11 = none (roughness > 1: the reference's chains leave the initial value untouched) */
B2_HD int glass_exponent_segment(float roughness)
{
    int k = 11;
    k = roughness <= 1.0f ? 10 : k;
    k = roughness <= 0.9f ? 9 : k;
    k = roughness <= 0.8f ? 8 : k;
    k = roughness <= 0.7f ? 7 : k;
    k = roughness <= 0.6f ? 6 : k;
    k = roughness <= 0.5f ? 5 : k;
    k = roughness <= 0.4f ? 4 : k;
    k = roughness <= 0.3f ? 3 : k;
    k = roughness <= 0.2f ? 2 : k;
    k = roughness <= 0.1f ? 1 : k;
    k = roughness <= 0.0f ? 0 : k;
    return k;
}

B2_HD float glass_exponent_curve(const B2ExponentCurve& curve, int k, float roughness, float fallback)
{
    if (k > 10)
        return fallback;
    const B2ExponentSegment s = curve.seg[k];
    if (!s.lerp)
        return s.a;
    return b2_lerp(s.a, s.b, (roughness - s.x0) / 0.1f);
}

B2_HD_NOINLINE float glass_energy_compensation_correction_exponent(float roughness, float relative_eta)
{
    if (b2_is_zero(roughness) || fabsf(1.0f - relative_eta) < 1.0e-3f)
        return 2.5f;

    /* namespace-scope tables (global memory on the device, .rodata on the host): a function-local const array is
     * re-materialised on the thread's local-memory stack at every call (3.4 KB of stores per call, measured as
     * 1.4 GB of DRAM writes per C2 sample in profiles/r01_ncu_c2_megakernel_first.md) */
#if defined(__CUDA_ARCH__)
    const B2ExponentCurve* lower_curves = b2_glass_exponent_lower_device;
    const B2ExponentCurve* higher_curves = b2_glass_exponent_higher_device;
#else
    const B2ExponentCurve* lower_curves = b2_glass_exponent_lower_host;
    const B2ExponentCurve* higher_curves = b2_glass_exponent_higher_host;
#endif
    const int k = glass_exponent_segment(roughness);

    /* lower chain: bracket (bound[i], bound[i + 1]] uses lower_curves[i]; outside every bracket the defaults stay.
     * The brackets are tested from the last to the first so that the first match of the reference's chain wins. */
    int lower_index = -1;
    float lower_relative_eta_bound = 1.01f;
#pragma unroll
    for (int i = B2_GLASS_EXPONENT_LOWER_COUNT - 1; i >= 0; i--)
    {
        const float lo = B2_GLASS_LOWER_BOUND(i);
        const float hi = B2_GLASS_HIGHER_BOUND(i + 1);
        const bool inside = relative_eta > lo && relative_eta <= hi;
        lower_index = inside ? i : lower_index;
        lower_relative_eta_bound = inside ? lo : lower_relative_eta_bound;
    }
    float lower_correction = 2.5f;
    if (lower_index >= 0)
        lower_correction = glass_exponent_curve(lower_curves[lower_index], k, roughness, 2.5f);

    /* higher chain: the first bound >= relative_eta, the last curve beyond all bounds */
    int j = B2_GLASS_EXPONENT_HIGHER_COUNT - 1;
    float higher_relative_eta_bound = B2_GLASS_HIGHER_BOUND(B2_GLASS_EXPONENT_HIGHER_COUNT - 1);
#pragma unroll
    for (int i = B2_GLASS_EXPONENT_HIGHER_COUNT - 2; i >= 0; i--)
    {
        const float bound = B2_GLASS_HIGHER_BOUND(i);
        const bool below = relative_eta <= bound;
        j = below ? i : j;
        higher_relative_eta_bound = below ? bound : higher_relative_eta_bound;
    }
    float higher_correction = glass_exponent_curve(higher_curves[j], k, roughness, 2.5f);

    if (higher_relative_eta_bound == lower_relative_eta_bound)
        return lower_correction;
    return b2_lerp(lower_correction, higher_correction, (relative_eta - lower_relative_eta_bound) / (higher_relative_eta_bound - lower_relative_eta_bound));
}

B2_HD float ggx_energy_compensation_dielectrics(const b2pt_render_data& rd, const Mat& material, float custom_roughness, bool inside_object, float eta_t, float eta_i,
                                                float relative_eta, float NoV, int current_bounce)
{
    bool smooth_enough = custom_roughness <= rd.bsdfs_data.energy_compensation_roughness_threshold;
    bool max_bounce_reached = current_bounce > rd.bsdfs_data.glass_energy_compensation_max_bounce && rd.bsdfs_data.glass_energy_compensation_max_bounce > -1;
    if (!material.do_glass_energy_compensation || smooth_enough || max_bounce_reached)
        return 1.0f;

    float compensation_term = 1.0f;
    if (material.thin_film < 1.0f)
    {
        float relative_eta_for_correction = inside_object ? 1.0f / relative_eta : relative_eta;
        float exponent_correction = 2.5f;
        if (!material.thin_walled)
            exponent_correction = glass_energy_compensation_correction_exponent(custom_roughness, relative_eta_for_correction);

        float view_direction_tex_fetch = b2f_pow(b2_max(1.0e-3f, NoV), 1.0f / exponent_correction);
        float F0 = F0_from_eta(eta_t, eta_i);
        float F0_remapped = (float)b2w_sqrt(b2w_sqrt(F0));

        const void* texture;
        if (material.thin_walled)
            texture = rd.bsdfs_data.GGX_thin_glass_directional_albedo;
        else
            texture = inside_object ? rd.bsdfs_data.GGX_glass_directional_albedo_inverse : rd.bsdfs_data.GGX_glass_directional_albedo;
        compensation_term = lut_fetch_3d(texture, view_direction_tex_fetch, custom_roughness, F0_remapped);
        compensation_term = b2_lerp(compensation_term, 1.0f, material.thin_film * custom_roughness);
    }
    return compensation_term;
}

/* ------------------------------------------------------------------------------------------------------------
 * Glossy-base / clearcoat energy compensation (PrincipledEnergyCompensation.h)
 * ------------------------------------------------------------------------------------------------------------ */
B2_HD float principled_specular_relative_ior(const Mat& material, float incident_medium_ior)
{
    if (material.coat == 0.0f)
        return material.ior;
    float incident_layer_ior = b2_lerp(incident_medium_ior, material.coat_ior, material.coat);
    float relative_ior = material.ior / incident_layer_ior;
    if (relative_ior < 1.0f)
        relative_ior = 1.0f / relative_ior;
    return relative_ior;
}

B2_HD float energy_compensation_glossy_base(const b2pt_render_data& rd, const Mat& material, float incident_medium_ior, float NoV, int current_bounce)
{
    bool energy_compensation_disabled = !material.do_specular_energy_compensation;
    bool roughness_low_enough = material.roughness < rd.bsdfs_data.energy_compensation_roughness_threshold;
    bool no_specular_layer = material.specular < 1.0e-3f;
    bool max_bounce_reached = current_bounce > rd.bsdfs_data.glossy_base_energy_compensation_max_bounce && rd.bsdfs_data.glossy_base_energy_compensation_max_bounce > -1;
    bool invalid_view_direction = NoV < 0.0f;
    if (energy_compensation_disabled || roughness_low_enough || no_specular_layer || max_bounce_reached || invalid_view_direction)
        return 1.0f;

    float ior = material.ior;
    float relative_ior = principled_specular_relative_ior(material, incident_medium_ior);
    if (fabsf(relative_ior - 1.0f) < 1.0e-3f)
        relative_ior += 1.0e-3f;

    float view_dir_remapped = (float)b2w_pow(NoV, 1.0f / 2.5f);
    float F0_remapped = (float)b2w_sqrt(b2w_sqrt(F0_from_eta_t_and_relative_ior(ior, relative_ior)));
    float multiple_scattering_compensation = lut_fetch_3d(rd.bsdfs_data.glossy_dielectric_directional_albedo, view_dir_remapped, material.roughness, F0_remapped);

    float ms_compensation = b2_lerp(1.0f, multiple_scattering_compensation, material.specular);
    ms_compensation = b2_lerp(ms_compensation, 1.0f, material.thin_film);
    return ms_compensation;
}

B2_HD float energy_compensation_clearcoat_lobe(const b2pt_render_data& rd, const Mat& material, float incident_medium_ior, float NoV, int current_bounce)
{
    bool energy_compensation_disabled = !material.do_specular_energy_compensation;
    bool no_coat_layer = material.coat < 1.0e-3f;
    bool max_bounce_reached = current_bounce > rd.bsdfs_data.clearcoat_energy_compensation_max_bounce && rd.bsdfs_data.clearcoat_energy_compensation_max_bounce > -1;
    bool invalid_view_direction = NoV < 0.0f;
    if (energy_compensation_disabled || no_coat_layer || max_bounce_reached || invalid_view_direction)
        return 1.0f;

    if (fabsf(material.coat_ior / incident_medium_ior - 1.0f) < 1.0e-3f)
        incident_medium_ior += 1.0e-3f;

    float view_dir_remapped = (float)b2w_pow(NoV, 1.0f / 2.5f);
    float F0_remapped = (float)b2w_sqrt(b2w_sqrt(F0_from_eta(material.coat_ior, incident_medium_ior)));
    float multiple_scattering_compensation = lut_fetch_3d(rd.bsdfs_data.glossy_dielectric_directional_albedo, view_dir_remapped, material.coat_roughness, F0_remapped);

    float ms_compensation = b2_lerp(1.0f, multiple_scattering_compensation, material.coat * (1.0f - material.specular_transmission));
    ms_compensation = b2_lerp(ms_compensation, 1.0f, material.thin_film);
    return ms_compensation;
}

/* ------------------------------------------------------------------------------------------------------------
 * Sheen (Zeltner et al. 2022 LTC fit; SheenLTC.h)
 * ------------------------------------------------------------------------------------------------------------ */
B2_HD float sheen_eval_ltc(V3 d, V3 AiBiRi)
{
    V3 o = v3(d.x * AiBiRi.x + d.z * AiBiRi.y, d.y * AiBiRi.x, d.z);
    float len = length(o);
    o = o / len;
    float M_inv_determinant = AiBiRi.x * AiBiRi.x;
    float jacobian = M_inv_determinant / (len * len * len);
    return o.z * B2_INV_PI * jacobian;
}

B2_HD V3 sheen_read_LTC_parameters(const b2pt_render_data& rd, float roughness, float cos_theta)
{
    return lut_fetch_2d(rd.bsdfs_data.sheen_ltc_parameters_texture, cos_theta, b2_clamp(0.0f, 1.0f, roughness));
}

B2_HD float sheen_get_phi(V3 direction)
{
    float p = (float)b2w_atan2(direction.y, direction.x);
    if (p < 0.0f)
        p += B2_TWO_PI;
    return p;
}

B2_HD V3 sheen_rotate_vector(V3 vec, V3 axis, float angle)
{
    b2w sin_angle_w, cos_angle_w;
    b2w_sincos(angle, sin_angle_w, cos_angle_w);
    float sin_angle = (float)sin_angle_w;
    float cos_angle = (float)cos_angle_w;
    return vec * cos_angle + axis * dot(vec, axis) * (1.0f - cos_angle) + sin_angle * cross(axis, vec);
}

B2_HD Rgb sheen_ltc_eval(const b2pt_render_data& rd, const Mat& material, V3 local_to_light_direction, V3 local_view_direction, float& out_pdf, float& out_sheen_reflectance)
{
    if (local_view_direction.z <= 0.0f || local_to_light_direction.z <= 0.0f)
    {
        out_pdf = 0.0f;
        if (local_view_direction.z > 0.0f)
            out_sheen_reflectance = sheen_read_LTC_parameters(rd, material.sheen_roughness, local_view_direction.z).z;
        else
            out_sheen_reflectance = 0.0f;
        return rgb(0.0f);
    }
    float phi = sheen_get_phi(local_view_direction);
    V3 to_light_standard_frame = sheen_rotate_vector(local_to_light_direction, v3(0.0f, 0.0f, 1.0f), -phi);
    V3 AiBiRi = sheen_read_LTC_parameters(rd, material.sheen_roughness, local_view_direction.z);
    float Do = sheen_eval_ltc(to_light_standard_frame, AiBiRi);
    out_pdf = Do;
    out_sheen_reflectance = AiBiRi.z;
    return material.sheen_color * AiBiRi.z * Do / local_to_light_direction.z;
}

B2_HD float sheen_ltc_pdf(const b2pt_render_data& rd, const Mat& material, V3 local_to_light_direction, V3 local_view_direction)
{
    if (local_view_direction.z <= 0.0f || local_to_light_direction.z <= 0.0f)
        return 0.0f;
    float phi = sheen_get_phi(local_view_direction);
    V3 to_light_standard_frame = sheen_rotate_vector(local_to_light_direction, v3(0.0f, 0.0f, 1.0f), -phi);
    V3 AiBiRi = sheen_read_LTC_parameters(rd, material.sheen_roughness, local_view_direction.z);
    return sheen_eval_ltc(to_light_standard_frame, AiBiRi);
}

B2_HD V3 sheen_ltc_sample(const b2pt_render_data& rd, const Mat& material, V3 local_view_direction, Rng& rng)
{
    V3 cosine_sample = cosine_weighted_sample_z_up_frame(rng);
    V3 AiBiRi = sheen_read_LTC_parameters(rd, material.sheen_roughness, local_view_direction.z);
    float Ai_inv = 1.0f / AiBiRi.x;
    float Bi = AiBiRi.y;
    V3 sampled_direction_ltc_space = normalize(v3(cosine_sample.x * Ai_inv - cosine_sample.z * Bi * Ai_inv, cosine_sample.y * Ai_inv, cosine_sample.z));
    return sheen_rotate_vector(sampled_direction_ltc_space, v3(0.0f, 0.0f, 1.0f), sheen_get_phi(local_view_direction));
}

/* ------------------------------------------------------------------------------------------------------------
 * Spectral dispersion of the glass lobe (Dispersion.h:433-522, WavelengthToRGBMethod = WAVELENGTH_TO_RGB_FIT)
 * ------------------------------------------------------------------------------------------------------------ */
/* compute_dispersion_ior (Dispersion.h:485-499): the IOR at `wavelength` of a medium whose base IOR is measured at 587.6 nm */
B2_HD float compute_dispersion_ior(float dispersion_abbe_number, float dispersion_scale, float base_IOR, float wavelength)
{
    if (dispersion_scale == 0.0f)
        return base_IOR;
    float abbe_number = dispersion_abbe_number / dispersion_scale;
    float B = (base_IOR - 1.0f) / (abbe_number * 0.00000191038851931481f);
    float A = base_IOR - B / (334777.96f);
    return A + B / (wavelength * wavelength);
}

/* wavelength_to_RGB_fit (Dispersion.h:433-458). The reference writes `exp` in the red and blue terms: unqualified, i.e. the double
 * function with the rest of the expression carried in double (b2_math.h "wide evaluation"); `expf` / `logf` stay float. */
B2_HD Rgb wavelength_to_RGB_fit(float wavelength)
{
    Rgb RGB;
    if (wavelength < 463.0f)
    {
        RGB.r = (float)((b2w)-1.2776028240727566e-01f / ((b2w)1.0f + b2w_exp((wavelength - 4.2680623367293401e+02f) / 8.2197460736637176e+00f)) +
                        (b2w)-1.3925673552505122e-11f * b2w_exp((wavelength - 45.0f) / 1.8175459086411596e+01f));
        RGB.r += 1.2898689750552100e-01f;
    }
    else if (wavelength > 553.0f)
    {
        RGB.r = (float)((b2w)(1.7963649137825513e+01f * (1.0f / 2.6577826611702449e+01f)) *
                        b2w_exp(-0.5f * b2_square((wavelength - 6.0625724092824566e+02f) * (1.0f / 2.6577826611702449e+01f))));
        RGB.r += 2.5574660155104657e-03f;
    }
    else
        RGB.r = 0.0f;
    RGB.g = 3.4962267376163049e+02f * b2f_exp_rounded(-0.5f * b2_square((wavelength - 5.4209217455705152e+02f) / -2.9598170255834638e+01f));
    RGB.g /= wavelength;
    RGB.b = (float)b2w_exp(3.2987659944421112e+03f + (-2.0975839709372405e+05f / wavelength) - 4.6368268395094020e+02f * b2f_log_rounded(wavelength));
    return RGB * 10.0f;
}

/* get_dispersion_ray_color (Dispersion.h:512-522): the colour filter of the sampled wavelength, applied once per path (a negative
 * wavelength = not applied yet; it is made positive here) */
B2_HD Rgb get_dispersion_ray_color(float& wavelength, float dispersion_scale)
{
    if (dispersion_scale == 0.0f)
        return rgb(1.0f);
    if (wavelength >= 0.0f)
        return rgb(1.0f);
    wavelength *= -1.0f;
    return wavelength_to_RGB_fit(wavelength);
}

/* ------------------------------------------------------------------------------------------------------------
 * Lobes (Principled.h:49-881)
 * ------------------------------------------------------------------------------------------------------------ */
B2_HD float medium_ior(const b2pt_render_data& rd, int material_index)
{
    /* material_index is -1 for a surface rebuilt from a G-buffer entry that holds no hit (ReSTIR neighbour whose
     * previous-frame pixel missed): the reference then reads ior[-1] (undefined); the device treats it as air */
    if (material_index < 0)
        return 1.0f;
    return material_index == (int)B2_MAX_MATERIAL_INDEX ? 1.0f : rd.buffers.materials_buffer.ior[material_index];
}

B2_HD Rgb principled_coat_eval(const b2pt_render_data& rd, const BsdfCtx& ctx, V3 local_view_direction, V3 local_to_light_direction, V3 local_halfway_vector,
                               float incident_medium_ior, float& out_pdf)
{
    const Mat& m = *ctx.material;
    float HoL = b2_clamp(1.0e-8f, 1.0f, dot(local_halfway_vector, local_to_light_direction));
    float regularized_roughness = regularize_reflection(rd, ctx.regularization_mode, m.coat_roughness, ctx.accumulated_path_roughness);
    int coat_delta_direction_sampled = is_specular_delta_reflection_sampled(m, regularized_roughness, m.coat_anisotropy, *ctx.incident_light_info);
    Rgb F = rgb(full_fresnel_dielectric(HoL, incident_medium_ior, m.coat_ior));
    return ggx_eval_reflect(rd, regularized_roughness, m.coat_anisotropy, F, local_view_direction, local_to_light_direction, local_halfway_vector, out_pdf, coat_delta_direction_sampled);
}

B2_HD float principled_coat_pdf(const b2pt_render_data& rd, const BsdfCtx& ctx, V3 local_view_direction, V3 local_to_light_direction, V3 local_halfway_vector)
{
    const Mat& m = *ctx.material;
    float regularized_roughness = regularize_reflection(rd, ctx.regularization_mode, m.coat_roughness, ctx.accumulated_path_roughness);
    int coat_delta_direction_sampled = is_specular_delta_reflection_sampled(m, regularized_roughness, m.coat_anisotropy, *ctx.incident_light_info);
    return ggx_pdf_reflect(regularized_roughness, m.coat_anisotropy, local_view_direction, local_to_light_direction, local_halfway_vector, coat_delta_direction_sampled);
}

B2_HD Rgb principled_metallic_eval(const b2pt_render_data& rd, const BsdfCtx& ctx, float roughness, float anisotropy, float incident_ior, V3 local_view_direction,
                                   V3 local_to_light_direction, V3 local_half_vector, float& pdf)
{
    const Mat& m = *ctx.material;
    float regularized_roughness = regularize_reflection(rd, ctx.regularization_mode, roughness, ctx.accumulated_path_roughness);
    int metal_delta_direction_sampled = is_specular_delta_reflection_sampled(m, regularized_roughness, anisotropy, *ctx.incident_light_info);
    if (metal_delta_direction_sampled == B2_SPECULAR_PEAK_NOT_SAMPLED)
    {
        pdf = 0.0f;
        return rgb(0.0f);
    }
    float HoL = b2_clamp(1.0e-8f, 1.0f, dot(local_half_vector, local_to_light_direction));
    Rgb F_metal = adobe_f82_tint_fresnel(m.base_color, m.metallic_F82, m.metallic_F90, m.metallic_F90_falloff_exponent, HoL);
    Rgb F_thin_film = thin_film_fresnel(m, incident_ior, HoL);
    Rgb F = rgb_lerp(F_metal, F_thin_film, m.thin_film);
    return ggx_eval_reflect_compensated(rd, regularized_roughness, anisotropy, m.do_metallic_energy_compensation, F, local_view_direction, local_to_light_direction,
                                        local_half_vector, pdf, metal_delta_direction_sampled, ctx.current_bounce);
}

B2_HD float principled_metallic_pdf(const b2pt_render_data& rd, const BsdfCtx& ctx, float roughness, float anisotropy, V3 local_view_direction, V3 local_to_light_direction,
                                    V3 local_half_vector)
{
    float regularized_roughness = regularize_reflection(rd, ctx.regularization_mode, roughness, ctx.accumulated_path_roughness);
    int metal_delta_direction_sampled = is_specular_delta_reflection_sampled(*ctx.material, regularized_roughness, anisotropy, *ctx.incident_light_info);
    if (metal_delta_direction_sampled == B2_SPECULAR_PEAK_NOT_SAMPLED)
        return 0.0f;
    return ggx_pdf_reflect(regularized_roughness, anisotropy, local_view_direction, local_to_light_direction, local_half_vector, metal_delta_direction_sampled);
}

B2_HD Rgb principled_specular_fresnel(const Mat& material, float relative_specular_ior, float cos_theta_i)
{
    float layer_above_IOR = material.ior / relative_specular_ior;
    float material_thin_film = material.thin_film;
    Rgb F_specular = rgb(0.0f);
    if (material_thin_film < 1.0f)
        F_specular = rgb(full_fresnel_dielectric(cos_theta_i, relative_specular_ior));
    Rgb F_thin_film = thin_film_fresnel(material, layer_above_IOR, cos_theta_i);
    return rgb_lerp(F_specular, F_thin_film, material_thin_film);
}

B2_HD Rgb principled_specular_eval(const b2pt_render_data& rd, const BsdfCtx& ctx, float relative_ior, V3 local_view_direction, V3 local_to_light_direction, V3 local_half_vector, float& pdf)
{
    const Mat& m = *ctx.material;
    float regularized_roughness = regularize_reflection(rd, ctx.regularization_mode, m.roughness, ctx.accumulated_path_roughness);
    int delta_sampled = is_specular_delta_reflection_sampled(m, regularized_roughness, m.anisotropy, *ctx.incident_light_info);
    Rgb F = principled_specular_fresnel(m, relative_ior, dot(local_to_light_direction, local_half_vector));
    return ggx_eval_reflect(rd, regularized_roughness, m.anisotropy, F, local_view_direction, local_to_light_direction, local_half_vector, pdf, delta_sampled);
}

B2_HD float principled_specular_pdf(const b2pt_render_data& rd, const BsdfCtx& ctx, V3 local_view_direction, V3 local_to_light_direction, V3 local_half_vector)
{
    const Mat& m = *ctx.material;
    float regularized_roughness = regularize_reflection(rd, ctx.regularization_mode, m.roughness, ctx.accumulated_path_roughness);
    int delta_sampled = is_specular_delta_reflection_sampled(m, regularized_roughness, m.anisotropy, *ctx.incident_light_info);
    return ggx_pdf_reflect(regularized_roughness, m.anisotropy, local_view_direction, local_to_light_direction, local_half_vector, delta_sampled);
}

B2_HD Rgb principled_beer_absorption(const b2pt_render_data& rd, VolumeState& vs)
{
    Rgb absorption_color;
    if (rd.bsdfs_data.white_furnace_mode)
        absorption_color = rgb(1.0f);
    else
        absorption_color = unpack_rgb24(rd.buffers.materials_buffer.absorption_color_packed[vs.incident_mat_index]);
    if (!(absorption_color.r == 1.0f && absorption_color.g == 1.0f && absorption_color.b == 1.0f))
    {
        vs.distance_in_volume = b2_max(vs.distance_in_volume, 1.0e-6f);
        Rgb absorption_coefficient = rgb_log(absorption_color) / rd.buffers.materials_buffer.absorption_at_distance[vs.incident_mat_index];
        return rgb_exp(absorption_coefficient * vs.distance_in_volume);
    }
    return rgb(1.0f);
}

/* Shared front half of principled_glass_eval / _pdf (Principled.h:359-483 and :560-684 are the same statements) */
struct GlassSetup
{
    bool valid;
    bool reflecting;
    bool thin_walled;
    float eta_i, eta_t, relative_eta;
    float scaled_roughness;
    float f_reflect_proba;
    V3 local_half_vector;
    Rgb F;
};

B2_HD GlassSetup principled_glass_setup(const b2pt_render_data& rd, const BsdfCtx& ctx, V3 local_view_direction, V3 local_to_light_direction)
{
    const Mat& m = *ctx.material;
    GlassSetup g;
    g.valid = false;

    float NoV = local_view_direction.z;
    float NoL = local_to_light_direction.z;
    if (fabsf(NoL) < 1.0e-8f)
        return g;

    g.reflecting = NoL * NoV > 0;
    g.eta_i = medium_ior(rd, ctx.volume_state->incident_mat_index);
    g.eta_t = medium_ior(rd, ctx.volume_state->outgoing_mat_index);
    /* Principled.h:377-380, 578-581 */
    g.eta_i = compute_dispersion_ior(m.dispersion_abbe_number, m.dispersion_scale, g.eta_i, fabsf(ctx.volume_state->sampled_wavelength));
    g.eta_t = compute_dispersion_ior(m.dispersion_abbe_number, m.dispersion_scale, g.eta_t, fabsf(ctx.volume_state->sampled_wavelength));
    g.relative_eta = g.eta_t / g.eta_i;
    if (fabsf(g.relative_eta - 1.0f) < 1.0e-5f)
        g.relative_eta = 1.0f + 1.0e-5f;

    g.thin_walled = m.thin_walled;
    g.scaled_roughness = thin_walled_roughness(g.thin_walled, m.roughness, g.relative_eta);

    /* microfacet regularization is ON, so the specular fast path of the reference (:411) is compiled out */
    V3 h;
    if (g.reflecting)
        h = local_to_light_direction + local_view_direction;
    else if (g.thin_walled)
        h = local_to_light_direction * v3(1.0f, 1.0f, -1.0f) + local_view_direction;
    else
        h = local_to_light_direction * g.relative_eta + local_view_direction;
    h = normalize(h);
    if (h.z < 0.0f)
        h = -h;
    g.local_half_vector = h;

    float HoL = dot(local_to_light_direction, h);
    float HoV = dot(local_view_direction, h);
    if (HoL * NoL < 0.0f || HoV * NoV < 0.0f)
        return g;

    float thin_film = m.thin_film;
    Rgb F_thin_film = thin_film_fresnel(m, g.eta_i, HoV);
    Rgb F_no_thin_film = rgb(0.0f);
    if (thin_film < 1.0f)
        F_no_thin_film = rgb(full_fresnel_dielectric(HoV, g.relative_eta));
    Rgb F = rgb_lerp(F_no_thin_film, F_thin_film, thin_film);
    if (g.thin_walled && F.r < 1.0f && thin_film == 0.0f && g.scaled_roughness < 0.1f)
        F += rgb(b2_square(1.0f - F.r) * F.r / (1.0f - b2_square(F.r)));
    g.F = F;
    g.f_reflect_proba = luminance(F);
    g.valid = true;
    return g;
}

B2_HD float glass_regularized_roughness(const b2pt_render_data& rd, const BsdfCtx& ctx, const GlassSetup& g, bool reflecting)
{
    if (ctx.regularization_mode == B2_REGULARIZATION_MIS)
        return regularize_mix_reflection_refraction(rd, ctx.regularization_mode, g.scaled_roughness, ctx.accumulated_path_roughness, g.eta_i, g.eta_t);
    if (ctx.regularization_mode == B2_REGULARIZATION_CLASSIC)
    {
        if (reflecting)
            return regularize_reflection(rd, ctx.regularization_mode, g.scaled_roughness, ctx.accumulated_path_roughness);
        return regularize_refraction(rd, ctx.regularization_mode, g.scaled_roughness, ctx.accumulated_path_roughness, g.eta_i, g.eta_t);
    }
    return g.scaled_roughness;
}

B2_HD_NOINLINE Rgb principled_glass_eval(const b2pt_render_data& rd, BsdfCtx& ctx, V3 local_view_direction, V3 local_to_light_direction, float& pdf)
{
    pdf = 0.0f;
    const Mat& m = *ctx.material;
    VolumeState& vs = *ctx.volume_state;
    GlassSetup g = principled_glass_setup(rd, ctx, local_view_direction, local_to_light_direction);
    if (!g.valid)
        return rgb(0.0f);

    Rgb color;
    if (g.reflecting)
    {
        float regularized_roughness = glass_regularized_roughness(rd, ctx, g, true);
        int delta_glass_direction_sampled = is_specular_delta_reflection_sampled(m, g.scaled_roughness, m.anisotropy, *ctx.incident_light_info);
        color = ggx_eval_reflect(rd, regularized_roughness, m.anisotropy, g.F, local_view_direction, local_to_light_direction, g.local_half_vector, pdf, delta_glass_direction_sampled);
        float compensation_term = ggx_energy_compensation_dielectrics(rd, m, m.roughness, vs.inside_material, g.eta_t, g.eta_i, g.relative_eta, local_view_direction.z, ctx.current_bounce);
        color /= compensation_term;
        pdf *= g.f_reflect_proba;
    }
    else
    {
        float regularized_roughness = glass_regularized_roughness(rd, ctx, g, false);
        color = ggx_eval_refract(m, regularized_roughness, g.relative_eta, g.F, local_view_direction, local_to_light_direction, g.local_half_vector, pdf, *ctx.incident_light_info);
        pdf *= 1.0f - g.f_reflect_proba;
        float compensation_term = ggx_energy_compensation_dielectrics(rd, m, regularized_roughness, vs.inside_material, g.eta_t, g.eta_i, g.relative_eta, local_view_direction.z, ctx.current_bounce);
        color /= compensation_term;

        if (g.thin_walled)
            color *= m.base_color;

        if (g.thin_walled && ctx.update_ray_volume_state)
            vs.interior_stack.pop(vs.inside_material);
        else if (vs.incident_mat_index != (int)B2_MAX_MATERIAL_INDEX)
        {
            color *= principled_beer_absorption(rd, vs);
            if (ctx.update_ray_volume_state)
            {
                vs.distance_in_volume = 0.0f;
                if (vs.inside_material)
                    vs.interior_stack.pop(vs.inside_material);
            }
        }
    }
    return color;
}

B2_HD_NOINLINE float principled_glass_pdf(const b2pt_render_data& rd, BsdfCtx& ctx, V3 local_view_direction, V3 local_to_light_direction)
{
    const Mat& m = *ctx.material;
    GlassSetup g = principled_glass_setup(rd, ctx, local_view_direction, local_to_light_direction);
    if (!g.valid)
        return 0.0f;
    float pdf;
    if (g.reflecting)
    {
        float regularized_roughness = glass_regularized_roughness(rd, ctx, g, true);
        int delta_glass_direction_sampled = is_specular_delta_reflection_sampled(m, g.scaled_roughness, m.anisotropy, *ctx.incident_light_info);
        pdf = ggx_pdf_reflect(regularized_roughness, m.anisotropy, local_view_direction, local_to_light_direction, g.local_half_vector, delta_glass_direction_sampled);
        pdf *= g.f_reflect_proba;
    }
    else
    {
        float regularized_roughness = glass_regularized_roughness(rd, ctx, g, false);
        pdf = ggx_pdf_refract(m, regularized_roughness, g.relative_eta, local_view_direction, local_to_light_direction, g.local_half_vector, *ctx.incident_light_info);
        pdf *= 1.0f - g.f_reflect_proba;
    }
    return pdf;
}

/* returns the sampled direction in the local frame of `local_view_direction` */
B2_HD_NOINLINE V3 principled_glass_sample(const b2pt_render_data& rd, BsdfCtx& ctx, V3 local_view_direction, Rng& rng)
{
    const Mat& m = *ctx.material;
    VolumeState& vs = *ctx.volume_state;
    float eta_i = medium_ior(rd, vs.incident_mat_index);
    float eta_t = medium_ior(rd, vs.outgoing_mat_index);
    /* Principled.h:731-734 */
    eta_i = compute_dispersion_ior(m.dispersion_abbe_number, m.dispersion_scale, eta_i, fabsf(vs.sampled_wavelength));
    eta_t = compute_dispersion_ior(m.dispersion_abbe_number, m.dispersion_scale, eta_t, fabsf(vs.sampled_wavelength));
    float relative_eta = eta_t / eta_i;
    if (fabsf(relative_eta - 1.0f) < 1.0e-5f)
        relative_eta = 1.0f + 1.0e-5f;

    bool thin_walled = m.thin_walled;
    float scaled_roughness = thin_walled_roughness(thin_walled, m.roughness, relative_eta);
    if (ctx.regularization_mode == B2_REGULARIZATION_MIS)
        scaled_roughness = regularize_mix_reflection_refraction(rd, ctx.regularization_mode, scaled_roughness, ctx.accumulated_path_roughness, eta_i, eta_t);

    float alpha_x, alpha_y;
    material_alphas(scaled_roughness, m.anisotropy, alpha_x, alpha_y);
    V3 microfacet_normal = GGX_anisotropic_sample_microfacet(local_view_direction, alpha_x, alpha_y, rng);

    float HoV = dot(local_view_direction, microfacet_normal);
    float thin_film = m.thin_film;
    Rgb F_thin_film = thin_film_fresnel(m, eta_i, HoV);
    Rgb F_no_thin_film = rgb(0.0f);
    if (thin_film < 1.0f)
        F_no_thin_film = rgb(full_fresnel_dielectric(HoV, relative_eta));
    Rgb F = rgb_lerp(F_no_thin_film, F_thin_film, thin_film);
    if (thin_walled && F.r < 1.0f && thin_film == 0.0f && scaled_roughness < 0.1f)
        F += rgb(b2_square(1.0f - F.r) * F.r / (1.0f - b2_square(F.r)));
    float f_reflect_proba = luminance(F);

    float rand_1 = rng.next_float();
    V3 sampled_direction;
    if (rand_1 < f_reflect_proba)
    {
        sampled_direction = reflect_ray(local_view_direction, microfacet_normal);
        *ctx.incident_light_info = B2_LOBE_GLASS_REFLECT;
        if (ctx.update_ray_volume_state)
            vs.interior_stack.pop(false);
    }
    else
    {
        *ctx.incident_light_info = B2_LOBE_GLASS_REFRACT;
        if (dot(microfacet_normal, local_view_direction) < 0.0f)
            microfacet_normal = -microfacet_normal;
        if (thin_walled)
        {
            V3 reflected = reflect_ray(local_view_direction, microfacet_normal);
            reflected.z *= -1.0f;
            if (ctx.update_ray_volume_state)
                vs.interior_stack.pop(false);
            return reflected;
        }
        sampled_direction = refract_ray(local_view_direction, microfacet_normal, relative_eta);
    }
    return sampled_direction;
}

B2_HD Rgb principled_diffuse_transmission_eval(const b2pt_render_data& rd, const Mat& material, VolumeState& vs, bool update_ray_volume_state, V3 local_view_direction,
                                               V3 local_to_light_direction, float& diffuse_transmission_pdf)
{
    diffuse_transmission_pdf = 0.0f;
    if (local_view_direction.z * local_to_light_direction.z > 0.0f)
        return rgb(0.0f);

    Rgb color = material.base_color * B2_INV_PI;
    if (vs.incident_mat_index != (int)B2_MAX_MATERIAL_INDEX)
    {
        color *= principled_beer_absorption(rd, vs);
        if (update_ray_volume_state)
        {
            vs.distance_in_volume = 0.0f;
            if (vs.inside_material)
                vs.interior_stack.pop(vs.inside_material);
        }
    }
    diffuse_transmission_pdf = fabsf(local_to_light_direction.z * B2_INV_PI);
    return color;
}

/* ------------------------------------------------------------------------------------------------------------
 * Layering (Principled.h:883-1510)
 * ------------------------------------------------------------------------------------------------------------ */
B2_HD Rgb principled_coat_compute_darkening(const Mat& material, float relative_eta, float view_dir_fresnel)
{
    if (material.coat_darkening == 0.0f)
        return rgb(1.0f);
    float Kr = 1.0f - (1.0f - fresnel_hemispherical_albedo_fit(relative_eta)) / (relative_eta * relative_eta);
    float Ks = view_dir_fresnel;
    float base_roughness = b2_lerp(1.0f, material.roughness, b2_max(material.specular_transmission, b2_max(material.metallic, material.specular)));
    float K = b2_lerp(Ks, Kr, base_roughness);
    float sheen = material.sheen;
    Rgb base_albedo = (material.base_color + material.sheen_color * sheen) / (1.0f + sheen);
    Rgb darkening = (1.0f - K) / (rgb(1.0f) - base_albedo * K);
    return rgb_lerp(rgb(1.0f), darkening, material.coat * material.coat_darkening * (1.0f - material.diffuse_transmission));
}

B2_HD Rgb principled_specular_compute_darkening(const Mat& material, float relative_eta, float view_dir_fresnel)
{
    (void)view_dir_fresnel;
    if (material.specular_darkening == 0.0f)
        return rgb(1.0f);
    float Kr = 1.0f - (1.0f - fresnel_hemispherical_albedo_fit(relative_eta)) / (relative_eta * relative_eta);
    float K = Kr;
    Rgb base_albedo = material.base_color;
    Rgb darkening = (1.0f - K) / (rgb(1.0f) - base_albedo * K);
    return rgb_lerp(rgb(1.0f), darkening, material.specular * material.specular_darkening * (1.0f - material.diffuse_transmission));
}

B2_HD Rgb internal_eval_coat_layer(const b2pt_render_data& rd, const BsdfCtx& ctx, V3 local_view_direction, V3 local_to_light_direction, V3 local_half_vector,
                                   float incident_ior, bool refracting, float coat_weight, float coat_proba, Rgb& layers_throughput, float& out_cumulative_pdf)
{
    const Mat& m = *ctx.material;
    float coat_ior = m.coat_ior;
    if (coat_weight > 0.0f && ((local_view_direction.z > 0.0f && local_to_light_direction.z > 0.0f) || refracting))
    {
        float coat_pdf = 0.0f;
        Rgb contribution = rgb(0.0f);
        if (!refracting)
        {
            contribution = principled_coat_eval(rd, ctx, local_view_direction, local_to_light_direction, local_half_vector, incident_ior, coat_pdf);
            contribution *= coat_weight;
            contribution *= layers_throughput;
        }
        out_cumulative_pdf += coat_pdf * coat_proba;

        Rgb layer_below_attenuation = rgb(1.0f);
        layer_below_attenuation *= 1.0f - full_fresnel_dielectric(fabsf(local_to_light_direction.z), incident_ior, coat_ior);
        float view_dir_fresnel = full_fresnel_dielectric(fabsf(local_view_direction.z), incident_ior, coat_ior);
        layer_below_attenuation *= 1.0f - view_dir_fresnel;

        const Rgb absorption = m.coat_medium_absorption;
        if (!(absorption.r == 1.0f && absorption.g == 1.0f && absorption.b == 1.0f))
        {
            float incident_refracted_angle = b2_max(1.0e-6f, sqrtf(1.0f - (1.0f - local_to_light_direction.z * local_to_light_direction.z) / (coat_ior * coat_ior)));
            float outgoing_refracted_angle = b2_max(1.0e-6f, sqrtf(1.0f - (1.0f - local_view_direction.z * local_view_direction.z) / (coat_ior * coat_ior)));
            float traveled_distance_angle = 1.0f / incident_refracted_angle + 1.0f / outgoing_refracted_angle;
            Rgb coat_absorption = rgb_exp(-(rgb(1.0f) - rgb_pow(rgb_sqrt(absorption), traveled_distance_angle)) * m.coat_medium_thickness);
            layer_below_attenuation *= coat_absorption;
        }

        layer_below_attenuation *= principled_coat_compute_darkening(m, coat_ior / incident_ior, view_dir_fresnel);
        layer_below_attenuation = rgb_lerp(rgb(1.0f), layer_below_attenuation, m.coat);
        layers_throughput *= layer_below_attenuation;
        return contribution;
    }
    return rgb(0.0f);
}

B2_HD float internal_pdf_coat_layer(const b2pt_render_data& rd, const BsdfCtx& ctx, V3 local_view_direction, V3 local_to_light_direction, V3 local_half_vector, bool refracting,
                                    float coat_weight, float coat_proba)
{
    if (coat_weight > 0.0f && ((local_view_direction.z > 0.0f && local_to_light_direction.z > 0.0f) || refracting))
    {
        float coat_pdf = 0.0f;
        if (!refracting)
            coat_pdf = principled_coat_pdf(rd, ctx, local_view_direction, local_to_light_direction, local_half_vector);
        return coat_pdf * coat_proba;
    }
    return 0.0f;
}

B2_HD Rgb internal_eval_sheen_layer(const b2pt_render_data& rd, const Mat& material, V3 local_view_direction, V3 local_to_light_direction, bool refracting, float sheen_weight,
                                    float sheen_proba, Rgb& layers_throughput, float& out_cumulative_pdf)
{
    if ((sheen_weight > 0.0f && local_view_direction.z > 0.0f && local_to_light_direction.z > 0.0f) || refracting)
    {
        float sheen_reflectance;
        float sheen_pdf;
        Rgb contribution = sheen_ltc_eval(rd, material, local_to_light_direction, local_view_direction, sheen_pdf, sheen_reflectance);
        contribution *= sheen_weight;
        contribution *= layers_throughput;
        out_cumulative_pdf += sheen_pdf * sheen_proba;
        layers_throughput *= 1.0f - material.sheen * sheen_reflectance;
        return contribution;
    }
    return rgb(0.0f);
}

B2_HD float internal_pdf_sheen_layer(const b2pt_render_data& rd, const Mat& material, V3 local_view_direction, V3 local_to_light_direction, bool refracting, float sheen_weight, float sheen_proba)
{
    if ((sheen_weight > 0.0f && local_view_direction.z > 0.0f && local_to_light_direction.z > 0.0f) || refracting)
        return sheen_ltc_pdf(rd, material, local_to_light_direction, local_view_direction) * sheen_proba;
    return 0.0f;
}

B2_HD Rgb internal_eval_metal_layer(const b2pt_render_data& rd, const BsdfCtx& ctx, float roughness, float anisotropy, V3 local_view_direction, V3 local_to_light_direction,
                                    V3 local_half_vector, float incident_ior, float metal_weight, float metal_proba, Rgb layers_throughput, float& out_cumulative_pdf)
{
    if (metal_weight > 0.0f && local_view_direction.z > 0.0f && local_to_light_direction.z > 0.0f)
    {
        float metal_pdf = 0.0f;
        Rgb contribution = principled_metallic_eval(rd, ctx, roughness, anisotropy, incident_ior, local_view_direction, local_to_light_direction, local_half_vector, metal_pdf);
        contribution *= metal_weight;
        contribution *= layers_throughput;
        out_cumulative_pdf += metal_pdf * metal_proba;
        return contribution;
    }
    return rgb(0.0f);
}

B2_HD float internal_pdf_metal_layer(const b2pt_render_data& rd, const BsdfCtx& ctx, float roughness, float anisotropy, V3 local_view_direction, V3 local_to_light_direction,
                                     V3 local_half_vector, float metal_weight, float metal_proba)
{
    if (metal_weight > 0.0f && local_view_direction.z > 0.0f && local_to_light_direction.z > 0.0f)
        return principled_metallic_pdf(rd, ctx, roughness, anisotropy, local_view_direction, local_to_light_direction, local_half_vector) * metal_proba;
    return 0.0f;
}

B2_HD Rgb internal_eval_specular_layer(const b2pt_render_data& rd, const BsdfCtx& ctx, V3 local_view_direction, V3 local_to_light_direction, V3 local_half_vector,
                                       float incident_medium_ior, float specular_weight, bool refracting, float specular_proba, Rgb& layers_throughput, float& out_cumulative_pdf)
{
    const Mat& m = *ctx.material;
    if (specular_weight > 0.0f && ((local_view_direction.z > 0.0f && local_to_light_direction.z > 0.0f) || (refracting && *ctx.incident_light_info == B2_LOBE_DIFFUSE_TRANSMISSION)))
    {
        float relative_ior = principled_specular_relative_ior(m, incident_medium_ior);
        float specular_pdf = 0.0f;
        Rgb contribution = principled_specular_eval(rd, ctx, relative_ior, local_view_direction, local_to_light_direction, local_half_vector, specular_pdf);
        contribution *= rgb_lerp(rgb(1.0f), m.specular_tint * m.specular_color, m.specular);
        contribution *= specular_weight;
        contribution *= layers_throughput;

        Rgb layer_below_attenuation = rgb(1.0f);
        Rgb light_dir_fresnel = principled_specular_fresnel(m, relative_ior, fabsf(local_to_light_direction.z));
        light_dir_fresnel *= (1.0f - m.diffuse_transmission);
        layer_below_attenuation *= rgb(1.0f) - light_dir_fresnel;
        Rgb view_dir_fresnel = principled_specular_fresnel(m, relative_ior, fabsf(local_view_direction.z));
        layer_below_attenuation *= rgb(1.0f) - view_dir_fresnel;
        layer_below_attenuation *= principled_specular_compute_darkening(m, relative_ior, luminance(view_dir_fresnel));
        layer_below_attenuation = rgb_lerp(rgb(1.0f), layer_below_attenuation, m.specular);
        layers_throughput *= layer_below_attenuation;

        out_cumulative_pdf += specular_pdf * specular_proba;
        return contribution;
    }
    return rgb(0.0f);
}

B2_HD float internal_pdf_specular_layer(const b2pt_render_data& rd, const BsdfCtx& ctx, V3 local_view_direction, V3 local_to_light_direction, V3 local_half_vector, float specular_weight,
                                        bool refracting, float specular_proba)
{
    if (specular_weight > 0.0f && ((local_view_direction.z > 0.0f && local_to_light_direction.z > 0.0f) || (refracting && *ctx.incident_light_info == B2_LOBE_DIFFUSE_TRANSMISSION)))
        return principled_specular_pdf(rd, ctx, local_view_direction, local_to_light_direction, local_half_vector) * specular_proba;
    return 0.0f;
}

/* ------------------------------------------------------------------------------------------------------------
 * Lobe weights / sampling probabilities (Principled.h:1515-1618)
 * ------------------------------------------------------------------------------------------------------------ */
struct LobeWeights { float coat, sheen, metal_1, metal_2, specular, diffuse, glass, diffuse_transmission; };

B2_HD LobeWeights principled_lobe_weights(const Mat& material, bool outside_object)
{
    LobeWeights w;
    float outside = outside_object ? 1.0f : 0.0f;
    w.coat = material.coat * outside;
    w.sheen = material.sheen * outside;
    float metallic = material.metallic;
    w.metal_1 = metallic * outside;
    w.metal_2 = metallic * outside;
    float second_roughness_weight = material.second_roughness_weight;
    w.metal_1 = b2_lerp(w.metal_1, 0.0f, second_roughness_weight);
    w.metal_2 = b2_lerp(0.0f, w.metal_2, second_roughness_weight);
    float specular_transmission = material.specular_transmission;
    float diffuse_transmission = material.diffuse_transmission;
    w.glass = !outside_object ? (1.0f - diffuse_transmission) : (1.0f - metallic) * (1.0f - diffuse_transmission) * specular_transmission;
    w.diffuse_transmission = !outside_object ? diffuse_transmission : (1.0f - metallic) * diffuse_transmission;
    w.specular = (1.0f - metallic) * (1.0f - specular_transmission * (1.0f - diffuse_transmission)) * material.specular * outside;
    w.diffuse = (1.0f - metallic) * (1.0f - specular_transmission) * (1.0f - diffuse_transmission) * outside;
    return w;
}

B2_HD LobeWeights principled_lobe_probabilities(const LobeWeights& w)
{
    float normalize_factor = 1.0f / (w.coat + w.sheen + w.metal_1 + w.metal_2 + w.specular + w.diffuse + w.glass + w.diffuse_transmission);
    LobeWeights p;
    p.coat = w.coat * normalize_factor;
    p.sheen = w.sheen * normalize_factor;
    p.metal_1 = w.metal_1 * normalize_factor;
    p.metal_2 = w.metal_2 * normalize_factor;
    p.specular = w.specular * normalize_factor;
    p.diffuse = w.diffuse * normalize_factor;
    p.glass = w.glass * normalize_factor;
    p.diffuse_transmission = w.diffuse_transmission * normalize_factor;
    return p;
}

/* ------------------------------------------------------------------------------------------------------------
 * eval / pdf / sample (Principled.h:1620-1939)
 * ------------------------------------------------------------------------------------------------------------ */
struct PrincipledFrames
{
    V3 v, l, h;    /* un-rotated local frame */
    V3 vr, lr, hr; /* frame rotated by the anisotropy rotation */
};

B2_HD PrincipledFrames principled_frames(const BsdfCtx& ctx)
{
    PrincipledFrames f;
    V3 T, B;
    build_ONB(ctx.shading_normal, T, B);
    f.v = world_to_local_frame(T, B, ctx.shading_normal, ctx.view_direction);
    f.l = world_to_local_frame(T, B, ctx.shading_normal, ctx.to_light_direction);
    f.h = normalize(f.v + f.l);
    V3 TR, BR;
    build_rotated_ONB(ctx.shading_normal, TR, BR, ctx.material->anisotropy_rotation * B2_PI);
    f.vr = world_to_local_frame(TR, BR, ctx.shading_normal, ctx.view_direction);
    f.lr = world_to_local_frame(TR, BR, ctx.shading_normal, ctx.to_light_direction);
    f.hr = normalize(f.vr + f.lr);
    return f;
}

B2_HD_NOINLINE Rgb principled_bsdf_eval(const b2pt_render_data& rd, BsdfCtx& ctx, float& pdf)
{
    pdf = 0.0f;
    const Mat& m = *ctx.material;
    VolumeState& vs = *ctx.volume_state;

    bool outside_object = !vs.inside_material;
    bool refracting = dot(ctx.shading_normal, ctx.to_light_direction) < 0.0f && outside_object;
    PrincipledFrames f = principled_frames(ctx);
    float incident_medium_ior = medium_ior(rd, vs.incident_mat_index);

    LobeWeights w = principled_lobe_weights(m, outside_object);
    LobeWeights p = principled_lobe_probabilities(w);
    float not_refracting = refracting ? 0.0f : 1.0f;

    Rgb layers_throughput = rgb(1.0f);
    Rgb final_color = rgb(0.0f);

    final_color += internal_eval_coat_layer(rd, ctx, f.v, f.l, f.h, incident_medium_ior, refracting, w.coat, p.coat, layers_throughput, pdf);
    final_color += internal_eval_sheen_layer(rd, m, f.v, f.l, refracting, w.sheen, p.sheen, layers_throughput, pdf);
    final_color += internal_eval_metal_layer(rd, ctx, m.roughness, m.anisotropy, f.vr, f.lr, f.hr, incident_medium_ior, w.metal_1 * not_refracting, p.metal_1, layers_throughput, pdf);
    final_color += internal_eval_metal_layer(rd, ctx, m.second_roughness, m.anisotropy, f.vr, f.lr, f.hr, incident_medium_ior, w.metal_2 * not_refracting, p.metal_2, layers_throughput, pdf);

    /* glass before the glossy base: it sits next to the specular layer, not below it */
    if (w.glass > 0.0f)
    {
        float glass_pdf = 0.0f;
        Rgb contribution = principled_glass_eval(rd, ctx, f.vr, f.lr, glass_pdf);
        contribution *= w.glass;
        contribution *= layers_throughput;
        pdf += glass_pdf * p.glass;
        final_color += contribution;
    }

    /* glossy base = specular GGX over Lambertian diffuse */
    {
        Rgb glossy_base_contribution = rgb(0.0f);
        glossy_base_contribution += internal_eval_specular_layer(rd, ctx, f.vr, f.lr, f.hr, incident_medium_ior, w.specular, refracting, p.specular, layers_throughput, pdf);
        float diffuse_weight = w.diffuse * not_refracting;
        Rgb diffuse_contribution = rgb(0.0f);
        if (diffuse_weight > 0.0f && f.v.z > 0.0f && f.l.z > 0.0f)
        {
            float diffuse_pdf;
            diffuse_contribution = lambertian_eval(m, f.l.z, diffuse_pdf);
            diffuse_contribution *= diffuse_weight;
            diffuse_contribution *= layers_throughput;
            pdf += diffuse_pdf * p.diffuse;
        }
        glossy_base_contribution += diffuse_contribution;
        float glossy_base_energy_compensation = energy_compensation_glossy_base(rd, m, incident_medium_ior, f.v.z, ctx.current_bounce);
        final_color += glossy_base_contribution / glossy_base_energy_compensation;
    }

    if (w.diffuse_transmission > 0.0f && f.l.z < 0.0f)
    {
        float diffuse_transmission_pdf = 0.0f;
        Rgb contribution = principled_diffuse_transmission_eval(rd, m, vs, ctx.update_ray_volume_state, f.v, f.l, diffuse_transmission_pdf);
        contribution *= w.diffuse_transmission;
        contribution *= layers_throughput;
        pdf += diffuse_transmission_pdf * p.diffuse_transmission;
        final_color += contribution;
    }

    final_color /= energy_compensation_clearcoat_lobe(rd, m, incident_medium_ior, f.v.z, ctx.current_bounce);
    return final_color;
}

B2_HD_NOINLINE float principled_bsdf_pdf(const b2pt_render_data& rd, BsdfCtx& ctx)
{
    float pdf = 0.0f;
    const Mat& m = *ctx.material;
    VolumeState& vs = *ctx.volume_state;

    bool outside_object = !vs.inside_material;
    bool refracting = dot(ctx.shading_normal, ctx.to_light_direction) < 0.0f && outside_object;
    PrincipledFrames f = principled_frames(ctx);

    LobeWeights w = principled_lobe_weights(m, outside_object);
    LobeWeights p = principled_lobe_probabilities(w);
    float not_refracting = refracting ? 0.0f : 1.0f;

    pdf += internal_pdf_coat_layer(rd, ctx, f.v, f.l, f.h, refracting, w.coat, p.coat);
    pdf += internal_pdf_sheen_layer(rd, m, f.v, f.l, refracting, w.sheen, p.sheen);
    pdf += internal_pdf_metal_layer(rd, ctx, m.roughness, m.anisotropy, f.vr, f.lr, f.hr, w.metal_1 * not_refracting, p.metal_1);
    pdf += internal_pdf_metal_layer(rd, ctx, m.second_roughness, m.anisotropy, f.vr, f.lr, f.hr, w.metal_2 * not_refracting, p.metal_2);
    if (w.glass > 0.0f)
        pdf += principled_glass_pdf(rd, ctx, f.vr, f.lr) * p.glass;
    {
        float glossy = 0.0f;
        glossy += internal_pdf_specular_layer(rd, ctx, f.vr, f.lr, f.hr, w.specular, refracting, p.specular);
        float diffuse_weight = w.diffuse * not_refracting;
        float diffuse = 0.0f;
        if (diffuse_weight > 0.0f && f.v.z > 0.0f && f.l.z > 0.0f)
            diffuse = lambertian_pdf(f.l.z) * p.diffuse;
        glossy += diffuse;
        pdf += glossy;
    }
    if (w.diffuse_transmission > 0.0f && f.l.z < 0.0f)
    {
        float dt = 0.0f;
        if (!(f.v.z * f.l.z > 0.0f))
            dt = fabsf(f.l.z * B2_INV_PI);
        pdf += dt * p.diffuse_transmission;
    }
    return pdf;
}

B2_HD_NOINLINE Rgb principled_bsdf_sample(const b2pt_render_data& rd, BsdfCtx& ctx, V3& output_direction, float& pdf, Rng& rng)
{
    pdf = 0.0f;
    const Mat& m = *ctx.material;
    VolumeState& vs = *ctx.volume_state;

    bool is_outside_object = !vs.inside_material;
    LobeWeights w = principled_lobe_weights(m, is_outside_object);
    LobeWeights p = principled_lobe_probabilities(w);

    float cdf0 = p.coat;
    float cdf1 = cdf0 + p.sheen;
    float cdf2 = cdf1 + p.metal_1;
    float cdf3 = cdf2 + p.metal_2;
    float cdf4 = cdf3 + p.specular;
    float cdf5 = cdf4 + p.diffuse;
    float cdf6 = cdf5 + p.diffuse_transmission;

    float rand_1 = rng.next_float();
    bool sampling_diffuse_transmission_lobe = rand_1 > cdf5 && rand_1 < cdf6;
    bool sampling_glass_lobe = rand_1 > cdf6;

    if (ctx.update_ray_volume_state)
        if (!sampling_glass_lobe && !sampling_diffuse_transmission_lobe)
            vs.interior_stack.pop(false);

    V3 TR, BR;
    build_rotated_ONB(ctx.shading_normal, TR, BR, m.anisotropy_rotation * B2_PI);
    V3 local_view_direction_rotated = world_to_local_frame(TR, BR, ctx.shading_normal, ctx.view_direction);

    if (rand_1 < cdf0)
    {
        V3 TR_coat, BR_coat;
        build_rotated_ONB(ctx.shading_normal, TR_coat, BR_coat, m.coat_anisotropy_rotation * B2_PI);
        V3 local_view_direction_rotated_coat = world_to_local_frame(TR_coat, BR_coat, ctx.shading_normal, ctx.view_direction);
        *ctx.incident_light_info = B2_LOBE_COAT;
        float regularized_roughness = regularize_reflection(rd, ctx.regularization_mode, m.coat_roughness, ctx.accumulated_path_roughness);
        output_direction = local_to_world_frame(TR_coat, BR_coat, ctx.shading_normal, microfacet_GGX_sample_reflection(regularized_roughness, m.coat_anisotropy, local_view_direction_rotated_coat, rng));
    }
    else if (rand_1 < cdf1)
    {
        V3 T, B;
        build_ONB(ctx.shading_normal, T, B);
        V3 local_view_direction = world_to_local_frame(T, B, ctx.shading_normal, ctx.view_direction);
        output_direction = local_to_world_frame(T, B, ctx.shading_normal, sheen_ltc_sample(rd, m, local_view_direction, rng));
    }
    else if (rand_1 < cdf2)
    {
        *ctx.incident_light_info = B2_LOBE_FIRST_METAL;
        float regularized_roughness = regularize_reflection(rd, ctx.regularization_mode, m.roughness, ctx.accumulated_path_roughness);
        output_direction = local_to_world_frame(TR, BR, ctx.shading_normal, microfacet_GGX_sample_reflection(regularized_roughness, m.anisotropy, local_view_direction_rotated, rng));
    }
    else if (rand_1 < cdf3)
    {
        *ctx.incident_light_info = B2_LOBE_SECOND_METAL;
        float regularized_roughness = regularize_reflection(rd, ctx.regularization_mode, m.second_roughness, ctx.accumulated_path_roughness);
        output_direction = local_to_world_frame(TR, BR, ctx.shading_normal, microfacet_GGX_sample_reflection(regularized_roughness, m.anisotropy, local_view_direction_rotated, rng));
    }
    else if (rand_1 < cdf4)
    {
        *ctx.incident_light_info = B2_LOBE_SPECULAR;
        float regularized_roughness = regularize_reflection(rd, ctx.regularization_mode, m.roughness, ctx.accumulated_path_roughness);
        output_direction = local_to_world_frame(TR, BR, ctx.shading_normal, microfacet_GGX_sample_reflection(regularized_roughness, m.anisotropy, local_view_direction_rotated, rng));
    }
    else if (rand_1 < cdf5)
    {
        *ctx.incident_light_info = B2_LOBE_DIFFUSE;
        output_direction = cosine_weighted_sample_around_normal_world_space(ctx.shading_normal, rng);
    }
    else if (rand_1 < cdf6)
    {
        *ctx.incident_light_info = B2_LOBE_DIFFUSE_TRANSMISSION;
        output_direction = cosine_weighted_sample_around_normal_world_space(-ctx.shading_normal, rng);
    }
    else
        output_direction = local_to_world_frame(TR, BR, ctx.shading_normal, principled_glass_sample(rd, ctx, local_view_direction_rotated, rng));

    if (dot(output_direction, ctx.geometric_normal) < 0.0f && !sampling_glass_lobe && !sampling_diffuse_transmission_lobe)
        return rgb(0.0f);

    ctx.to_light_direction = output_direction;
    return principled_bsdf_eval(rd, ctx, pdf);
}

#endif
