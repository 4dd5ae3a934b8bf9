/*
 * b2_material.h -- material records of the hot path: the unpacked working form, the SoA scene form it is
 * gathered from, the 108-byte packed G-buffer form, and the 24-bit octahedral normal.
 *
 * Data formats are the reference's (they are what its host uploads): 8-bit quantised fields and bit positions
 * from /root/reference/src/HostDeviceCommon/Packing.h:16-238, record layout from
 * HostDeviceCommon/Material/MaterialPacked.h:500-539, SoA from MaterialPackedSoA.h:85-120. Quantisation is
 * `(uint8)(v * 255.0f)` (truncation) on pack and `k * (1/255)` on unpack; because a G-buffer round trip re-packs
 * unpacked values, the exact float expressions are kept so re-quantisation drifts exactly like the reference.
 */
#ifndef B2_MATERIAL_H
#define B2_MATERIAL_H

#include "b2_math.h"
#include "b2_bvh.h"
#include "b2pt_render_data.h"

#define B2_NO_TEXTURE 65535
#define B2_ROUGHNESS_CLAMP 1.0e-4f
#define B2_PERFECTLY_SMOOTH_ROUGHNESS_THRESHOLD 1.0e-2f
#define B2_DELTA_DISTRIBUTION_HIGH_VALUE 1.0e9f
#define B2_DELTA_DISTRIBUTION_ALIGNEMENT_THRESHOLD 0.999999f

/* Working form of a surface material (reference: DeviceUnpackedEffectiveMaterial, MaterialUnpacked.h:11-106) */
struct Mat
{
    Rgb emission;
    Rgb base_color;
    float roughness;
    float oren_nayar_sigma;
    float metallic;
    float metallic_F90_falloff_exponent;
    Rgb metallic_F82;
    Rgb metallic_F90;
    float anisotropy;
    float anisotropy_rotation;
    float second_roughness_weight;
    float second_roughness;
    float specular;
    float specular_tint;
    Rgb specular_color;
    float specular_darkening;
    float coat;
    Rgb coat_medium_absorption;
    float coat_medium_thickness;
    float coat_roughness;
    float coat_roughening;
    float coat_darkening;
    float coat_anisotropy;
    float coat_anisotropy_rotation;
    float coat_ior;
    float sheen;
    float sheen_roughness;
    Rgb sheen_color;
    float ior;
    float specular_transmission;
    float diffuse_transmission;
    float absorption_at_distance;
    Rgb absorption_color;
    float dispersion_scale;
    float dispersion_abbe_number;
    float thin_film;
    float thin_film_ior;
    float thin_film_thickness;
    float thin_film_kappa_3;
    float thin_film_hue_shift_degrees;
    float thin_film_base_ior_override;
    float alpha_opacity;
    uint8_t energy_preservation_monte_carlo_samples;
    uint8_t dielectric_priority;
    bool do_metallic_energy_compensation;
    bool do_specular_energy_compensation;
    bool do_coat_energy_compensation;
    bool thin_walled;
    bool do_glass_energy_compensation;
    bool thin_film_do_ior_override;
    bool enforce_strong_energy_conservation;
    bool emissive_texture_used;
};

B2_HD Mat mat_defaults()
{
    Mat m;
    m.emission = rgb(0.0f);
    m.base_color = rgb(1.0f);
    m.roughness = 0.3f;
    m.oren_nayar_sigma = 0.34906585039886591538f;
    m.metallic = 0.0f;
    m.metallic_F90_falloff_exponent = 5.0f;
    m.metallic_F82 = rgb(1.0f);
    m.metallic_F90 = rgb(1.0f);
    m.anisotropy = 0.0f;
    m.anisotropy_rotation = 0.0f;
    m.second_roughness_weight = 0.0f;
    m.second_roughness = 0.5f;
    m.specular = 1.0f;
    m.specular_tint = 1.0f;
    m.specular_color = rgb(1.0f);
    m.specular_darkening = 0.0f;
    m.coat = 0.0f;
    m.coat_medium_absorption = rgb(1.0f);
    m.coat_medium_thickness = 5.0f;
    m.coat_roughness = 0.0f;
    m.coat_roughening = 1.0f;
    m.coat_darkening = 1.0f;
    m.coat_anisotropy = 0.0f;
    m.coat_anisotropy_rotation = 0.0f;
    m.coat_ior = 1.5f;
    m.sheen = 0.0f;
    m.sheen_roughness = 0.5f;
    m.sheen_color = rgb(1.0f);
    m.ior = 1.40f;
    m.specular_transmission = 0.0f;
    m.diffuse_transmission = 0.0f;
    m.absorption_at_distance = 1.0f;
    m.absorption_color = rgb(1.0f);
    m.dispersion_scale = 0.0f;
    m.dispersion_abbe_number = 20.0f;
    m.thin_film = 0.0f;
    m.thin_film_ior = 1.3f;
    m.thin_film_thickness = 500.0f;
    m.thin_film_kappa_3 = 0.0f;
    m.thin_film_hue_shift_degrees = 0.0f;
    m.thin_film_base_ior_override = 1.0f;
    m.alpha_opacity = 1.0f;
    m.energy_preservation_monte_carlo_samples = 12;
    m.dielectric_priority = 0;
    m.do_metallic_energy_compensation = true;
    m.do_specular_energy_compensation = true;
    m.do_coat_energy_compensation = true;
    m.thin_walled = false;
    m.do_glass_energy_compensation = true;
    m.thin_film_do_ior_override = false;
    m.enforce_strong_energy_conservation = false;
    m.emissive_texture_used = false;
    return m;
}

B2_HD bool mat_is_emissive(const Mat& m)
{
    return !b2_is_zero(m.emission.r) || !b2_is_zero(m.emission.g) || !b2_is_zero(m.emission.b) || m.emissive_texture_used;
}

/* ---- 8-bit field codecs (Packing.h:38-157) ---- */
#define B2_INV_255 (1.0f / 255.0f)
#define B2_INV_255_SHL_8 (1.0f / (float)(255 << 8))

B2_HD Rgb unpack_rgb24(uint32_t w)
{
    float r = (float)(w & 0x000000FFu) * B2_INV_255;
    float g = (float)(w & 0x0000FF00u) * B2_INV_255_SHL_8;
    float b = (float)((w & 0x00FF0000u) >> 16) * B2_INV_255;
    return rgb(r, g, b);
}
B2_HD float unpack_hi8(uint32_t w) { return (float)((w & 0xFF000000u) >> 24) * B2_INV_255; }
B2_HD float unpack_f8(uint32_t w, int index) { return (float)((w >> (index * 8)) & 0xFFu) * 0.00392156862745098039f; }
B2_HD uint32_t quant8(float v) { return (uint32_t)(uint8_t)(v * 255.0f); }
B2_HD uint32_t pack_rgb24_f8(Rgb c, float f) { return quant8(c.r) | (quant8(c.g) << 8) | (quant8(c.b) << 16) | (quant8(f) << 24); }
B2_HD uint32_t pack_f8x4(float a, float b, float c, float d) { return quant8(a) | (quant8(b) << 8) | (quant8(c) << 16) | (quant8(d) << 24); }

/* flag bit positions (MaterialPacked.h:18-28) */
enum
{
    B2_FLAG_THIN_WALLED = 0,
    B2_FLAG_EMISSIVE_TEXTURE_USED = 1,
    B2_FLAG_THIN_FILM_DO_IOR_OVERRIDE = 2,
    B2_FLAG_GLASS_ENERGY_COMPENSATION = 3,
    B2_FLAG_COAT_ENERGY_COMPENSATION = 4,
    B2_FLAG_METALLIC_ENERGY_COMPENSATION = 5,
    B2_FLAG_SPECULAR_ENERGY_COMPENSATION = 6,
    B2_FLAG_ENFORCE_STRONG_ENERGY_CONSERVATION = 7
};

/* The 108-byte packed G-buffer record, word by word (27 words). Reference member order, MaterialPacked.h:500-539. */
enum
{
    PM_FLAGS = 0,            /* byte 0 only */
    PM_EMISSION_R = 1, PM_EMISSION_G = 2, PM_EMISSION_B = 3,
    PM_BASE_COLOR_ROUGHNESS = 4,
    PM_OREN_SIGMA = 5,
    PM_METALLIC_F90_AND_METALLIC = 6,
    PM_METALLIC_F82_AND_DIFFUSE_TRANSMISSION = 7,
    PM_METALLIC_F90_FALLOFF = 8,
    PM_ANISO_ROT_SECOND_ROUGHNESS = 9,
    PM_SPECULAR_COLOR_TINT = 10,
    PM_SPECULAR_DARKENING_COAT_ROUGHNESS = 11,
    PM_COAT_MEDIUM_THICKNESS = 12,
    PM_COAT_AND_MEDIUM_ABSORPTION = 13,
    PM_COAT_ROUGHENING_DARKENING_ANISO = 14,
    PM_COAT_IOR = 15,
    PM_SHEEN_AND_COLOR = 16,
    PM_IOR = 17,
    PM_ABSORPTION_COLOR = 18,
    PM_ABSORPTION_DISTANCE = 19,
    PM_SHEEN_ROUGHNESS_TRANSMISSION_DISPERSION_THIN_FILM = 20,
    PM_DISPERSION_ABBE = 21,
    PM_THIN_FILM_IOR = 22,
    PM_THIN_FILM_THICKNESS = 23,
    PM_THIN_FILM_KAPPA_3 = 24,
    PM_THIN_FILM_BASE_IOR_OVERRIDE = 25,
    PM_ALPHA_HUE_PRIORITY = 26,
    PM_WORDS = 27
};

struct PackedMat { uint32_t w[PM_WORDS]; };

/* reference: DevicePackedEffectiveMaterial::pack, MaterialPacked.h:87-155 */
template <int BSDF_OVERRIDE>
B2_HD PackedMat pack_material(const Mat& m)
{
    PackedMat p;
    uint32_t flags = 0;
    flags |= (m.thin_walled ? 1u : 0u) << B2_FLAG_THIN_WALLED;
    flags |= (m.emissive_texture_used ? 1u : 0u) << B2_FLAG_EMISSIVE_TEXTURE_USED;
    flags |= (m.thin_film_do_ior_override ? 1u : 0u) << B2_FLAG_THIN_FILM_DO_IOR_OVERRIDE;
    flags |= (m.do_glass_energy_compensation ? 1u : 0u) << B2_FLAG_GLASS_ENERGY_COMPENSATION;
    flags |= (m.do_coat_energy_compensation ? 1u : 0u) << B2_FLAG_COAT_ENERGY_COMPENSATION;
    flags |= (m.do_metallic_energy_compensation ? 1u : 0u) << B2_FLAG_METALLIC_ENERGY_COMPENSATION;
    flags |= (m.do_specular_energy_compensation ? 1u : 0u) << B2_FLAG_SPECULAR_ENERGY_COMPENSATION;
    flags |= (m.enforce_strong_energy_conservation ? 1u : 0u) << B2_FLAG_ENFORCE_STRONG_ENERGY_CONSERVATION;
    p.w[PM_FLAGS] = flags;
    p.w[PM_EMISSION_R] = b2_float_as_uint(m.emission.r);
    p.w[PM_EMISSION_G] = b2_float_as_uint(m.emission.g);
    p.w[PM_EMISSION_B] = b2_float_as_uint(m.emission.b);
    p.w[PM_BASE_COLOR_ROUGHNESS] = pack_rgb24_f8(m.base_color, m.roughness);
    p.w[PM_OREN_SIGMA] = b2_float_as_uint(m.oren_nayar_sigma);
    p.w[PM_METALLIC_F90_AND_METALLIC] = pack_rgb24_f8(m.metallic_F90, m.metallic);
    p.w[PM_METALLIC_F82_AND_DIFFUSE_TRANSMISSION] = pack_rgb24_f8(m.metallic_F82, m.diffuse_transmission);
    p.w[PM_METALLIC_F90_FALLOFF] = b2_float_as_uint(m.metallic_F90_falloff_exponent);
    p.w[PM_ANISO_ROT_SECOND_ROUGHNESS] = pack_f8x4(m.anisotropy, m.anisotropy_rotation, m.second_roughness_weight, m.second_roughness);
    p.w[PM_SPECULAR_COLOR_TINT] = pack_rgb24_f8(m.specular_color, m.specular_tint);
    p.w[PM_SPECULAR_DARKENING_COAT_ROUGHNESS] = quant8(m.specular) | (quant8(m.specular_darkening) << 8) | (quant8(m.coat_roughness) << 16);
    p.w[PM_COAT_MEDIUM_THICKNESS] = b2_float_as_uint(m.coat_medium_thickness);
    p.w[PM_COAT_AND_MEDIUM_ABSORPTION] = pack_rgb24_f8(m.coat_medium_absorption, m.coat);
    p.w[PM_COAT_ROUGHENING_DARKENING_ANISO] = pack_f8x4(m.coat_roughening, m.coat_darkening, m.coat_anisotropy, m.coat_anisotropy_rotation);
    p.w[PM_COAT_IOR] = b2_float_as_uint(m.coat_ior);
    p.w[PM_SHEEN_AND_COLOR] = pack_rgb24_f8(m.sheen_color, m.sheen);
    p.w[PM_IOR] = b2_float_as_uint(m.ior);
    p.w[PM_ABSORPTION_COLOR] = pack_rgb24_f8(m.absorption_color, 0.0f);
    p.w[PM_ABSORPTION_DISTANCE] = b2_float_as_uint(m.absorption_at_distance);
    p.w[PM_SHEEN_ROUGHNESS_TRANSMISSION_DISPERSION_THIN_FILM] = pack_f8x4(m.sheen_roughness, m.specular_transmission, m.dispersion_scale, m.thin_film);
    p.w[PM_DISPERSION_ABBE] = b2_float_as_uint(m.dispersion_abbe_number);
    p.w[PM_THIN_FILM_IOR] = b2_float_as_uint(m.thin_film_ior);
    p.w[PM_THIN_FILM_THICKNESS] = b2_float_as_uint(m.thin_film_thickness);
    p.w[PM_THIN_FILM_KAPPA_3] = b2_float_as_uint(m.thin_film_kappa_3);
    /* reference quirk kept: the setter takes a bool, so the stored float is 0 or 1 (MaterialPacked.h:485) */
    p.w[PM_THIN_FILM_BASE_IOR_OVERRIDE] = b2_float_as_uint(m.thin_film_base_ior_override != 0.0f ? 1.0f : 0.0f);
    /* get_dielectric_priority() is 0 under the Lambertian / Oren-Nayar overrides (MaterialUnpacked.h:92-99) */
    uint32_t priority = (BSDF_OVERRIDE == 1 || BSDF_OVERRIDE == 2) ? 0u : (uint32_t)m.dielectric_priority;
    p.w[PM_ALPHA_HUE_PRIORITY] = quant8(m.alpha_opacity) | (quant8(m.thin_film_hue_shift_degrees) << 8) | (priority << 16) | ((uint32_t)m.energy_preservation_monte_carlo_samples << 24);
    return p;
}

/* reference: DevicePackedEffectiveMaterial::unpack, MaterialPacked.h:157-225 (unconditional: every field is read) */
B2_HD Mat unpack_material(const PackedMat& p)
{
    Mat m;
    uint32_t flags = p.w[PM_FLAGS];
    m.emission = rgb(b2_uint_as_float(p.w[PM_EMISSION_R]), b2_uint_as_float(p.w[PM_EMISSION_G]), b2_uint_as_float(p.w[PM_EMISSION_B]));
    m.emissive_texture_used = (flags >> B2_FLAG_EMISSIVE_TEXTURE_USED) & 1u;
    m.base_color = unpack_rgb24(p.w[PM_BASE_COLOR_ROUGHNESS]);
    m.roughness = unpack_hi8(p.w[PM_BASE_COLOR_ROUGHNESS]);
    m.oren_nayar_sigma = b2_uint_as_float(p.w[PM_OREN_SIGMA]);
    m.metallic = unpack_hi8(p.w[PM_METALLIC_F90_AND_METALLIC]);
    m.metallic_F90_falloff_exponent = b2_uint_as_float(p.w[PM_METALLIC_F90_FALLOFF]);
    m.metallic_F82 = unpack_rgb24(p.w[PM_METALLIC_F82_AND_DIFFUSE_TRANSMISSION]);
    m.metallic_F90 = unpack_rgb24(p.w[PM_METALLIC_F90_AND_METALLIC]);
    m.anisotropy = unpack_f8(p.w[PM_ANISO_ROT_SECOND_ROUGHNESS], 0);
    m.anisotropy_rotation = unpack_f8(p.w[PM_ANISO_ROT_SECOND_ROUGHNESS], 1);
    m.second_roughness_weight = unpack_f8(p.w[PM_ANISO_ROT_SECOND_ROUGHNESS], 2);
    m.second_roughness = unpack_f8(p.w[PM_ANISO_ROT_SECOND_ROUGHNESS], 3);
    m.do_metallic_energy_compensation = (flags >> B2_FLAG_METALLIC_ENERGY_COMPENSATION) & 1u;
    m.specular = unpack_f8(p.w[PM_SPECULAR_DARKENING_COAT_ROUGHNESS], 0);
    m.specular_tint = unpack_hi8(p.w[PM_SPECULAR_COLOR_TINT]);
    m.specular_color = unpack_rgb24(p.w[PM_SPECULAR_COLOR_TINT]);
    m.specular_darkening = unpack_f8(p.w[PM_SPECULAR_DARKENING_COAT_ROUGHNESS], 1);
    m.do_specular_energy_compensation = (flags >> B2_FLAG_SPECULAR_ENERGY_COMPENSATION) & 1u;
    m.coat = unpack_hi8(p.w[PM_COAT_AND_MEDIUM_ABSORPTION]);
    m.coat_medium_absorption = unpack_rgb24(p.w[PM_COAT_AND_MEDIUM_ABSORPTION]);
    m.coat_medium_thickness = b2_uint_as_float(p.w[PM_COAT_MEDIUM_THICKNESS]);
    m.coat_roughness = unpack_f8(p.w[PM_SPECULAR_DARKENING_COAT_ROUGHNESS], 2);
    m.coat_roughening = unpack_f8(p.w[PM_COAT_ROUGHENING_DARKENING_ANISO], 0);
    m.coat_darkening = unpack_f8(p.w[PM_COAT_ROUGHENING_DARKENING_ANISO], 1);
    m.coat_anisotropy = unpack_f8(p.w[PM_COAT_ROUGHENING_DARKENING_ANISO], 2);
    m.coat_anisotropy_rotation = unpack_f8(p.w[PM_COAT_ROUGHENING_DARKENING_ANISO], 3);
    m.coat_ior = b2_uint_as_float(p.w[PM_COAT_IOR]);
    m.do_coat_energy_compensation = (flags >> B2_FLAG_COAT_ENERGY_COMPENSATION) & 1u;
    m.sheen = unpack_hi8(p.w[PM_SHEEN_AND_COLOR]);
    m.sheen_roughness = unpack_f8(p.w[PM_SHEEN_ROUGHNESS_TRANSMISSION_DISPERSION_THIN_FILM], 0);
    m.sheen_color = unpack_rgb24(p.w[PM_SHEEN_AND_COLOR]);
    m.ior = b2_uint_as_float(p.w[PM_IOR]);
    m.specular_transmission = unpack_f8(p.w[PM_SHEEN_ROUGHNESS_TRANSMISSION_DISPERSION_THIN_FILM], 1);
    m.diffuse_transmission = unpack_hi8(p.w[PM_METALLIC_F82_AND_DIFFUSE_TRANSMISSION]);
    m.absorption_at_distance = b2_uint_as_float(p.w[PM_ABSORPTION_DISTANCE]);
    m.absorption_color = unpack_rgb24(p.w[PM_ABSORPTION_COLOR]);
    m.dispersion_scale = unpack_f8(p.w[PM_SHEEN_ROUGHNESS_TRANSMISSION_DISPERSION_THIN_FILM], 2);
    m.dispersion_abbe_number = b2_uint_as_float(p.w[PM_DISPERSION_ABBE]);
    m.thin_walled = (flags >> B2_FLAG_THIN_WALLED) & 1u;
    m.do_glass_energy_compensation = (flags >> B2_FLAG_GLASS_ENERGY_COMPENSATION) & 1u;
    m.thin_film = unpack_f8(p.w[PM_SHEEN_ROUGHNESS_TRANSMISSION_DISPERSION_THIN_FILM], 3);
    m.thin_film_ior = b2_uint_as_float(p.w[PM_THIN_FILM_IOR]);
    m.thin_film_thickness = b2_uint_as_float(p.w[PM_THIN_FILM_THICKNESS]);
    m.thin_film_kappa_3 = b2_uint_as_float(p.w[PM_THIN_FILM_KAPPA_3]);
    m.thin_film_hue_shift_degrees = unpack_f8(p.w[PM_ALPHA_HUE_PRIORITY], 1);
    m.thin_film_base_ior_override = b2_uint_as_float(p.w[PM_THIN_FILM_BASE_IOR_OVERRIDE]);
    m.thin_film_do_ior_override = (flags >> B2_FLAG_THIN_FILM_DO_IOR_OVERRIDE) & 1u;
    m.alpha_opacity = unpack_f8(p.w[PM_ALPHA_HUE_PRIORITY], 0);
    m.dielectric_priority = (uint8_t)((p.w[PM_ALPHA_HUE_PRIORITY] >> 16) & 0xFFu);
    m.energy_preservation_monte_carlo_samples = (uint8_t)((p.w[PM_ALPHA_HUE_PRIORITY] >> 24) & 0xFFu);
    m.enforce_strong_energy_conservation = (flags >> B2_FLAG_ENFORCE_STRONG_ENERGY_CONSERVATION) & 1u;
    return m;
}

/*
 * Scene SoA -> working material for an UNTEXTURED material. Composition of the reference's
 * read_partial_material (MaterialPackedSoA.h:381-551, lobe-conditional loads into a default-initialised packed
 * record), DevicePackedTexturedMaterial::unpack (MaterialPacked.h:541-668, same conditions, into a
 * default-initialised unpacked record) and get_intersection_material's coat roughening (Material.h:60-150).
 * Because both steps test the same predicates on the same quantised values, a field is either read from the SoA
 * or left at the UNPACKED default; fields the second step reads unconditionally but the first step skipped come
 * out as the PACKED default (all-zero words), which only concerns textured materials.
 * Textured materials (texture indices != 65535) are a "next" row: the kernels refuse scenes that contain them.
 */
B2_HD Mat load_scene_material(const b2pt_materials_soa& s, int i, bool white_furnace)
{
    Mat m = mat_defaults();
    uint32_t flags = s.flags[i];

    m.emissive_texture_used = (flags >> B2_FLAG_EMISSIVE_TEXTURE_USED) & 1u;
    if (!m.emissive_texture_used)
        m.emission = rgb(s.emission[i].r, s.emission[i].g, s.emission[i].b);

    uint32_t bcr = s.base_color_roughness[i];
    m.base_color = unpack_rgb24(bcr);
    m.roughness = unpack_hi8(bcr);
    m.oren_nayar_sigma = s.oren_nayar_sigma[i];

    uint32_t mf90 = s.metallic_F90_and_metallic[i];
    m.metallic = unpack_hi8(mf90);
    uint32_t aniso = s.anisotropy_and_rotation_and_second_roughness[i];
    uint32_t mf82 = s.metallic_F82_packed_and_diffuse_transmission[i];
    if (m.metallic > 0.0f)
    {
        m.metallic_F90_falloff_exponent = s.metallic_F90_falloff_exponent[i];
        m.metallic_F82 = unpack_rgb24(mf82);
        m.metallic_F90 = unpack_rgb24(mf90);
        m.second_roughness_weight = unpack_f8(aniso, 2);
        m.second_roughness = unpack_f8(aniso, 3);
        m.do_metallic_energy_compensation = (flags >> B2_FLAG_METALLIC_ENERGY_COMPENSATION) & 1u;
    }
    m.anisotropy = unpack_f8(aniso, 0);
    m.anisotropy_rotation = unpack_f8(aniso, 1);

    uint32_t spec = s.specular_and_darkening_and_coat_roughness[i];
    m.specular = unpack_f8(spec, 0);
    if (m.specular > 0.0f)
    {
        uint32_t sct = s.specular_color_and_tint_factor[i];
        m.specular_tint = unpack_hi8(sct);
        m.specular_color = unpack_rgb24(sct);
        m.specular_darkening = unpack_f8(spec, 1);
        m.do_specular_energy_compensation = (flags >> B2_FLAG_SPECULAR_ENERGY_COMPENSATION) & 1u;
    }

    uint32_t coat_w = s.coat_and_medium_absorption[i];
    m.coat = unpack_hi8(coat_w);
    if (m.coat > 0.0f)
    {
        uint32_t crd = s.coat_roughening_darkening_anisotropy_and_rotation[i];
        m.coat_medium_absorption = unpack_rgb24(coat_w);
        m.coat_medium_thickness = s.coat_medium_thickness[i];
        m.coat_roughness = unpack_f8(spec, 2);
        m.coat_roughening = unpack_f8(crd, 0);
        m.coat_darkening = unpack_f8(crd, 1);
        m.coat_anisotropy = unpack_f8(crd, 2);
        m.coat_anisotropy_rotation = unpack_f8(crd, 3);
        m.coat_ior = s.coat_ior[i];
        m.do_coat_energy_compensation = (flags >> B2_FLAG_COAT_ENERGY_COMPENSATION) & 1u;
    }

    uint32_t sheen_w = s.sheen_and_color[i];
    uint32_t srtd = s.sheen_roughness_transmission_dispersion_thin_film[i];
    m.sheen = unpack_hi8(sheen_w);
    if (m.sheen > 0.0f)
    {
        m.sheen_roughness = unpack_f8(srtd, 0);
        m.sheen_color = unpack_rgb24(sheen_w);
    }

    m.ior = s.ior[i];
    m.diffuse_transmission = unpack_hi8(mf82);
    m.specular_transmission = unpack_f8(srtd, 1);
    if (m.specular_transmission > 0.0f)
    {
        m.dispersion_scale = unpack_f8(srtd, 2);
        m.dispersion_abbe_number = s.dispersion_abbe_number[i];
        m.thin_walled = (flags >> B2_FLAG_THIN_WALLED) & 1u;
        m.do_glass_energy_compensation = (flags >> B2_FLAG_GLASS_ENERGY_COMPENSATION) & 1u;
    }
    if (m.specular_transmission > 0.0f || m.diffuse_transmission > 0.0f)
    {
        m.absorption_at_distance = s.absorption_at_distance[i];
        m.absorption_color = unpack_rgb24(s.absorption_color_packed[i]);
    }

    uint32_t ahp = s.alpha_thin_film_hue_dielectric_priority[i];
    m.thin_film = unpack_f8(srtd, 3);
    if (m.thin_film > 0.0f)
    {
        m.thin_film_ior = s.thin_film_ior[i];
        m.thin_film_thickness = s.thin_film_thickness[i];
        m.thin_film_kappa_3 = s.thin_film_kappa_3[i];
        m.thin_film_hue_shift_degrees = unpack_f8(ahp, 1);
        /* read_partial_material funnels this float through a bool setter (MaterialPacked.h:485) */
        m.thin_film_base_ior_override = s.thin_film_base_ior_override[i] != 0.0f ? 1.0f : 0.0f;
        m.thin_film_do_ior_override = (flags >> B2_FLAG_THIN_FILM_DO_IOR_OVERRIDE) & 1u;
    }

    m.alpha_opacity = unpack_f8(ahp, 0);
    m.dielectric_priority = (uint8_t)((ahp >> 16) & 0xFFu);
    /* read_partial_material never copies the strong-energy-conservation flag: stays false, samples stay 12 */
    /* get_intersection_material: emissive_texture_used = (emission texture index != NO_TEXTURE), Material.h:122 */
    m.emissive_texture_used = false;

    if (white_furnace)
        m.base_color = rgb(1.0f);

    /* coat roughening of the base layer(s), Material.h:127-147 */
    if (m.coat > 0.0f && m.coat_roughening > 0.0f)
    {
        float base_roughness = m.roughness;
        float coat_roughness = m.coat_roughness;
        float target = b2_pow_1_4(b2_min(1.0f, b2_pow_4(base_roughness) + 2.0f * b2_pow_4(coat_roughness)));
        float roughened = b2_lerp(base_roughness, target, m.coat);
        m.roughness = b2_lerp(base_roughness, roughened, m.coat_roughening);
        if (m.second_roughness_weight > 0.0f)
        {
            float second = m.second_roughness;
            float target2 = b2_pow_1_4(b2_min(1.0f, b2_pow_4(second) + 2.0f * b2_pow_4(coat_roughness)));
            float roughened2 = b2_lerp(second, target2, m.coat);
            m.second_roughness = b2_lerp(second, roughened2, m.coat_roughening);
        }
    }
    return m;
}

/* ---- material textures --------------------------------------------------------------------------------------------------------
 * What RenderBuffers::material_textures points to on this device layer: an array of these descriptors (the reference's array of
 * oroTextureObject_t on its GPU path, of Image8Bit on its CPU path). 8-bit texels, 1 or 4 channels, row 0 first; the descriptor is
 * what b2pt_texture_create returns. */
/* struct b2pt_texture: b2_bvh.h (the traversal samples the base-colour alpha for the alpha test) */

/* Image8Bit::sample_rgba32f (src/Image/Image.cpp:161-195): nearest texel, repeat addressing, v flipped; channels beyond the
 * image's keep ColorRGBA32F's defaults (0, 0, 0, 1). Then sample_texture_rgba's sRGB decode of ALL four channels. */
B2_HD void sample_texture_rgba(const void* textures, int texture_index, bool is_srgb, V2 uv, float (&rgba)[4])
{
    const b2pt_texture& t = reinterpret_cast<const b2pt_texture*>(textures)[texture_index];
    float u = uv.x;
    if (u != 1.0f)
        u -= (float)(int)uv.x;
    float v = uv.y;
    if (v != 1.0f)
        v -= (float)(int)uv.y;
    u = u < 0.0f ? 1.0f + u : u;
    v = v < 0.0f ? 1.0f + v : v;
    v = 1.0f - v;
    int x = (int)(u * (float)(t.width - 1));
    int y = (int)(v * (float)(t.height - 1));
    rgba[0] = rgba[1] = rgba[2] = 0.0f;
    rgba[3] = 1.0f;
    const uint8_t* texel = t.texels + ((size_t)x + (size_t)y * (size_t)t.width) * (size_t)t.channels;
    for (int c = 0; c < t.channels; c++)
        rgba[c] = (float)texel[c] / 255.0f;
    if (is_srgb)
    {
        /* intrin_pow(rgba, 2.2f): on texel channels a table lookup; on defaulted channels pow(0) = 0 and pow(1) = 1 */
        for (int c = 0; c < t.channels; c++)
            rgba[c] = t.srgb_to_linear[texel[c]];
    }
}

/* texture indices of a material: the six packed words of DevicePackedTexturedMaterialSoA and the reference's getters
 * (MaterialPacked.h:713-723) -- including get_anisotropic_texture_index reading the ROUGHNESS slot of the roughness/metallic word */
struct MatTextureIndices
{
    int normal_map, emission, base_color, roughness_metallic, roughness, metallic, anisotropic, specular, coat, sheen, specular_transmission;
};
B2_HD MatTextureIndices load_texture_indices(const b2pt_materials_soa& s, int i)
{
    MatTextureIndices t;
    const uint32_t w0 = s.normal_map_emission_index[i], w1 = s.base_color_roughness_metallic_index[i], w2 = s.roughness_and_metallic_index[i];
    const uint32_t w3 = s.anisotropic_specular_index[i], w4 = s.coat_sheen_index[i], w5 = s.specular_transmission_index[i];
    t.normal_map = (int)(w0 & 0xFFFFu);
    t.emission = (int)(w0 >> 16);
    t.base_color = (int)(w1 & 0xFFFFu);
    t.roughness_metallic = (int)(w1 >> 16);
    t.roughness = (int)(w2 & 0xFFFFu);
    t.metallic = (int)(w2 >> 16);
    t.anisotropic = (int)(w2 & 0xFFFFu);
    t.specular = (int)(w3 >> 16);
    t.coat = (int)(w4 & 0xFFFFu);
    t.sheen = (int)(w4 >> 16);
    t.specular_transmission = (int)(w5 & 0xFFFFu);
    return t;
}

#define B2_CONSTANT_EMISSIVE_TEXTURE 65534

/* get_material_property<float> (Material.h:213-221): the red channel, 0 without texture */
B2_HD float material_property_float(const void* textures, V2 texcoords, int texture_index)
{
    if (texture_index == B2_NO_TEXTURE || texture_index == B2_CONSTANT_EMISSIVE_TEXTURE)
        return 0.0f;
    float rgba[4];
    sample_texture_rgba(textures, texture_index, false, texcoords, rgba);
    return rgba[0];
}

/*
 * Scene SoA + textures -> working material: get_intersection_material (Material.h:60-150) over read_partial_material
 * (MaterialPackedSoA.h:381-551) and DevicePackedTexturedMaterial::unpack (MaterialPacked.h:541-668). A parameter that has a texture
 * is skipped by read_partial_material (its packed word stays 0) and replaced by the texel afterwards; a lobe's secondary parameters
 * are loaded when the lobe weight is > 0 OR the weight has a texture.
 */
B2_HD Mat load_scene_material_textured(const b2pt_materials_soa& s, const void* textures, int i, V2 texcoords, bool white_furnace)
{
    const MatTextureIndices tx = load_texture_indices(s, i);
    Mat m = mat_defaults();
    uint32_t flags = s.flags[i];

    bool packed_emissive_texture_used = (flags >> B2_FLAG_EMISSIVE_TEXTURE_USED) & 1u;
    if (!packed_emissive_texture_used)
        m.emission = rgb(s.emission[i].r, s.emission[i].g, s.emission[i].b);

    uint32_t bcr = s.base_color_roughness[i];
    if (tx.base_color == B2_NO_TEXTURE)
        m.base_color = unpack_rgb24(bcr);
    /* roughness: skipped by read_partial_material when a roughness / roughness-metallic texture exists (the packed 0 is unpacked) */
    m.roughness = (tx.roughness == B2_NO_TEXTURE && tx.roughness_metallic == B2_NO_TEXTURE) ? unpack_hi8(bcr) : 0.0f;
    m.oren_nayar_sigma = s.oren_nayar_sigma[i];

    uint32_t mf90 = s.metallic_F90_and_metallic[i];
    m.metallic = unpack_hi8(mf90);
    uint32_t aniso = s.anisotropy_and_rotation_and_second_roughness[i];
    uint32_t mf82 = s.metallic_F82_packed_and_diffuse_transmission[i];
    if (m.metallic > 0.0f || tx.metallic != B2_NO_TEXTURE || tx.roughness_metallic != B2_NO_TEXTURE)
    {
        m.metallic_F90_falloff_exponent = s.metallic_F90_falloff_exponent[i];
        m.metallic_F82 = unpack_rgb24(mf82);
        m.metallic_F90 = unpack_rgb24(mf90);
        m.second_roughness_weight = unpack_f8(aniso, 2);
        m.second_roughness = unpack_f8(aniso, 3);
        m.do_metallic_energy_compensation = (flags >> B2_FLAG_METALLIC_ENERGY_COMPENSATION) & 1u;
    }
    m.anisotropy = tx.anisotropic == B2_NO_TEXTURE ? unpack_f8(aniso, 0) : 0.0f;
    m.anisotropy_rotation = unpack_f8(aniso, 1);

    uint32_t spec = s.specular_and_darkening_and_coat_roughness[i];
    m.specular = unpack_f8(spec, 0);
    if (m.specular > 0.0f || tx.specular != B2_NO_TEXTURE)
    {
        uint32_t sct = s.specular_color_and_tint_factor[i];
        m.specular_tint = unpack_hi8(sct);
        m.specular_color = unpack_rgb24(sct);
        m.specular_darkening = unpack_f8(spec, 1);
        m.do_specular_energy_compensation = (flags >> B2_FLAG_SPECULAR_ENERGY_COMPENSATION) & 1u;
    }

    uint32_t coat_w = s.coat_and_medium_absorption[i];
    m.coat = unpack_hi8(coat_w);
    if (m.coat > 0.0f || tx.coat != B2_NO_TEXTURE)
    {
        uint32_t crd = s.coat_roughening_darkening_anisotropy_and_rotation[i];
        m.coat_medium_absorption = unpack_rgb24(coat_w);
        m.coat_medium_thickness = s.coat_medium_thickness[i];
        m.coat_roughness = unpack_f8(spec, 2);
        m.coat_roughening = unpack_f8(crd, 0);
        m.coat_darkening = unpack_f8(crd, 1);
        m.coat_anisotropy = unpack_f8(crd, 2);
        m.coat_anisotropy_rotation = unpack_f8(crd, 3);
        m.coat_ior = s.coat_ior[i];
        m.do_coat_energy_compensation = (flags >> B2_FLAG_COAT_ENERGY_COMPENSATION) & 1u;
    }

    uint32_t sheen_w = s.sheen_and_color[i];
    uint32_t srtd = s.sheen_roughness_transmission_dispersion_thin_film[i];
    m.sheen = unpack_hi8(sheen_w);
    if (m.sheen > 0.0f || tx.sheen != B2_NO_TEXTURE)
    {
        m.sheen_roughness = unpack_f8(srtd, 0);
        m.sheen_color = unpack_rgb24(sheen_w);
    }

    m.ior = s.ior[i];
    m.diffuse_transmission = unpack_hi8(mf82);
    m.specular_transmission = unpack_f8(srtd, 1);
    if (m.specular_transmission > 0.0f || tx.specular_transmission != B2_NO_TEXTURE)
    {
        m.dispersion_scale = unpack_f8(srtd, 2);
        m.dispersion_abbe_number = s.dispersion_abbe_number[i];
        m.thin_walled = (flags >> B2_FLAG_THIN_WALLED) & 1u;
        m.do_glass_energy_compensation = (flags >> B2_FLAG_GLASS_ENERGY_COMPENSATION) & 1u;
    }
    if (m.specular_transmission > 0.0f || m.diffuse_transmission > 0.0f || tx.specular_transmission != B2_NO_TEXTURE)
    {
        m.absorption_at_distance = s.absorption_at_distance[i];
        m.absorption_color = unpack_rgb24(s.absorption_color_packed[i]);
    }

    uint32_t ahp = s.alpha_thin_film_hue_dielectric_priority[i];
    m.thin_film = unpack_f8(srtd, 3);
    if (m.thin_film > 0.0f)
    {
        m.thin_film_ior = s.thin_film_ior[i];
        m.thin_film_thickness = s.thin_film_thickness[i];
        m.thin_film_kappa_3 = s.thin_film_kappa_3[i];
        m.thin_film_hue_shift_degrees = unpack_f8(ahp, 1);
        m.thin_film_base_ior_override = s.thin_film_base_ior_override[i] != 0.0f ? 1.0f : 0.0f;
        m.thin_film_do_ior_override = (flags >> B2_FLAG_THIN_FILM_DO_IOR_OVERRIDE) & 1u;
    }
    m.alpha_opacity = unpack_f8(ahp, 0);
    m.dielectric_priority = (uint8_t)((ahp >> 16) & 0xFFu);

    /* ---- get_intersection_material (Material.h:60-150) ---- */
    if (white_furnace)
        m.base_color = rgb(1.0f);
    else if (tx.base_color != B2_NO_TEXTURE)
    {
        float rgba[4];
        sample_texture_rgba(textures, tx.base_color, true, texcoords, rgba);
        m.base_color = rgb(rgba[0], rgba[1], rgba[2]);
    }

    /* get_metallic_roughness: (roughness, metallic) */
    float tex_roughness, tex_metallic;
    if (tx.roughness_metallic != B2_NO_TEXTURE)
    {
        float rgba[4];
        sample_texture_rgba(textures, tx.roughness_metallic, false, texcoords, rgba);
        tex_roughness = rgba[1];
        tex_metallic = rgba[2];
        m.roughness = tex_roughness;
        m.metallic = tex_metallic;
    }
    else
    {
        tex_roughness = material_property_float(textures, texcoords, tx.roughness);
        tex_metallic = material_property_float(textures, texcoords, tx.metallic);
        if (tx.roughness != B2_NO_TEXTURE)
            m.roughness = tex_roughness;
        if (tx.metallic != B2_NO_TEXTURE)
            m.metallic = tex_metallic;
        tex_roughness = m.roughness;
    }

    if (tx.anisotropic != B2_NO_TEXTURE)
        m.anisotropy = material_property_float(textures, texcoords, tx.anisotropic);
    if (tx.specular != B2_NO_TEXTURE)
        m.specular = material_property_float(textures, texcoords, tx.specular);
    float coat = material_property_float(textures, texcoords, tx.coat);
    if (tx.coat != B2_NO_TEXTURE)
        m.coat = coat;
    else
        coat = m.coat;
    if (tx.sheen != B2_NO_TEXTURE)
        m.sheen = material_property_float(textures, texcoords, tx.sheen);
    if (tx.specular_transmission != B2_NO_TEXTURE)
        m.specular_transmission = material_property_float(textures, texcoords, tx.specular_transmission);

    Rgb emission = m.emission;
    if (!(tx.emission == B2_NO_TEXTURE || tx.emission == B2_CONSTANT_EMISSIVE_TEXTURE))
    {
        float rgba[4];
        sample_texture_rgba(textures, tx.emission, false, texcoords, rgba);
        emission = rgb(rgba[0], rgba[1], rgba[2]);
    }
    m.emissive_texture_used = tx.emission != B2_NO_TEXTURE;
    m.emission = emission;

    /* coat roughening of the base layer(s), Material.h:127-147: from the roughness READ ABOVE (texture or parameter) */
    if (coat > 0.0f && m.coat_roughening > 0.0f)
    {
        float base_roughness = tex_roughness;
        float coat_roughness = m.coat_roughness;
        float target = b2_pow_1_4(b2_min(1.0f, b2_pow_4(base_roughness) + 2.0f * b2_pow_4(coat_roughness)));
        float roughened = b2_lerp(base_roughness, target, coat);
        m.roughness = b2_lerp(base_roughness, roughened, m.coat_roughening);
        if (m.second_roughness_weight > 0.0f)
        {
            float second = m.second_roughness;
            float target2 = b2_pow_1_4(b2_min(1.0f, b2_pow_4(second) + 2.0f * b2_pow_4(coat_roughness)));
            float roughened2 = b2_lerp(second, target2, coat);
            m.second_roughness = b2_lerp(second, roughened2, m.coat_roughening);
        }
    }
    return m;
}

/* MaterialUtils (HostDeviceCommon/Material/MaterialUtils.h:16-178) */
B2_HD void material_alphas(float roughness, float anisotropy, float& alpha_x, float& alpha_y)
{
    float aspect = sqrtf(1.0f - 0.9f * anisotropy);
    alpha_x = b2_max(B2_ROUGHNESS_CLAMP, roughness * roughness / aspect);
    alpha_y = b2_max(B2_ROUGHNESS_CLAMP, roughness * roughness * aspect);
}
B2_HD bool material_is_perfectly_smooth(float roughness, float threshold = B2_PERFECTLY_SMOOTH_ROUGHNESS_THRESHOLD) { return roughness <= threshold; }

/* ---- octahedral 2x12-bit normal in 3 bytes + 1 pad (Packing.h:166-283) ---- */
B2_HD float oct_sign_not_zero(float k) { return k >= 0.0f ? 1.0f : -1.0f; }

B2_HD uint32_t pack_normal_oct24(V3 n)
{
    float l1norm_inv = 1.0f / (fabsf(n.x) + fabsf(n.y) + fabsf(n.z));
    float ex = n.x * l1norm_inv, ey = n.y * l1norm_inv;
    if (n.z < 0.0f)
    {
        float rx = (1.0f - fabsf(ey)) * oct_sign_not_zero(ex);
        float ry = (1.0f - fabsf(ex)) * oct_sign_not_zero(ey);
        ex = rx;
        ey = ry;
    }
    float sx = roundf(b2_clamp(0.0f, 2.0f, ex + 1.0f) * 2047.0f);
    float sy = roundf(b2_clamp(0.0f, 2.0f, ey + 1.0f) * 2047.0f);

    float ux = sx / 16.0f;
    float t = floorf(sy / 256.0f);
    float uy = ((ux - floorf(ux)) * 256.0f) + t;
    float uz = sy - (t * 256.0f);
    /* float -> unsigned char conversions, as the reference's `out_x = u.x` */
    return (uint32_t)(uint8_t)ux | ((uint32_t)(uint8_t)uy << 8) | ((uint32_t)(uint8_t)uz << 16);
}

B2_HD V3 unpack_normal_oct24(uint32_t w)
{
    float x = (float)(w & 0xFFu), y = (float)((w >> 8) & 0xFFu), z = (float)((w >> 16) & 0xFFu);
    float temp = y / 16.0f;
    float sx = x * 16.0f + floorf(temp);
    float sy = (temp - floorf(temp)) * 256.0f * 16.0f + z;
    float vx = b2_clamp(-1.0f, 1.0f, (sx / 2047.0f) - 1.0f);
    float vy = b2_clamp(-1.0f, 1.0f, (sy / 2047.0f) - 1.0f);
    V3 v = v3(vx, vy, 1.0f - fabsf(vx) - fabsf(vy));
    if (v.z < 0.0f)
    {
        float tx = v.x, ty = v.y;
        v.x = (1.0f - fabsf(ty)) * oct_sign_not_zero(tx);
        v.y = (1.0f - fabsf(tx)) * oct_sign_not_zero(ty);
    }
    return normalize(v);
}

#endif
