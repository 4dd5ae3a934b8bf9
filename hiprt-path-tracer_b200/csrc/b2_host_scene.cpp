/* b2_host_scene.cpp -- see b2_host_scene.h for what this restates and why. */
#include "b2_host_scene.h"

#include "b2_material.h"
#include "b2pt_alias_table.h"

#include <chrono>
#include <cmath>

namespace
{
/* hippt::clamp on the host = fminf(max, fmaxf(min, v)) (HostDeviceCommon/Math.h:257-258) */
inline float clamp01(float v) { return fminf(1.0f, fmaxf(0.0f, v)); }
inline float clampf(float lo, float hi, float v) { return fminf(hi, fmaxf(lo, v)); }
inline float maxf_ref(float a, float b) { return a > b ? a : b; }

struct SafeMaterial
{
    float f[B2MAT_FLOAT_COUNT];
    int dielectric_priority;
};

/* CPUMaterial::make_safe, MaterialCPU.h:156-215 */
SafeMaterial make_safe(const float* rec)
{
    SafeMaterial m;
    for (int i = 0; i < B2MAT_FLOAT_COUNT; i++)
        m.f[i] = rec[i];
    auto clamp_rgb01 = [&](int i) { for (int k = 0; k < 3; k++) m.f[i + k] = clamp01(m.f[i + k]); };

    clamp_rgb01(B2MAT_base_color_r);
    m.f[B2MAT_metallic] = clamp01(m.f[B2MAT_metallic]);
    clamp_rgb01(B2MAT_metallic_F82_r);
    clamp_rgb01(B2MAT_metallic_F90_r);
    m.f[B2MAT_anisotropy] = clamp01(m.f[B2MAT_anisotropy]);
    m.f[B2MAT_anisotropy_rotation] = clamp01(m.f[B2MAT_anisotropy_rotation]);
    m.f[B2MAT_second_roughness_weight] = clamp01(m.f[B2MAT_second_roughness_weight]);
    m.f[B2MAT_second_roughness] = clamp01(m.f[B2MAT_second_roughness]);

    m.f[B2MAT_specular] = clamp01(m.f[B2MAT_specular]);
    m.f[B2MAT_specular_tint] = clamp01(m.f[B2MAT_specular_tint]);
    clamp_rgb01(B2MAT_specular_color_r);
    m.f[B2MAT_specular_darkening] = clamp01(m.f[B2MAT_specular_darkening]);

    m.f[B2MAT_coat] = clamp01(m.f[B2MAT_coat]);
    clamp_rgb01(B2MAT_coat_medium_absorption_r);
    m.f[B2MAT_coat_roughness] = clampf(B2_ROUGHNESS_CLAMP, 1.0f, m.f[B2MAT_coat_roughness]);
    m.f[B2MAT_coat_roughening] = clamp01(m.f[B2MAT_coat_roughening]);
    m.f[B2MAT_coat_darkening] = clamp01(m.f[B2MAT_coat_darkening]);
    m.f[B2MAT_coat_anisotropy] = clamp01(m.f[B2MAT_coat_anisotropy]);
    m.f[B2MAT_coat_anisotropy_rotation] = clamp01(m.f[B2MAT_coat_anisotropy_rotation]);

    m.f[B2MAT_sheen] = clamp01(m.f[B2MAT_sheen]);
    m.f[B2MAT_sheen_roughness] = clampf(B2_ROUGHNESS_CLAMP, 1.0f, m.f[B2MAT_sheen_roughness]);
    clamp_rgb01(B2MAT_sheen_color_r);

    m.f[B2MAT_specular_transmission] = clamp01(m.f[B2MAT_specular_transmission]);
    m.f[B2MAT_absorption_at_distance] = maxf_ref(m.f[B2MAT_absorption_at_distance], 1.0e-4f);
    for (int k = 0; k < 3; k++)
        m.f[B2MAT_absorption_color_r + k] = clamp01(maxf_ref(m.f[B2MAT_absorption_color_r + k], 1.0f / 512.0f));

    m.f[B2MAT_dispersion_abbe_number] = maxf_ref(1.0e-5f, m.f[B2MAT_dispersion_abbe_number]);
    m.f[B2MAT_dispersion_scale] = clamp01(m.f[B2MAT_dispersion_scale]);

    m.f[B2MAT_thin_film] = clamp01(m.f[B2MAT_thin_film]);
    m.f[B2MAT_thin_film_hue_shift_degrees] = clampf(0.0f, 360.0f, m.f[B2MAT_thin_film_hue_shift_degrees]);
    m.f[B2MAT_thin_film_ior] = maxf_ref(1.0005f, m.f[B2MAT_thin_film_ior]);

    m.f[B2MAT_alpha_opacity] = clamp01(m.f[B2MAT_alpha_opacity]);

    int priority = (int)m.f[B2MAT_dielectric_priority];
    priority = std::min(15, std::max(0, priority));
    for (int k = 0; k < 3; k++)
        m.f[B2MAT_emission_r + k] = maxf_ref(0.0f, m.f[B2MAT_emission_r + k]);
    if (m.f[B2MAT_specular_transmission] == 0.0f && m.f[B2MAT_diffuse_transmission] == 0.0f)
        priority = (1 << 15) - 1;
    m.dielectric_priority = priority;
    return m;
}

inline Rgb rgb_at(const SafeMaterial& m, int i) { return rgb(m.f[i], m.f[i + 1], m.f[i + 2]); }

/* CPUMaterial::pack_to_GPU, MaterialCPU.h:31-120 */
B2PackedTexturedMaterial pack_to_gpu(const SafeMaterial& m, const int* tex /* B2TEX_SLOT_COUNT indices, or null */)
{
    B2PackedTexturedMaterial out;
    Mat u = mat_defaults();
    Rgb e = rgb_at(m, B2MAT_emission_r) * m.f[B2MAT_emission_strength] * m.f[B2MAT_global_emissive_factor];
    u.emission = e;
    u.emissive_texture_used = false;
    u.base_color = rgb_at(m, B2MAT_base_color_r);
    u.roughness = m.f[B2MAT_roughness];
    u.oren_nayar_sigma = m.f[B2MAT_oren_nayar_sigma];
    u.metallic = m.f[B2MAT_metallic];
    u.metallic_F90_falloff_exponent = m.f[B2MAT_metallic_F90_falloff_exponent];
    u.metallic_F82 = rgb_at(m, B2MAT_metallic_F82_r);
    u.metallic_F90 = rgb_at(m, B2MAT_metallic_F90_r);
    u.anisotropy = m.f[B2MAT_anisotropy];
    u.anisotropy_rotation = m.f[B2MAT_anisotropy_rotation];
    u.second_roughness_weight = m.f[B2MAT_second_roughness_weight];
    u.second_roughness = m.f[B2MAT_second_roughness];
    u.do_metallic_energy_compensation = m.f[B2MAT_do_metallic_energy_compensation] != 0.0f;
    u.specular = m.f[B2MAT_specular];
    u.specular_tint = m.f[B2MAT_specular_tint];
    u.specular_color = rgb_at(m, B2MAT_specular_color_r);
    u.specular_darkening = m.f[B2MAT_specular_darkening];
    u.do_specular_energy_compensation = m.f[B2MAT_do_specular_energy_compensation] != 0.0f;
    u.coat = m.f[B2MAT_coat];
    u.coat_medium_absorption = rgb_at(m, B2MAT_coat_medium_absorption_r);
    u.coat_medium_thickness = m.f[B2MAT_coat_medium_thickness];
    u.coat_roughness = m.f[B2MAT_coat_roughness];
    u.coat_roughening = m.f[B2MAT_coat_roughening];
    u.coat_darkening = m.f[B2MAT_coat_darkening];
    u.coat_anisotropy = m.f[B2MAT_coat_anisotropy];
    u.coat_anisotropy_rotation = m.f[B2MAT_coat_anisotropy_rotation];
    u.coat_ior = m.f[B2MAT_coat_ior];
    u.do_coat_energy_compensation = m.f[B2MAT_do_coat_energy_compensation] != 0.0f;
    u.sheen = m.f[B2MAT_sheen];
    u.sheen_roughness = m.f[B2MAT_sheen_roughness];
    u.sheen_color = rgb_at(m, B2MAT_sheen_color_r);
    u.ior = m.f[B2MAT_ior];
    u.specular_transmission = m.f[B2MAT_specular_transmission];
    u.diffuse_transmission = m.f[B2MAT_diffuse_transmission];
    u.absorption_at_distance = m.f[B2MAT_absorption_at_distance];
    u.absorption_color = rgb_at(m, B2MAT_absorption_color_r);
    u.dispersion_scale = m.f[B2MAT_dispersion_scale];
    u.dispersion_abbe_number = m.f[B2MAT_dispersion_abbe_number];
    u.thin_walled = m.f[B2MAT_thin_walled] != 0.0f;
    u.do_glass_energy_compensation = m.f[B2MAT_do_glass_energy_compensation] != 0.0f;
    u.thin_film = m.f[B2MAT_thin_film];
    u.thin_film_ior = m.f[B2MAT_thin_film_ior];
    u.thin_film_thickness = m.f[B2MAT_thin_film_thickness];
    u.thin_film_kappa_3 = m.f[B2MAT_thin_film_kappa_3];
    u.thin_film_hue_shift_degrees = m.f[B2MAT_thin_film_hue_shift_degrees] / 360.0f;
    u.thin_film_base_ior_override = m.f[B2MAT_thin_film_base_ior_override];
    u.thin_film_do_ior_override = m.f[B2MAT_thin_film_do_ior_override] != 0.0f;
    u.alpha_opacity = m.f[B2MAT_alpha_opacity];
    /* set_dielectric_priority takes an unsigned char: 32767 becomes 255 */
    u.dielectric_priority = (uint8_t)m.dielectric_priority;
    /* pack_to_GPU never sets the strong-energy-conservation flag nor the sample count: packed defaults (0) */
    u.enforce_strong_energy_conservation = false;
    u.energy_preservation_monte_carlo_samples = 0;

    PackedMat p = pack_material<0>(u);
    for (int i = 0; i < 27; i++)
        out.effective[i] = p.w[i];

    /* texture indices, with the reference's setters (MaterialPacked.h:725-735, in CPUMaterial::pack_to_GPU's order): the anisotropic
     * index is written into the ROUGHNESS slot of the roughness/metallic word (:731), after the roughness index, so it replaces it;
     * the anisotropic half of the anisotropic/specular word is never written and stays 0 */
    auto idx = [&](int slot) -> uint32_t { return tex ? (uint32_t)tex[slot] & 0xFFFFu : 0xFFFFu; };
    out.texture_indices[0] = idx(B2TEX_normal_map) | (idx(B2TEX_emission) << 16);
    out.texture_indices[1] = idx(B2TEX_base_color) | (idx(B2TEX_roughness_metallic) << 16);
    out.texture_indices[2] = idx(B2TEX_anisotropic) | (idx(B2TEX_metallic) << 16);
    out.texture_indices[3] = idx(B2TEX_specular) << 16;
    out.texture_indices[4] = idx(B2TEX_coat) | (idx(B2TEX_sheen) << 16);
    out.texture_indices[5] = idx(B2TEX_specular_transmission);
    return out;
}
} // namespace

bool b2_host_scene_prepare(B2HostScene& scene, std::string& error, bool build_bvh)
{
    const B2Scene& blob = scene.blob;
    const size_t n_mat = blob.n_materials(), n_tri = blob.n_triangles(), n_vert = blob.n_vertices();
    if (n_mat >= 0x03FFFFFFu)
    {
        error = "too many materials for the nested-dielectrics stack entries";
        return false;
    }
    for (size_t t = 0; t < n_tri; t++)
        if (blob.material_indices[t] < 0 || (size_t)blob.material_indices[t] >= n_mat)
        {
            error = "material index out of range";
            return false;
        }
    /* vertex indices are dereferenced below (triangle areas, emissive power) before the BVH builder would look at them */
    if (blob.indices.size() != 3 * n_tri || blob.positions.size() != 3 * n_vert || blob.normals.size() != 3 * n_vert || blob.texcoords.size() != 2 * n_vert)
    {
        error = "scene arrays have inconsistent sizes";
        return false;
    }
    for (size_t i = 0; i < 3 * n_tri; i++)
        if (blob.indices[i] < 0 || (size_t)blob.indices[i] >= n_vert)
        {
            error = "vertex index out of range";
            return false;
        }

    std::vector<SafeMaterial> safe(n_mat);
    scene.has_alpha_tested_materials = false;
    scene.packed_materials.resize(n_mat);
    scene.material_opaque.resize(n_mat);
    for (size_t i = 0; i < n_mat; i++)
    {
        safe[i] = make_safe(blob.material(i));
        scene.packed_materials[i] = pack_to_gpu(safe[i], blob.texture_indices.empty() ? nullptr : &blob.texture_indices[i * B2TEX_SLOT_COUNT]);
        /* CPURenderer.cpp:312: opaque = fully opaque base colour texture (or none) and alpha_opacity == 1 */
        scene.material_opaque[i] = blob.base_color_opaque(i) && safe[i].f[B2MAT_alpha_opacity] == 1.0f;
        if (!scene.material_opaque[i])
            scene.has_alpha_tested_materials = true; /* the kernels are then asked for the alpha test (B2PT_ALPHA_TEST) */
    }

    /* AoS -> SoA */
    B2MaterialSoAHost& s = scene.materials;
    auto col = [&](std::vector<uint32_t>& dst, int word) { dst.resize(n_mat); for (size_t i = 0; i < n_mat; i++) dst[i] = scene.packed_materials[i].effective[word]; };
    auto colf = [&](std::vector<float>& dst, int word) { dst.resize(n_mat); for (size_t i = 0; i < n_mat; i++) dst[i] = b2_uint_as_float(scene.packed_materials[i].effective[word]); };
    auto coli = [&](std::vector<uint32_t>& dst, int word) { dst.resize(n_mat); for (size_t i = 0; i < n_mat; i++) dst[i] = scene.packed_materials[i].texture_indices[word]; };
    s.flags.resize(n_mat);
    s.emission.resize(3 * n_mat);
    for (size_t i = 0; i < n_mat; i++)
    {
        s.flags[i] = (uint8_t)scene.packed_materials[i].effective[PM_FLAGS];
        for (int k = 0; k < 3; k++)
            s.emission[3 * i + k] = b2_uint_as_float(scene.packed_materials[i].effective[PM_EMISSION_R + k]);
    }
    col(s.base_color_roughness, PM_BASE_COLOR_ROUGHNESS);
    colf(s.oren_nayar_sigma, PM_OREN_SIGMA);
    col(s.metallic_F90_and_metallic, PM_METALLIC_F90_AND_METALLIC);
    col(s.metallic_F82_packed_and_diffuse_transmission, PM_METALLIC_F82_AND_DIFFUSE_TRANSMISSION);
    colf(s.metallic_F90_falloff_exponent, PM_METALLIC_F90_FALLOFF);
    col(s.anisotropy_and_rotation_and_second_roughness, PM_ANISO_ROT_SECOND_ROUGHNESS);
    col(s.specular_color_and_tint_factor, PM_SPECULAR_COLOR_TINT);
    col(s.specular_and_darkening_and_coat_roughness, PM_SPECULAR_DARKENING_COAT_ROUGHNESS);
    colf(s.coat_medium_thickness, PM_COAT_MEDIUM_THICKNESS);
    col(s.coat_and_medium_absorption, PM_COAT_AND_MEDIUM_ABSORPTION);
    col(s.coat_roughening_darkening_anisotropy_and_rotation, PM_COAT_ROUGHENING_DARKENING_ANISO);
    colf(s.coat_ior, PM_COAT_IOR);
    col(s.sheen_and_color, PM_SHEEN_AND_COLOR);
    colf(s.ior, PM_IOR);
    col(s.absorption_color_packed, PM_ABSORPTION_COLOR);
    colf(s.absorption_at_distance, PM_ABSORPTION_DISTANCE);
    col(s.sheen_roughness_transmission_dispersion_thin_film, PM_SHEEN_ROUGHNESS_TRANSMISSION_DISPERSION_THIN_FILM);
    colf(s.dispersion_abbe_number, PM_DISPERSION_ABBE);
    colf(s.thin_film_ior, PM_THIN_FILM_IOR);
    colf(s.thin_film_thickness, PM_THIN_FILM_THICKNESS);
    colf(s.thin_film_kappa_3, PM_THIN_FILM_KAPPA_3);
    colf(s.thin_film_base_ior_override, PM_THIN_FILM_BASE_IOR_OVERRIDE);
    col(s.alpha_thin_film_hue_dielectric_priority, PM_ALPHA_HUE_PRIORITY);
    coli(s.normal_map_emission_index, 0);
    coli(s.base_color_roughness_metallic_index, 1);
    coli(s.roughness_and_metallic_index, 2);
    coli(s.anisotropic_specular_index, 3);
    coli(s.coat_sheen_index, 4);
    coli(s.specular_transmission_index, 5);

    /* triangle areas (ThreadFunctions.cpp:221-243) and the emissive list (:136-194) */
    scene.triangle_areas.resize(n_tri);
    scene.emissive_triangles.clear();
    scene.emissive_triangles_and_textures.clear();
    /* ThreadFunctions.cpp:136-194: a mesh is emissive when its material emits or has a (non-constant) emission texture; triangles
     * with an emission texture are not importance sampled: they are left out of the sampled list but kept in the second one */
    std::vector<uint8_t> mat_emissive(n_mat), mat_emission_textured(n_mat);
    for (size_t i = 0; i < n_mat; i++)
    {
        const int emission_texture = blob.texture_index(i, B2TEX_emission);
        mat_emission_textured[i] = emission_texture != B2TEX_NONE && emission_texture != B2TEX_CONSTANT_EMISSIVE;
        mat_emissive[i] = !b2_is_zero(safe[i].f[B2MAT_emission_r]) || !b2_is_zero(safe[i].f[B2MAT_emission_g]) || !b2_is_zero(safe[i].f[B2MAT_emission_b]) || mat_emission_textured[i];
    }
    auto vertex = [&](int idx) { return v3(blob.positions[3 * (size_t)idx], blob.positions[3 * (size_t)idx + 1], blob.positions[3 * (size_t)idx + 2]); };
    for (size_t t = 0; t < n_tri; t++)
    {
        V3 A = vertex(blob.indices[3 * t]), B = vertex(blob.indices[3 * t + 1]), C = vertex(blob.indices[3 * t + 2]);
        V3 normal = cross(B - A, C - A);
        scene.triangle_areas[t] = length(normal) * 0.5f;
        if (mat_emissive[blob.material_indices[t]])
        {
            if (!mat_emission_textured[blob.material_indices[t]])
                scene.emissive_triangles.push_back((int)t);
            scene.emissive_triangles_and_textures.push_back((int)t);
        }
    }

    /* power-proportional alias table (CPURenderer.cpp:420-465): built for every scene, used by the POWER variants */
    std::vector<float> power_list(scene.emissive_triangles.size());
    float power_sum = 0.0f;
    for (size_t i = 0; i < scene.emissive_triangles.size(); i++)
    {
        int t = scene.emissive_triangles[i];
        V3 A = vertex(blob.indices[3 * (size_t)t]), B = vertex(blob.indices[3 * (size_t)t + 1]), C = vertex(blob.indices[3 * (size_t)t + 2]);
        V3 normal = cross(B - A, C - A);
        float triangle_area = 0.5f * length(normal);
        const SafeMaterial& m = safe[blob.material_indices[t]];
        float emission_luminance = luminance(rgb_at(m, B2MAT_emission_r)) * m.f[B2MAT_emission_strength] * m.f[B2MAT_global_emissive_factor];
        float area_power = emission_luminance * triangle_area;
        power_list[i] = area_power;
        power_sum += area_power;
    }
    scene.power_alias_probas.clear();
    scene.power_alias_alias.clear();
    b2pt_build_alias_table(power_list, power_sum, scene.power_alias_probas, scene.power_alias_alias);
    scene.power_sum = power_sum;

    if (!build_bvh)
        return true;
    auto t0 = std::chrono::high_resolution_clock::now();
    if (!b2_build_cw8(blob.indices.data(), n_tri, blob.positions.data(), n_vert, scene.bvh, error))
        return false;
    scene.bvh_build_seconds = std::chrono::duration<double>(std::chrono::high_resolution_clock::now() - t0).count();
    return true;
}

void b2_host_envmap_prepare(B2HostEnvmap& out, const float* rgba, int width, int height)
{
    const size_t n = (size_t)width * height;
    out.width = width;
    out.height = height;
    std::vector<float> lum(n);
    out.packed.resize(n);
    float total = 0.0f;
    const float weights[3] = { 0.3086f, 0.6094f, 0.0820f };
    for (size_t i = 0; i < n; i++)
    {
        float l = 0.0f;
        for (int c = 0; c < 3; c++)
            l += rgba[4 * i + c] * weights[c];
        total += l;
        lum[i] = l;

        const float k_max = b2_uint_as_float(0x477F8000u), k_min = b2_uint_as_float(0x37800000u);
        float cr = fminf(k_max, fmaxf(0.0f, rgba[4 * i + 0])), cg = fminf(k_max, fmaxf(0.0f, rgba[4 * i + 1])), cb = fminf(k_max, fmaxf(0.0f, rgba[4 * i + 2]));
        float max_channel = b2_max(b2_max(k_min, cr), b2_max(cg, cb));
        float bias = b2_uint_as_float((b2_float_as_uint(max_channel) + 0x07804000u) & 0x7F800000u);
        uint32_t R = b2_float_as_uint(cr + bias), G = b2_float_as_uint(cg + bias), B = b2_float_as_uint(cb + bias);
        uint32_t E = (b2_float_as_uint(bias) << 4) + 0x10000000u;
        out.packed[i] = E | B << 18 | G << 9 | (R & 0x1FFu);
    }
    out.total_luminance = total;
    b2pt_build_alias_table(lum, total, out.alias_probas, out.alias_alias);
}

bool b2_envmap_blob_read(const char* path, std::vector<float>& rgba, int& width, int& height)
{
    FILE* f = std::fopen(path, "rb");
    char magic[8];
    int32_t dims[2];
    if (!f || std::fread(magic, 1, 8, f) != 8 || std::memcmp(magic, "B2ENVMP1", 8) != 0 || std::fread(dims, 4, 2, f) != 2 || dims[0] <= 0 || dims[1] <= 0)
    {
        if (f) std::fclose(f);
        return false;
    }
    rgba.resize((size_t)dims[0] * dims[1] * 4);
    bool ok = std::fread(rgba.data(), 4, rgba.size(), f) == rgba.size();
    std::fclose(f);
    width = dims[0];
    height = dims[1];
    return ok;
}

/* Radiance RGBE (.hdr) reader, written from the format description (Ward, "Real Pixels", Graphics Gems II; the file header of
 * radiance/src/common/color.c): text header up to an empty line with FORMAT=32-bit_rle_rgbe, a resolution line "-Y h +X w", then
 * scanlines, each either flat (4 bytes per pixel) or "new RLE" (2, 2, width hi, width lo, then the four channels run-length coded one
 * after the other). A pixel (r, g, b, e) decodes to (r, g, b) * 2^(e - 136), e = 0 to black -- what stb_image's loader, which the
 * reference calls (Image32Bit::read_image_hdr, src/Image/Image.cpp:389-417), computes; alpha = 1. flip_y: last scanline first. */
bool b2_hdr_read(const char* path, bool flip_y, std::vector<float>& rgba, int& width, int& height, std::string& error)
{
    FILE* f = std::fopen(path, "rb");
    if (!f) { error = std::string("cannot open ") + path; return false; }
    auto fail = [&](const char* what) { error = std::string(what) + ": " + path; std::fclose(f); return false; };
    char line[1024];
    if (!std::fgets(line, sizeof(line), f) || (std::strncmp(line, "#?RADIANCE", 10) != 0 && std::strncmp(line, "#?RGBE", 6) != 0))
        return fail("not a Radiance .hdr file");
    bool format_ok = false;
    while (true)
    {
        if (!std::fgets(line, sizeof(line), f))
            return fail("truncated .hdr header");
        if (line[0] == '\n' || (line[0] == '\r' && line[1] == '\n'))
            break;
        if (std::strncmp(line, "FORMAT=32-bit_rle_rgbe", 22) == 0)
            format_ok = true;
    }
    if (!format_ok)
        return fail("unsupported .hdr pixel format");
    if (!std::fgets(line, sizeof(line), f) || std::sscanf(line, "-Y %d +X %d", &height, &width) != 2 || width <= 0 || height <= 0 || width > 65535 || height > 65535)
        return fail("unsupported .hdr resolution line");

    rgba.assign((size_t)width * height * 4, 0.0f);
    std::vector<unsigned char> scan((size_t)width * 4);
    auto decode = [&](const unsigned char* p, float* out) {
        if (p[3] != 0)
        {
            const float scale = std::ldexp(1.0f, (int)p[3] - 136);
            out[0] = p[0] * scale;
            out[1] = p[1] * scale;
            out[2] = p[2] * scale;
        }
        else
            out[0] = out[1] = out[2] = 0.0f;
        out[3] = 1.0f;
    };
    for (int y = 0; y < height; y++)
    {
        unsigned char head[4];
        if (std::fread(head, 1, 4, f) != 4)
            return fail("truncated .hdr pixel data");
        const bool rle = width >= 8 && width < 32768 && head[0] == 2 && head[1] == 2 && (head[2] & 0x80) == 0;
        if (rle)
        {
            if (((int)head[2] << 8 | head[3]) != width)
                return fail("corrupt .hdr scanline");
            for (int c = 0; c < 4; c++)
            {
                int x = 0;
                while (x < width)
                {
                    int count = std::fgetc(f);
                    if (count == EOF)
                        return fail("truncated .hdr pixel data");
                    if (count > 128)
                    {
                        count -= 128;
                        const int value = std::fgetc(f);
                        if (value == EOF || count == 0 || x + count > width)
                            return fail("corrupt .hdr run");
                        for (int i = 0; i < count; i++)
                            scan[(size_t)(x++) * 4 + c] = (unsigned char)value;
                    }
                    else
                    {
                        if (count == 0 || x + count > width)
                            return fail("corrupt .hdr run");
                        for (int i = 0; i < count; i++)
                        {
                            const int value = std::fgetc(f);
                            if (value == EOF)
                                return fail("truncated .hdr pixel data");
                            scan[(size_t)(x++) * 4 + c] = (unsigned char)value;
                        }
                    }
                }
            }
        }
        else
        {
            std::memcpy(scan.data(), head, 4);
            if (width > 1 && std::fread(scan.data() + 4, 4, (size_t)width - 1, f) != (size_t)width - 1)
                return fail("truncated .hdr pixel data");
        }
        float* row = rgba.data() + (size_t)(flip_y ? height - 1 - y : y) * width * 4;
        for (int x = 0; x < width; x++)
            decode(scan.data() + (size_t)x * 4, row + (size_t)x * 4);
    }
    std::fclose(f);
    return true;
}
