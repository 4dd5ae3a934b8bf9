/*
 * b2pt_scene_blob.h -- the flat, seeded, self-describing scene container shared by the product host
 * (hiprt-path-tracer_b200/csrc/host) and by the test oracle (oracle/).
 *
 * It plays the role of the reference's in-memory `Scene` struct after `SceneParser::parse_scene_file`
 * (/root/reference/src/Scene/SceneParser.cpp:22-225) + `parse_camera` (:227-283): flat vertex / index /
 * per-triangle-material arrays, unpacked material parameter records and one camera. Both sides read
 * the very same bytes, so "same scene and seeds" (BASELINE.json north_star) holds by construction.
 *
 * Layout (little endian):
 *   B2SceneHeader
 *   float3  positions[n_vertices]
 *   float3  normals[n_vertices]
 *   float2  texcoords[n_vertices]
 *   int32   indices[3 * n_triangles]
 *   int32   material_indices[n_triangles]
 *   uint8   has_normals[n_vertices]            (padded with zeros to a multiple of 4 bytes)
 *   float   materials[n_materials][B2MAT_FLOAT_COUNT]
 *   -- optional, only when the scene has textures (what the reference's Scene holds after ThreadFunctions::load_scene_texture,
 *      src/Threads/ThreadFunctions.cpp:33-133) --
 *   char    "B2TEXTR1"
 *   uint32  n_textures
 *   { uint32 width, height, channels; uint8 texels[height][width][channels] (row 0 first, as stb_image returns them, padded to 4 B) } x n_textures
 *   int32   texture_indices[n_materials][B2TEX_SLOT_COUNT]   (index into the textures, B2TEX_NONE, or B2TEX_CONSTANT_EMISSIVE)
 *   uint8   base_color_texture_opaque[n_materials]            (Scene::material_has_opaque_base_color_texture; 1 without texture)
 */
#ifndef B2PT_SCENE_BLOB_H
#define B2PT_SCENE_BLOB_H

#include <cstdint>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

/* One line per unpacked material parameter, in file order. Mirrors the field set of the reference's
 * CPUMaterial (/root/reference/src/HostDeviceCommon/Material/MaterialCPU.h:217-329); bools are 0/1,
 * the integer dielectric priority is stored as a float. Texture indices are not stored: textured
 * materials are a "next" row. */
#define B2MAT_FIELDS(X)                                                                                  \
    X(emission_r) X(emission_g) X(emission_b) X(emission_strength) X(global_emissive_factor)            \
    X(base_color_r) X(base_color_g) X(base_color_b)                                                     \
    X(roughness) X(oren_nayar_sigma)                                                                    \
    X(metallic) X(metallic_F90_falloff_exponent)                                                        \
    X(metallic_F82_r) X(metallic_F82_g) X(metallic_F82_b)                                               \
    X(metallic_F90_r) X(metallic_F90_g) X(metallic_F90_b)                                               \
    X(anisotropy) X(anisotropy_rotation) X(second_roughness_weight) X(second_roughness)                 \
    X(do_metallic_energy_compensation)                                                                  \
    X(specular) X(specular_tint) X(specular_color_r) X(specular_color_g) X(specular_color_b)            \
    X(specular_darkening) X(do_specular_energy_compensation)                                            \
    X(coat) X(coat_medium_absorption_r) X(coat_medium_absorption_g) X(coat_medium_absorption_b)         \
    X(coat_medium_thickness) X(coat_roughness) X(coat_roughening) X(coat_darkening)                     \
    X(coat_anisotropy) X(coat_anisotropy_rotation) X(coat_ior) X(do_coat_energy_compensation)           \
    X(sheen) X(sheen_roughness) X(sheen_color_r) X(sheen_color_g) X(sheen_color_b)                      \
    X(ior) X(specular_transmission) X(diffuse_transmission) X(absorption_at_distance)                   \
    X(absorption_color_r) X(absorption_color_g) X(absorption_color_b)                                   \
    X(dispersion_scale) X(dispersion_abbe_number) X(thin_walled) X(do_glass_energy_compensation)        \
    X(thin_film) X(thin_film_ior) X(thin_film_thickness) X(thin_film_kappa_3)                           \
    X(thin_film_hue_shift_degrees) X(thin_film_base_ior_override) X(thin_film_do_ior_override)          \
    X(alpha_opacity) X(dielectric_priority)

enum B2MatField : int
{
#define X(name) B2MAT_##name,
    B2MAT_FIELDS(X)
#undef X
    B2MAT_FLOAT_COUNT
};

/* texture slots of a material, in the order of CPUMaterial's *_texture_index members (MaterialCPU.h:313-328) */
enum B2TexSlot : int
{
    B2TEX_normal_map, B2TEX_emission, B2TEX_base_color, B2TEX_roughness_metallic, B2TEX_roughness, B2TEX_metallic, B2TEX_anisotropic,
    B2TEX_specular, B2TEX_coat, B2TEX_sheen, B2TEX_specular_transmission,
    B2TEX_SLOT_COUNT
};
#define B2TEX_NONE 65535              /* MaterialConstants::NO_TEXTURE */
#define B2TEX_CONSTANT_EMISSIVE 65534 /* MaterialConstants::CONSTANT_EMISSIVE_TEXTURE */

struct B2SceneTexture
{
    uint32_t width = 0, height = 0, channels = 0;
    std::vector<uint8_t> texels;
};

struct B2SceneCamera
{
    /* Row-major 4x4 matrices, right-multiplied by column points; same meaning as the reference's
     * HIPRTCamera (/root/reference/src/HostDeviceCommon/HIPRTCamera.h:11-25). */
    float inverse_view[16];
    float inverse_projection[16];
    float view_projection[16];
    float position[3];
    float vertical_fov;
};

struct B2SceneHeader
{
    char magic[8]; /* "B2SCENE1" */
    uint32_t n_vertices;
    uint32_t n_triangles;
    uint32_t n_materials;
    uint32_t mat_float_count; /* must equal B2MAT_FLOAT_COUNT */
    B2SceneCamera camera;
};

struct B2Scene
{
    B2SceneCamera camera;
    std::vector<float> positions;        /* 3 * n_vertices */
    std::vector<float> normals;          /* 3 * n_vertices */
    std::vector<float> texcoords;        /* 2 * n_vertices */
    std::vector<int32_t> indices;        /* 3 * n_triangles */
    std::vector<int32_t> material_indices; /* n_triangles */
    std::vector<uint8_t> has_normals;    /* n_vertices */
    std::vector<float> materials;        /* n_materials * B2MAT_FLOAT_COUNT */
    std::vector<B2SceneTexture> textures;           /* optional */
    std::vector<int32_t> texture_indices;           /* n_materials * B2TEX_SLOT_COUNT, or empty (no textures) */
    std::vector<uint8_t> base_color_texture_opaque; /* n_materials, or empty */

    int texture_index(size_t material, int slot) const { return texture_indices.empty() ? B2TEX_NONE : texture_indices[material * B2TEX_SLOT_COUNT + slot]; }
    bool base_color_opaque(size_t material) const { return base_color_texture_opaque.empty() ? true : base_color_texture_opaque[material] != 0; }

    size_t n_vertices() const { return has_normals.size(); }
    size_t n_triangles() const { return material_indices.size(); }
    size_t n_materials() const { return materials.size() / B2MAT_FLOAT_COUNT; }
    const float* material(size_t i) const { return materials.data() + i * B2MAT_FLOAT_COUNT; }
};

/* Returns an empty string on success, an error message otherwise. */
inline std::string b2scene_read(const char* path, B2Scene& out)
{
    FILE* f = std::fopen(path, "rb");
    if (!f)
        return std::string("cannot open scene blob: ") + path;

    B2SceneHeader h;
    if (std::fread(&h, sizeof(h), 1, f) != 1 || std::memcmp(h.magic, "B2SCENE1", 8) != 0)
    {
        std::fclose(f);
        return std::string("bad scene blob header: ") + path;
    }
    if (h.mat_float_count != B2MAT_FLOAT_COUNT)
    {
        std::fclose(f);
        return "scene blob material record size mismatch";
    }

    {
        /* a bad header must not size the vectors: every count is bounded by what the file can hold */
        long here = std::ftell(f);
        std::fseek(f, 0, SEEK_END);
        const unsigned long long file_bytes = (unsigned long long)std::ftell(f);
        std::fseek(f, here, SEEK_SET);
        const unsigned long long need = sizeof(h) + 33ull * h.n_vertices + ((4 - (h.n_vertices & 3)) & 3) + 16ull * h.n_triangles + 4ull * h.n_materials * B2MAT_FLOAT_COUNT;
        if (need > file_bytes)
        {
            std::fclose(f);
            return std::string("scene blob header counts exceed the file size: ") + path;
        }
    }
    out.camera = h.camera;
    out.positions.resize(3ull * h.n_vertices);
    out.normals.resize(3ull * h.n_vertices);
    out.texcoords.resize(2ull * h.n_vertices);
    out.indices.resize(3ull * h.n_triangles);
    out.material_indices.resize(h.n_triangles);
    out.has_normals.resize(h.n_vertices);
    out.materials.resize((size_t)h.n_materials * B2MAT_FLOAT_COUNT);

    bool ok = true;
    auto rd = [&](void* dst, size_t bytes) { if (bytes && std::fread(dst, 1, bytes, f) != bytes) ok = false; };
    rd(out.positions.data(), out.positions.size() * 4);
    rd(out.normals.data(), out.normals.size() * 4);
    rd(out.texcoords.data(), out.texcoords.size() * 4);
    rd(out.indices.data(), out.indices.size() * 4);
    rd(out.material_indices.data(), out.material_indices.size() * 4);
    rd(out.has_normals.data(), out.has_normals.size());
    size_t pad = (4 - (h.n_vertices & 3)) & 3;
    if (pad)
        std::fseek(f, (long)pad, SEEK_CUR);
    rd(out.materials.data(), out.materials.size() * 4);
    out.textures.clear();
    out.texture_indices.clear();
    out.base_color_texture_opaque.clear();
    char tag[8];
    if (ok && std::fread(tag, 1, 8, f) == 8 && std::memcmp(tag, "B2TEXTR1", 8) == 0)
    {
        long here = std::ftell(f);
        std::fseek(f, 0, SEEK_END);
        const unsigned long long file_bytes = (unsigned long long)std::ftell(f);
        std::fseek(f, here, SEEK_SET);
        uint32_t n_textures = 0;
        rd(&n_textures, 4);
        if (ok && 12ull * n_textures > file_bytes)
            ok = false;
        for (uint32_t t = 0; ok && t < n_textures; t++)
        {
            B2SceneTexture tex;
            uint32_t dims[3] = { 0, 0, 0 };
            rd(dims, 12);
            const unsigned long long bytes = (unsigned long long)dims[0] * dims[1] * dims[2];
            if (!ok || dims[2] < 1 || dims[2] > 4 || bytes > file_bytes)
            {
                ok = false;
                break;
            }
            tex.width = dims[0];
            tex.height = dims[1];
            tex.channels = dims[2];
            tex.texels.resize((size_t)bytes);
            rd(tex.texels.data(), tex.texels.size());
            const long pad = (long)((4 - (bytes & 3)) & 3);
            if (pad)
                std::fseek(f, pad, SEEK_CUR);
            out.textures.push_back(std::move(tex));
        }
        if (ok)
        {
            out.texture_indices.resize((size_t)h.n_materials * B2TEX_SLOT_COUNT);
            out.base_color_texture_opaque.resize(h.n_materials);
            rd(out.texture_indices.data(), out.texture_indices.size() * 4);
            rd(out.base_color_texture_opaque.data(), out.base_color_texture_opaque.size());
            for (int32_t idx : out.texture_indices)
                if (!(idx == B2TEX_NONE || idx == B2TEX_CONSTANT_EMISSIVE || (idx >= 0 && (size_t)idx < out.textures.size())))
                    ok = false;
        }
    }
    std::fclose(f);

    if (!ok)
        return std::string("truncated or corrupt scene blob: ") + path;
    return "";
}

inline std::string b2scene_write(const char* path, const B2Scene& s)
{
    FILE* f = std::fopen(path, "wb");
    if (!f)
        return std::string("cannot create scene blob: ") + path;

    B2SceneHeader h;
    std::memcpy(h.magic, "B2SCENE1", 8);
    h.n_vertices = (uint32_t)s.n_vertices();
    h.n_triangles = (uint32_t)s.n_triangles();
    h.n_materials = (uint32_t)s.n_materials();
    h.mat_float_count = B2MAT_FLOAT_COUNT;
    h.camera = s.camera;
    std::fwrite(&h, sizeof(h), 1, f);
    std::fwrite(s.positions.data(), 4, s.positions.size(), f);
    std::fwrite(s.normals.data(), 4, s.normals.size(), f);
    std::fwrite(s.texcoords.data(), 4, s.texcoords.size(), f);
    std::fwrite(s.indices.data(), 4, s.indices.size(), f);
    std::fwrite(s.material_indices.data(), 4, s.material_indices.size(), f);
    std::fwrite(s.has_normals.data(), 1, s.has_normals.size(), f);
    const char zeros[4] = { 0, 0, 0, 0 };
    std::fwrite(zeros, 1, (4 - (s.n_vertices() & 3)) & 3, f);
    std::fwrite(s.materials.data(), 4, s.materials.size(), f);
    if (!s.textures.empty())
    {
        std::fwrite("B2TEXTR1", 1, 8, f);
        const uint32_t n_textures = (uint32_t)s.textures.size();
        std::fwrite(&n_textures, 4, 1, f);
        for (const B2SceneTexture& t : s.textures)
        {
            const uint32_t dims[3] = { t.width, t.height, t.channels };
            std::fwrite(dims, 4, 3, f);
            std::fwrite(t.texels.data(), 1, t.texels.size(), f);
            std::fwrite(zeros, 1, (4 - (t.texels.size() & 3)) & 3, f);
        }
        std::fwrite(s.texture_indices.data(), 4, s.texture_indices.size(), f);
        std::fwrite(s.base_color_texture_opaque.data(), 1, s.base_color_texture_opaque.size(), f);
    }
    std::fclose(f);
    return "";
}

#endif
