/*
 * b2_host_scene.h -- host-side preparation of the scene arrays the kernels read through HIPRTRenderData.
 *
 * In the reference this work is spread over the scene loader and the renderer's set_scene():
 *   CPUMaterial::make_safe / pack_to_GPU            src/HostDeviceCommon/Material/MaterialCPU.h:31-215
 *   AoS -> SoA expansion of packed materials        src/Renderer/CPUDataStructures/MaterialPackedSoACPUData.h:117-170
 *   material_opaque                                 src/Renderer/CPURenderer.cpp:308-312
 *   emissive triangle list, triangle areas          src/Threads/ThreadFunctions.cpp:136-243
 *   power alias table                               src/Renderer/CPURenderer.cpp:420-465 + src/Utils/Utils.cpp:63-131
 * It is restated here (host C++, no CUDA) because the B200 host mirror (b2_renderer.cpp) has to produce the
 * same bytes the reference host would upload. tests/ compare these arrays with the oracle's dumps bit for bit.
 */
#ifndef B2_HOST_SCENE_H
#define B2_HOST_SCENE_H

#include "b2_bvh_build.h"
#include "b2pt_scene_blob.h"

#include <string>
#include <vector>

/* One packed material record as the reference lays it out in memory (DevicePackedTexturedMaterial, 132 bytes):
 * 27 words of DevicePackedEffectiveMaterial followed by 6 words of texture indices. */
struct B2PackedTexturedMaterial
{
    uint32_t effective[27];
    uint32_t texture_indices[6];
};
static_assert(sizeof(B2PackedTexturedMaterial) == 132, "packed textured material must be 132 bytes");

struct B2MaterialSoAHost
{
    std::vector<uint8_t> flags;
    std::vector<float> emission; /* 3 per material */
    std::vector<uint32_t> base_color_roughness;
    std::vector<float> oren_nayar_sigma;
    std::vector<uint32_t> metallic_F90_and_metallic;
    std::vector<uint32_t> metallic_F82_packed_and_diffuse_transmission;
    std::vector<float> metallic_F90_falloff_exponent;
    std::vector<uint32_t> anisotropy_and_rotation_and_second_roughness;
    std::vector<uint32_t> specular_color_and_tint_factor;
    std::vector<uint32_t> specular_and_darkening_and_coat_roughness;
    std::vector<float> coat_medium_thickness;
    std::vector<uint32_t> coat_and_medium_absorption;
    std::vector<uint32_t> coat_roughening_darkening_anisotropy_and_rotation;
    std::vector<float> coat_ior;
    std::vector<uint32_t> sheen_and_color;
    std::vector<float> ior;
    std::vector<uint32_t> absorption_color_packed;
    std::vector<float> absorption_at_distance;
    std::vector<uint32_t> sheen_roughness_transmission_dispersion_thin_film;
    std::vector<float> dispersion_abbe_number;
    std::vector<float> thin_film_ior;
    std::vector<float> thin_film_thickness;
    std::vector<float> thin_film_kappa_3;
    std::vector<float> thin_film_base_ior_override;
    std::vector<uint32_t> alpha_thin_film_hue_dielectric_priority;
    std::vector<uint32_t> normal_map_emission_index;
    std::vector<uint32_t> base_color_roughness_metallic_index;
    std::vector<uint32_t> roughness_and_metallic_index;
    std::vector<uint32_t> anisotropic_specular_index;
    std::vector<uint32_t> coat_sheen_index;
    std::vector<uint32_t> specular_transmission_index;
};

struct B2HostScene
{
    B2Scene blob;
    std::vector<B2PackedTexturedMaterial> packed_materials;
    B2MaterialSoAHost materials;
    std::vector<uint8_t> material_opaque;
    bool has_alpha_tested_materials = false; /* some material_opaque entry is 0 */
    std::vector<float> triangle_areas;
    std::vector<int> emissive_triangles;
    std::vector<int> emissive_triangles_and_textures; /* + triangles whose material has an emission texture (not importance sampled) */
    std::vector<float> power_alias_probas;
    std::vector<int> power_alias_alias;
    float power_sum = 0.0f;
    B2BvhHost bvh;
    double bvh_build_seconds = 0.0;
};

/* Environment map as the kernels read it: RGBE9995-packed texels + luminance alias table.
 * CPURenderer::set_envmap (src/Renderer/CPURenderer.cpp:467-508, ESS_ALIAS_TABLE branch), Image32Bit::compute_alias_table
 * (src/Image/Image.cpp:536-545,652-671), RGBE9995Packed::pack (src/HostDeviceCommon/Packing.h:443-468). */
struct B2HostEnvmap
{
    int width = 0, height = 0;
    std::vector<uint32_t> packed;
    std::vector<float> alias_probas;
    std::vector<int> alias_alias;
    float total_luminance = 0.0f;
};
/* rgba: width*height*4 floats, row 0 first (what Image32Bit holds after read_image_hdr(path, 4, flipY = true)) */
void b2_host_envmap_prepare(B2HostEnvmap& out, const float* rgba, int width, int height);
/* "B2ENVMP1" blob: int32 width, height, then width*height*4 floats. Returns false when the file cannot be read. */
bool b2_envmap_blob_read(const char* path, std::vector<float>& rgba, int& width, int& height);

/* Fills every derived array of `scene` from `scene.blob`. Returns false and sets `error` on unsupported input
 * (dispersive media, > 2^26 - 1 materials, invalid indices). */
bool b2_host_scene_prepare(B2HostScene& scene, std::string& error, bool build_bvh = true);

bool b2_hdr_read(const char* path, bool flip_y, std::vector<float>& rgba, int& width, int& height, std::string& error);

#endif
