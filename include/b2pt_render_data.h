/*
 * b2pt_render_data.h -- the by-value kernel argument, layout-identical to the reference's HIPRTRenderData.
 *
 * Every per-pixel kernel of the reference takes ONE argument: `HIPRTRenderData` by value, 2784 bytes on x86-64
 * (/root/reference/src/HostDeviceCommon/RenderData.h:93-129; launched at src/Compiler/GPUKernel.cpp:156-165 with
 * args = { &render_data }). The reference host (GPURenderer, the render passes) fills that struct and the buffers
 * it points to; a drop-in replacement of the device side therefore has to read THOSE bytes. This header restates
 * the layout as plain C structs: the fields the path-tracing hot path uses carry the reference's field names, the
 * sub-structs of features outside the hot path (ReSTIR GI, ReGIR, NEE++, GMoN) are opaque byte blocks of the
 * right size. oracle/layout_check.cpp static_asserts every offset and size below against the reference's own
 * headers (compiled here, where /root/reference exists), so a drift in either direction breaks the build.
 *
 * Conventions that differ from the reference and why it is still a drop-in:
 *  - `GPU_BVH` holds the device handle returned by b2pt_bvh_device_handle() (a b2pt wide-BVH descriptor) instead
 *    of a hiprtGeometry; the reference host only stores and forwards that pointer.
 *  - `bsdfs_data.*_albedo` hold b2pt LUT handles (device pointer to a b2pt_lut descriptor: linear fp32 texels,
 *    fetched nearest-texel exactly like the CPU reference, src/Image/Image.cpp:564-600,761-803) instead of
 *    oroTextureObject_t.
 *  - ReSTIR DI reservoir buffers are the same `width*height*36`-byte allocations the reference host makes
 *    (src/Renderer/RenderPasses/ReSTIRDIRenderPass.h:108-115) but the device reads and writes them as 9 planes of
 *    4-byte words (SoA) -- the host never looks inside them (it only ping-pongs the pointers).
 */
#ifndef B2PT_RENDER_DATA_H
#define B2PT_RENDER_DATA_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct { float x, y, z; } b2pt_float3;
typedef struct { float x, y; } b2pt_float2;
typedef struct { int x, y; } b2pt_int2;
typedef struct { float r, g, b; } b2pt_rgb;

/* src/Device/includes/AliasTable.h:11-30 */
typedef struct
{
    int* alias_table_alias;
    float* alias_table_probas;
    float sum_elements;
    unsigned int size;
} b2pt_alias_table;

/* DevicePackedTexturedMaterialSoA, src/HostDeviceCommon/Material/MaterialPackedSoA.h:16-120,381-565.
 * One array per packed field, indexed by material. Bit layouts: src/HostDeviceCommon/Packing.h. */
typedef struct
{
    uint8_t* flags;                                   /* UChar8BoolsPacked */
    b2pt_rgb* emission;
    uint32_t* base_color_roughness;                   /* rgb 3x8 | roughness 8 */
    float* oren_nayar_sigma;
    uint32_t* metallic_F90_and_metallic;              /* rgb | metallic */
    uint32_t* metallic_F82_packed_and_diffuse_transmission;
    float* metallic_F90_falloff_exponent;
    uint32_t* anisotropy_and_rotation_and_second_roughness; /* 4x8: anisotropy, rotation, 2nd weight, 2nd roughness */
    uint32_t* specular_color_and_tint_factor;         /* rgb | tint */
    uint32_t* specular_and_darkening_and_coat_roughness;    /* 4x8: specular, darkening, coat roughness, - */
    float* coat_medium_thickness;
    uint32_t* coat_and_medium_absorption;             /* rgb absorption | coat */
    uint32_t* coat_roughening_darkening_anisotropy_and_rotation;
    float* coat_ior;
    uint32_t* sheen_and_color;                        /* rgb | sheen */
    float* ior;
    uint32_t* absorption_color_packed;
    float* absorption_at_distance;
    uint32_t* sheen_roughness_transmission_dispersion_thin_film; /* 4x8 */
    float* dispersion_abbe_number;
    float* thin_film_ior;
    float* thin_film_thickness;
    float* thin_film_kappa_3;
    float* thin_film_base_ior_override;
    uint32_t* alpha_thin_film_hue_dielectric_priority; /* 2x8 floats | 2x8 uchars */
    uint32_t* normal_map_emission_index;              /* 2x16 texture indices each */
    uint32_t* base_color_roughness_metallic_index;
    uint32_t* roughness_and_metallic_index;
    uint32_t* anisotropic_specular_index;
    uint32_t* coat_sheen_index;
    uint32_t* specular_transmission_index;
} b2pt_materials_soa;

/* RenderBuffers, src/HostDeviceCommon/RenderBuffers.h:15-84 */
typedef struct
{
    b2pt_rgb* accumulated_ray_colors;
    uint8_t gmon_estimator_opaque[32];                /* GMoNDevice: out of scope */
    int* triangles_indices;
    b2pt_float3* vertices_positions;
    uint8_t* has_vertex_normals;
    b2pt_float3* vertex_normals;
    b2pt_float2* texcoords;
    float* triangles_areas;
    int* material_indices;
    b2pt_materials_soa materials_buffer;
    uint8_t* material_opaque;
    int emissive_triangles_count;
    int* emissive_triangles_primitive_indices;
    int* emissive_triangles_primitive_indices_and_emissive_textures;
    b2pt_alias_table emissives_power_alias_table;
    void* material_textures;
} b2pt_render_buffers;

/* BRDFsData, src/HostDeviceCommon/BSDFsData.h:24-77 */
typedef struct
{
    uint8_t white_furnace_mode;
    uint8_t white_furnace_mode_turn_off_emissives;
    void* sheen_ltc_parameters_texture;
    void* GGX_conductor_directional_albedo;
    void* glossy_dielectric_directional_albedo;
    void* GGX_glass_directional_albedo;
    void* GGX_glass_directional_albedo_inverse;
    void* GGX_thin_glass_directional_albedo;
    uint8_t use_hardware_tex_interpolation;
    int GGX_masking_shadowing;                        /* 0 = height-correlated, 1 = height-uncorrelated (default) */
    float energy_compensation_roughness_threshold;
    int glass_energy_compensation_max_bounce;
    int metal_energy_compensation_max_bounce;
    int clearcoat_energy_compensation_max_bounce;
    int glossy_base_energy_compensation_max_bounce;
    float microfacet_regularization_tau_0;
    float microfacet_regularization_min_roughness;
} b2pt_bsdfs_data;

/* AuxiliaryBuffers, src/HostDeviceCommon/RenderData.h:23-77 */
typedef struct
{
    uint8_t* pixel_active;
    b2pt_float3* denoiser_normals;
    b2pt_rgb* denoiser_albedo;
    int* pixel_sample_count;
    float* pixel_squared_luminance;
    int* pixel_converged_sample_count;
    uint8_t* still_one_ray_active;
    unsigned int* pixel_count_converged_so_far;
    void* restir_gi_reservoir_buffer_1;
    void* restir_gi_reservoir_buffer_2;
    void* restir_gi_reservoir_buffer_3;
} b2pt_aux_buffers;

/* GBufferDevice, src/Device/includes/GBufferDevice.h:17-33. materials: 108-byte DevicePackedEffectiveMaterial
 * records (src/HostDeviceCommon/Material/MaterialPacked.h:16-540); normals: Octahedral24BitNormalPadded32b. */
typedef struct
{
    void* materials;
    int* first_hit_prim_index;
    b2pt_float3* primary_hit_position;
    uint32_t* shading_normals;
    uint32_t* geometric_normals;
} b2pt_gbuffer;

/* src/HostDeviceCommon/ReSTIR/ReSTIRCommonSettings.h */
typedef struct
{
    uint8_t do_temporal_reuse_pass;
    uint8_t use_permutation_sampling;
    int permutation_sampling_random_bits;
    int max_neighbor_search_count;
    int neighbor_search_radius;
    uint8_t temporal_buffer_clear_requested;
} b2pt_restir_temporal_common;

typedef struct
{
    uint8_t do_spatial_reuse_pass;
    int spatial_pass_index;
    int number_of_passes;
    int reuse_radius;
    uint8_t auto_reuse_radius;
    int reuse_neighbor_count;
    uint8_t do_disocclusion_reuse_boost;
    int disocclusion_reuse_count;
    uint8_t debug_neighbor_location;
    int debug_neighbor_location_direction;
    uint8_t do_neighbor_rotation;
    unsigned int spatial_neighbors_rng_seed;
    uint8_t coalesced_spatial_reuse;
    uint8_t use_adaptive_directional_spatial_reuse;
    uint8_t allow_converged_neighbors_reuse;
    float converged_neighbor_reuse_probability;
    uint8_t do_visibility_only_last_pass;
    int neighbor_visibility_count;
    unsigned int* per_pixel_spatial_reuse_directions_mask_u;
    unsigned long long* per_pixel_spatial_reuse_directions_mask_ull;
    uint8_t* per_pixel_spatial_reuse_radius;
    int minimum_per_pixel_reuse_radius;
    unsigned long long current_pixel_directions_reuse_mask;
    uint8_t compute_spatial_reuse_hit_rate;
    unsigned long long* spatial_reuse_hit_rate_hits;
    unsigned long long* spatial_reuse_hit_rate_total;
} b2pt_restir_spatial_common;

typedef struct
{
    uint8_t use_normal_similarity_heuristic;
    float normal_similarity_angle_degrees;
    float normal_similarity_angle_precomp;
    uint8_t reject_using_geometric_normals;
    uint8_t use_plane_distance_heuristic;
    float plane_distance_threshold;
    uint8_t use_roughness_similarity_heuristic;
    float roughness_similarity_threshold;
} b2pt_restir_neighbor_similarity;

/* ReSTIRDISettings, src/HostDeviceCommon/ReSTIR/ReSTIRDISettings.h:14-161 */
typedef struct
{
    b2pt_restir_temporal_common common_temporal_pass;
    b2pt_restir_spatial_common common_spatial_pass;
    b2pt_restir_neighbor_similarity neighbor_similarity_settings;
    int m_cap;
    uint8_t use_confidence_weights;
    float symmetric_ratio_mis_weights_beta_exponent;

    struct
    {
        int number_of_initial_light_candidates;
        int number_of_initial_bsdf_candidates;
        float envmap_candidate_probability;
        void* output_reservoirs;
    } initial_candidates;
    struct { void* input_reservoirs; void* output_reservoirs; } temporal_pass;
    struct { void* input_reservoirs; void* output_reservoirs; } spatial_pass;
    struct
    {
        int number_of_subsets;
        int subset_size;
        int tile_size;
        void* light_samples;
    } light_presampling;

    uint8_t do_fused_spatiotemporal;
    uint8_t do_final_shading_visibility;
    void* restir_output_reservoirs;
} b2pt_restir_di_settings;

/* HIPRTRenderSettings, src/HostDeviceCommon/RenderSettings.h:26-278 */
typedef struct
{
    b2pt_int2 render_resolution;
    uint8_t need_to_reset;
    uint8_t DEBUG_gmon_auto_blending_weights;
    float DEBUG_GMON_DIVIDER;
    int DEBUG_GMON_WINDOW_SIZE;
    int DEBUG_REGIR_PRE_INTEGRATION_ITERATIONS;
    int DEBUG_REGIR_PRE_INTEGRATION_SAMPLE_COUNT_PER_RESERVOIR;
    uint8_t enable_direct;
    void* DEBUG_SUM_COUNT;
    void* DEBUG_SUM_TOTAL;
    uint8_t do_update_status_buffers;
    uint8_t accumulate;
    int denoiser_AOV_accumulation_counter;
    unsigned int sample_number;
    int output_debug_sample_N;
    int samples_per_frame;
    int nb_bounces;
    uint8_t do_russian_roulette;
    int russian_roulette_min_depth;
    float russian_roulette_throughput_clamp;
    int path_russian_roulette_method;                 /* 0 = MAX_THROUGHPUT, 1 = ARNOLD_2014 */
    int freeze_random;
    uint8_t display_NaNs;
    uint8_t allow_render_low_resolution;
    uint8_t wants_render_low_resolution;
    int render_low_resolution_scaling;
    uint8_t enable_adaptive_sampling;
    int adaptive_sampling_min_samples;
    float adaptive_sampling_noise_threshold;
    uint8_t use_pixel_stop_noise_threshold;
    float stop_pixel_percentage_converged;
    float stop_pixel_noise_threshold;
    float direct_contribution_clamp;
    float envmap_contribution_clamp;
    float indirect_contribution_clamp;
    float minimum_light_contribution;
    uint8_t do_alpha_testing;
    int alpha_testing_indirect_bounce;
    uint8_t do_normal_mapping;
    struct { int number_of_light_candidates; int number_of_bsdf_candidates; } ris_settings;
    b2pt_restir_di_settings restir_di_settings;
    uint64_t restir_gi_settings_opaque[264 / 8];      /* ReSTIR GI: out of scope */
    uint64_t regir_settings_opaque[648 / 8];          /* ReGIR: out of scope */
} b2pt_render_settings;

/* WorldSettings, src/HostDeviceCommon/WorldSettings.h:21-53. ambient_light_type: 0 NONE, 1 UNIFORM, 2 ENVMAP */
typedef struct
{
    int ambient_light_type;
    b2pt_rgb uniform_light_color;
    unsigned int envmap_width, envmap_height;
    float envmap_intensity;
    int envmap_scale_background_intensity;
    uint32_t* envmap;                                 /* RGBE9995Packed texels, Packing.h:441-485 */
    float envmap_total_sum;
    float* envmap_cdf;
    b2pt_alias_table envmap_alias_table;
    float envmap_to_world_matrix[3][3];
    float world_to_envmap_matrix[3][3];
} b2pt_world_settings;

/* HIPRTCamera, src/HostDeviceCommon/HIPRTCamera.h:11-53 */
typedef struct
{
    float inverse_view[4][4];
    float inverse_projection[4][4];
    float view_projection[4][4];
    b2pt_float3 position;
    float vertical_fov;
    int sensor_width, sensor_height;
    uint8_t do_jittering;
} b2pt_camera;

/* HIPRTRenderData, src/HostDeviceCommon/RenderData.h:93-129 */
typedef struct
{
    unsigned int random_number;
    void* GPU_BVH;
    void* light_GPU_BVH;
    void* hiprt_function_table;                       /* unused: the filter is compiled into the traversal */
    int global_traversal_stack_buffer_size;
    uint64_t global_traversal_stack_buffer_opaque[2]; /* unused: the traversal stack lives in shared memory */
    b2pt_render_buffers buffers;
    b2pt_bsdfs_data bsdfs_data;
    b2pt_aux_buffers aux_buffers;
    b2pt_gbuffer g_buffer;
    b2pt_gbuffer g_buffer_prev_frame;
    b2pt_render_settings render_settings;
    b2pt_world_settings world_settings;
    uint64_t nee_plus_plus_opaque[88 / 8];            /* NEE++: out of scope */
    b2pt_camera current_camera;
    b2pt_camera prev_camera;
    void* cpu_only[2];
} b2pt_render_data;

/* A LUT "texture" handle as stored in b2pt_bsdfs_data: linear fp32, 1 channel used (3 for the sheen LTC table). */
typedef struct
{
    const float* texels;
    int width, height, depth, channels;
} b2pt_lut;

#ifdef __cplusplus
}
static_assert(sizeof(b2pt_render_data) == 2784, "b2pt_render_data must be 2784 bytes like HIPRTRenderData");
#endif

#endif
