/*
 * oracle/layout_check.cpp -- BUILD-TIME CHECK (compiled with -fsyntax-only by oracle/Makefile, only where
 * /root/reference exists). Static-asserts that include/b2pt_render_data.h is byte-for-byte layout compatible
 * with the reference's HIPRTRenderData and the sub-structs the hot path reads, using the reference's own headers.
 */
#include "HostDeviceCommon/RenderData.h"
#include "b2pt_render_data.h"

#include <cstddef>

#define SAME_SIZE(REF, MINE) static_assert(sizeof(REF) == sizeof(MINE), "size mismatch: " #REF " vs " #MINE)
#define SAME_OFF(REF, RF, MINE, MF) static_assert(offsetof(REF, RF) == offsetof(MINE, MF), "offset mismatch: " #REF "::" #RF)
#define SAME(REF, MINE, F) SAME_OFF(REF, F, MINE, F)

SAME_SIZE(HIPRTRenderData, b2pt_render_data);
SAME(HIPRTRenderData, b2pt_render_data, random_number);
SAME(HIPRTRenderData, b2pt_render_data, GPU_BVH);
SAME(HIPRTRenderData, b2pt_render_data, light_GPU_BVH);
SAME(HIPRTRenderData, b2pt_render_data, hiprt_function_table);
SAME(HIPRTRenderData, b2pt_render_data, global_traversal_stack_buffer_size);
SAME_OFF(HIPRTRenderData, global_traversal_stack_buffer, b2pt_render_data, global_traversal_stack_buffer_opaque);
SAME(HIPRTRenderData, b2pt_render_data, buffers);
SAME(HIPRTRenderData, b2pt_render_data, bsdfs_data);
SAME(HIPRTRenderData, b2pt_render_data, aux_buffers);
SAME(HIPRTRenderData, b2pt_render_data, g_buffer);
SAME(HIPRTRenderData, b2pt_render_data, g_buffer_prev_frame);
SAME(HIPRTRenderData, b2pt_render_data, render_settings);
SAME(HIPRTRenderData, b2pt_render_data, world_settings);
SAME_OFF(HIPRTRenderData, nee_plus_plus, b2pt_render_data, nee_plus_plus_opaque);
SAME(HIPRTRenderData, b2pt_render_data, current_camera);
SAME(HIPRTRenderData, b2pt_render_data, prev_camera);
SAME(HIPRTRenderData, b2pt_render_data, cpu_only);

SAME_SIZE(RenderBuffers, b2pt_render_buffers);
SAME(RenderBuffers, b2pt_render_buffers, accumulated_ray_colors);
SAME_OFF(RenderBuffers, gmon_estimator, b2pt_render_buffers, gmon_estimator_opaque);
SAME(RenderBuffers, b2pt_render_buffers, triangles_indices);
SAME(RenderBuffers, b2pt_render_buffers, vertices_positions);
SAME(RenderBuffers, b2pt_render_buffers, has_vertex_normals);
SAME(RenderBuffers, b2pt_render_buffers, vertex_normals);
SAME(RenderBuffers, b2pt_render_buffers, texcoords);
SAME(RenderBuffers, b2pt_render_buffers, triangles_areas);
SAME(RenderBuffers, b2pt_render_buffers, material_indices);
SAME(RenderBuffers, b2pt_render_buffers, materials_buffer);
SAME(RenderBuffers, b2pt_render_buffers, material_opaque);
SAME(RenderBuffers, b2pt_render_buffers, emissive_triangles_count);
SAME(RenderBuffers, b2pt_render_buffers, emissive_triangles_primitive_indices);
SAME(RenderBuffers, b2pt_render_buffers, emissive_triangles_primitive_indices_and_emissive_textures);
SAME(RenderBuffers, b2pt_render_buffers, emissives_power_alias_table);
SAME(RenderBuffers, b2pt_render_buffers, material_textures);

SAME_SIZE(DeviceAliasTable, b2pt_alias_table);
SAME(DeviceAliasTable, b2pt_alias_table, alias_table_alias);
SAME(DeviceAliasTable, b2pt_alias_table, alias_table_probas);
SAME(DeviceAliasTable, b2pt_alias_table, sum_elements);
SAME(DeviceAliasTable, b2pt_alias_table, size);

SAME_SIZE(DevicePackedTexturedMaterialSoA, b2pt_materials_soa);
#define MAT(F) SAME(DevicePackedTexturedMaterialSoA, b2pt_materials_soa, F)
MAT(flags); MAT(emission); MAT(base_color_roughness); MAT(oren_nayar_sigma); MAT(metallic_F90_and_metallic);
MAT(metallic_F82_packed_and_diffuse_transmission); MAT(metallic_F90_falloff_exponent);
MAT(anisotropy_and_rotation_and_second_roughness); MAT(specular_color_and_tint_factor);
MAT(specular_and_darkening_and_coat_roughness); MAT(coat_medium_thickness); MAT(coat_and_medium_absorption);
MAT(coat_roughening_darkening_anisotropy_and_rotation); MAT(coat_ior); MAT(sheen_and_color); MAT(ior);
MAT(absorption_color_packed); MAT(absorption_at_distance); MAT(sheen_roughness_transmission_dispersion_thin_film);
MAT(dispersion_abbe_number); MAT(thin_film_ior); MAT(thin_film_thickness); MAT(thin_film_kappa_3);
MAT(thin_film_base_ior_override); MAT(alpha_thin_film_hue_dielectric_priority); MAT(normal_map_emission_index);
MAT(base_color_roughness_metallic_index); MAT(roughness_and_metallic_index); MAT(anisotropic_specular_index);
MAT(coat_sheen_index); MAT(specular_transmission_index);
static_assert(sizeof(DevicePackedEffectiveMaterial) == 108, "packed G-buffer material record must be 108 bytes");
static_assert(sizeof(Octahedral24BitNormalPadded32b) == 4, "packed normal must be 4 bytes");
static_assert(sizeof(ColorRGB24bFloat0_1Packed) == 4 && sizeof(Float4xPacked) == 4 && sizeof(Float2xUChar2xPacked) == 4 && sizeof(Uint2xPacked) == 4 && sizeof(UChar8BoolsPacked) == 1, "packed field sizes");

SAME_SIZE(BRDFsData, b2pt_bsdfs_data);
#define BS(F) SAME(BRDFsData, b2pt_bsdfs_data, F)
BS(white_furnace_mode); BS(white_furnace_mode_turn_off_emissives); BS(sheen_ltc_parameters_texture);
BS(GGX_conductor_directional_albedo); BS(glossy_dielectric_directional_albedo); BS(GGX_glass_directional_albedo);
BS(GGX_glass_directional_albedo_inverse); BS(GGX_thin_glass_directional_albedo); BS(use_hardware_tex_interpolation);
BS(GGX_masking_shadowing); BS(energy_compensation_roughness_threshold); BS(glass_energy_compensation_max_bounce);
BS(metal_energy_compensation_max_bounce); BS(clearcoat_energy_compensation_max_bounce); BS(glossy_base_energy_compensation_max_bounce);
SAME_OFF(BRDFsData, microfacet_regularization, b2pt_bsdfs_data, microfacet_regularization_tau_0);
static_assert(offsetof(MicrofacetRegularizationSettings, min_roughness) == 4, "regularization settings layout");

SAME_SIZE(AuxiliaryBuffers, b2pt_aux_buffers);
#define AUX(F) SAME(AuxiliaryBuffers, b2pt_aux_buffers, F)
AUX(pixel_active); AUX(denoiser_normals); AUX(denoiser_albedo); AUX(pixel_sample_count); AUX(pixel_squared_luminance);
AUX(pixel_converged_sample_count); AUX(still_one_ray_active); AUX(pixel_count_converged_so_far);
AUX(restir_gi_reservoir_buffer_1); AUX(restir_gi_reservoir_buffer_2); AUX(restir_gi_reservoir_buffer_3);

SAME_SIZE(GBufferDevice, b2pt_gbuffer);
SAME(GBufferDevice, b2pt_gbuffer, materials); SAME(GBufferDevice, b2pt_gbuffer, first_hit_prim_index);
SAME(GBufferDevice, b2pt_gbuffer, primary_hit_position); SAME(GBufferDevice, b2pt_gbuffer, shading_normals);
SAME(GBufferDevice, b2pt_gbuffer, geometric_normals);

SAME_SIZE(HIPRTRenderSettings, b2pt_render_settings);
#define RS(F) SAME(HIPRTRenderSettings, b2pt_render_settings, F)
RS(render_resolution); RS(need_to_reset); RS(DEBUG_gmon_auto_blending_weights); RS(DEBUG_GMON_DIVIDER); RS(DEBUG_GMON_WINDOW_SIZE);
RS(DEBUG_REGIR_PRE_INTEGRATION_ITERATIONS); RS(DEBUG_REGIR_PRE_INTEGRATION_SAMPLE_COUNT_PER_RESERVOIR); RS(enable_direct);
RS(DEBUG_SUM_COUNT); RS(DEBUG_SUM_TOTAL); RS(do_update_status_buffers); RS(accumulate); RS(denoiser_AOV_accumulation_counter);
RS(sample_number); RS(output_debug_sample_N); RS(samples_per_frame); RS(nb_bounces); RS(do_russian_roulette);
RS(russian_roulette_min_depth); RS(russian_roulette_throughput_clamp); RS(path_russian_roulette_method); RS(freeze_random);
RS(display_NaNs); RS(allow_render_low_resolution); RS(wants_render_low_resolution); RS(render_low_resolution_scaling);
RS(enable_adaptive_sampling); RS(adaptive_sampling_min_samples); RS(adaptive_sampling_noise_threshold);
RS(use_pixel_stop_noise_threshold); RS(stop_pixel_percentage_converged); RS(stop_pixel_noise_threshold);
RS(direct_contribution_clamp); RS(envmap_contribution_clamp); RS(indirect_contribution_clamp); RS(minimum_light_contribution);
RS(do_alpha_testing); RS(alpha_testing_indirect_bounce); RS(do_normal_mapping); RS(ris_settings); RS(restir_di_settings);
SAME_OFF(HIPRTRenderSettings, restir_gi_settings, b2pt_render_settings, restir_gi_settings_opaque);
SAME_OFF(HIPRTRenderSettings, regir_settings, b2pt_render_settings, regir_settings_opaque);
static_assert(sizeof(ReSTIRGISettings) == 264 && sizeof(ReGIRSettings) == 648 && sizeof(NEEPlusPlusDevice) == 88 && sizeof(GMoNDevice) == 32, "opaque block sizes");
static_assert(offsetof(RISSettings, number_of_bsdf_candidates) == 4, "RIS settings layout");

SAME_SIZE(ReSTIRDISettings, b2pt_restir_di_settings);
#define DI(F) SAME(ReSTIRDISettings, b2pt_restir_di_settings, F)
DI(common_temporal_pass); DI(common_spatial_pass); DI(neighbor_similarity_settings); DI(m_cap); DI(use_confidence_weights);
DI(symmetric_ratio_mis_weights_beta_exponent); DI(initial_candidates); DI(temporal_pass); DI(spatial_pass); DI(light_presampling);
DI(do_fused_spatiotemporal); DI(do_final_shading_visibility); DI(restir_output_reservoirs);
static_assert(offsetof(ReSTIRDIInitialCandidatesSettings, output_reservoirs) == 16 && offsetof(ReSTIRDIInitialCandidatesSettings, envmap_candidate_probability) == 8, "initial candidates layout");
static_assert(offsetof(ReSTIRDILightPresamplingSettings, light_samples) == 16, "presampling layout");

SAME_SIZE(ReSTIRCommonTemporalPassSettings, b2pt_restir_temporal_common);
#define TP(F) SAME(ReSTIRCommonTemporalPassSettings, b2pt_restir_temporal_common, F)
TP(do_temporal_reuse_pass); TP(use_permutation_sampling); TP(permutation_sampling_random_bits); TP(max_neighbor_search_count);
TP(neighbor_search_radius); TP(temporal_buffer_clear_requested);

SAME_SIZE(ReSTIRCommonSpatialPassSettings, b2pt_restir_spatial_common);
#define SP(F) SAME(ReSTIRCommonSpatialPassSettings, b2pt_restir_spatial_common, F)
SP(do_spatial_reuse_pass); SP(spatial_pass_index); SP(number_of_passes); SP(reuse_radius); SP(auto_reuse_radius);
SP(reuse_neighbor_count); SP(do_disocclusion_reuse_boost); SP(disocclusion_reuse_count); SP(debug_neighbor_location);
SP(debug_neighbor_location_direction); SP(do_neighbor_rotation); SP(spatial_neighbors_rng_seed); SP(coalesced_spatial_reuse);
SP(use_adaptive_directional_spatial_reuse); SP(allow_converged_neighbors_reuse); SP(converged_neighbor_reuse_probability);
SP(do_visibility_only_last_pass); SP(neighbor_visibility_count); SP(per_pixel_spatial_reuse_directions_mask_u);
SP(per_pixel_spatial_reuse_directions_mask_ull); SP(per_pixel_spatial_reuse_radius); SP(minimum_per_pixel_reuse_radius);
SP(current_pixel_directions_reuse_mask); SP(compute_spatial_reuse_hit_rate); SP(spatial_reuse_hit_rate_hits); SP(spatial_reuse_hit_rate_total);

SAME_SIZE(ReSTIRCommonNeighborSimiliaritySettings, b2pt_restir_neighbor_similarity);
#define NS(F) SAME(ReSTIRCommonNeighborSimiliaritySettings, b2pt_restir_neighbor_similarity, F)
NS(use_normal_similarity_heuristic); NS(normal_similarity_angle_degrees); NS(normal_similarity_angle_precomp);
NS(reject_using_geometric_normals); NS(use_plane_distance_heuristic); NS(plane_distance_threshold);
NS(use_roughness_similarity_heuristic); NS(roughness_similarity_threshold);

SAME_SIZE(WorldSettings, b2pt_world_settings);
#define WS(F) SAME(WorldSettings, b2pt_world_settings, F)
WS(ambient_light_type); WS(uniform_light_color); WS(envmap_width); WS(envmap_height); WS(envmap_intensity);
WS(envmap_scale_background_intensity); WS(envmap); WS(envmap_total_sum); WS(envmap_cdf); WS(envmap_alias_table);
WS(envmap_to_world_matrix); WS(world_to_envmap_matrix);
static_assert((int)AmbientLightType::UNIFORM == 1 && (int)AmbientLightType::ENVMAP == 2, "ambient light enum values");

SAME_SIZE(HIPRTCamera, b2pt_camera);
#define CAM(F) SAME(HIPRTCamera, b2pt_camera, F)
CAM(inverse_view); CAM(inverse_projection); CAM(view_projection); CAM(position); CAM(vertical_fov); CAM(sensor_width);
CAM(sensor_height); CAM(do_jittering);

/* ReSTIR DI reservoir record: 36 bytes = 9 four-byte words (the SoA plane count used by the kernels) */
static_assert(sizeof(ReSTIRDIReservoir) == 36 && offsetof(ReSTIRDIReservoir, sample) == 12 && offsetof(ReSTIRDISample, flags) == 20, "reservoir layout");

/* b2_restir_di.h reads ReSTIR GI's common_spatial_pass.coalesced_spatial_reuse (the seed quirk of the DI spatial
 * kernel) at the DI offset of the same field inside the opaque GI block: both structs start with ReSTIRCommonSettings */
static_assert(offsetof(ReSTIRGISettings, common_spatial_pass) + offsetof(ReSTIRCommonSpatialPassSettings, coalesced_spatial_reuse) ==
                  offsetof(b2pt_restir_di_settings, common_spatial_pass) + offsetof(b2pt_restir_spatial_common, coalesced_spatial_reuse),
              "ReSTIR GI coalesced_spatial_reuse offset");

int main() { return 0; }
