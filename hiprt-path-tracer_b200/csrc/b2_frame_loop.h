/*
 * b2_frame_loop.h -- host-side sample loop and ReSTIR DI buffer choreography, independent of where the passes run.
 *
 * This is the B200 host mirror of what drives the kernels in the reference:
 *   GPURendererThread::render_path_tracing / RenderGraph::launch_async   src/Renderer/GPURendererThread.cpp:311-361
 *   FillGBufferRenderPass / MegaKernelRenderPass::launch_async          src/Renderer/RenderPasses/*.cpp
 *   ReSTIRDIRenderPass::launch_async + configure_*                      src/Renderer/RenderPasses/ReSTIRDIRenderPass.cpp:343-676
 * with ONE deliberate choice: the order and number of host RNG draws is the CPURenderer's
 * (src/Renderer/CPURenderer.cpp:534-604,698-705,797-1068), because parity is defined against
 * the CPU path: camera(1) -> [presample(1)] -> initial(1) -> fused: 2 | temporal: 2 then 1 per spatial pass
 * -> megakernel(1) -> spare(1).
 * The caller supplies `launch(pass)`, which runs one pass over all pixels with the current `rd` (CUDA kernels in
 * b2_renderer.cpp; an OpenMP pixel loop in oracle/port_harness.cpp).
 */
#ifndef B2_FRAME_LOOP_H
#define B2_FRAME_LOOP_H

#include "b2_math.h"
#include "b2pt_render_data.h"

#include <cstring>

enum B2Pass
{
    B2_PASS_CAMERA_RAYS = 0,
    B2_PASS_RESTIR_DI_PRESAMPLING,
    B2_PASS_RESTIR_DIRECTIONAL_REUSE,
    B2_PASS_RESTIR_DI_INITIAL_CANDIDATES,
    B2_PASS_RESTIR_DI_TEMPORAL,
    B2_PASS_RESTIR_DI_SPATIAL,
    B2_PASS_RESTIR_DI_SPATIOTEMPORAL,
    B2_PASS_MEGAKERNEL,
    B2_PASS_COUNT
};

/* Reference defaults of every field the hot path reads (RenderSettings.h:26-230, ReSTIRDISettings.h:60-140,
 * RISSettings.h:11-15, WorldSettings.h:21-53, BSDFsData.h:24-77, HIPRTCamera.h:11-25). */
inline void b2pt_render_data_set_defaults(b2pt_render_data& rd)
{
    std::memset(&rd, 0, sizeof(rd));
    rd.random_number = 42;
    rd.global_traversal_stack_buffer_size = 32;

    b2pt_bsdfs_data& b = rd.bsdfs_data;
    b.white_furnace_mode = 0;
    b.white_furnace_mode_turn_off_emissives = 1;
    b.use_hardware_tex_interpolation = 0;
    b.GGX_masking_shadowing = 1; /* HeightUncorrelated */
    b.energy_compensation_roughness_threshold = 0.0f;
    b.glass_energy_compensation_max_bounce = -1;
    b.metal_energy_compensation_max_bounce = -1;
    b.clearcoat_energy_compensation_max_bounce = -1;
    b.glossy_base_energy_compensation_max_bounce = -1;
    b.microfacet_regularization_tau_0 = 30.0f;
    b.microfacet_regularization_min_roughness = 0.0f;

    b2pt_render_settings& s = rd.render_settings;
    s.render_resolution.x = 1280;
    s.render_resolution.y = 720;
    s.need_to_reset = 1;
    s.DEBUG_GMON_DIVIDER = 3.0f;
    s.DEBUG_GMON_WINDOW_SIZE = 3;
    s.DEBUG_REGIR_PRE_INTEGRATION_ITERATIONS = 4;
    s.DEBUG_REGIR_PRE_INTEGRATION_SAMPLE_COUNT_PER_RESERVOIR = 32;
    s.enable_direct = 1;
    s.do_update_status_buffers = 0;
    s.accumulate = 1;
    s.sample_number = 0;
    s.output_debug_sample_N = 1;
    s.samples_per_frame = 1;
    s.nb_bounces = 0;
    s.do_russian_roulette = 1;
    s.russian_roulette_min_depth = 0; /* min(5, nb_bounces / 2) frozen at construction with nb_bounces = 0 */
    s.russian_roulette_throughput_clamp = 10.0f;
    s.path_russian_roulette_method = 0;
    s.freeze_random = 0;
    s.display_NaNs = 1;
    s.render_low_resolution_scaling = 2;
    s.enable_adaptive_sampling = 0;
    s.adaptive_sampling_min_samples = 96;
    s.adaptive_sampling_noise_threshold = 0.075f;
    s.use_pixel_stop_noise_threshold = 1;
    s.stop_pixel_percentage_converged = 70.0f;
    s.stop_pixel_noise_threshold = 0.075f;
    s.do_alpha_testing = 1;
    s.alpha_testing_indirect_bounce = 1; /* nb_bounces + 1 frozen at construction */
    s.do_normal_mapping = 1;
    s.ris_settings.number_of_light_candidates = 4;
    s.ris_settings.number_of_bsdf_candidates = 1;

    b2pt_restir_di_settings& di = s.restir_di_settings;
    di.common_temporal_pass.do_temporal_reuse_pass = 1;
    di.common_temporal_pass.use_permutation_sampling = 0;
    di.common_temporal_pass.permutation_sampling_random_bits = 42;
    di.common_temporal_pass.max_neighbor_search_count = 8;
    di.common_temporal_pass.neighbor_search_radius = 4;
    di.common_spatial_pass.do_spatial_reuse_pass = 1;
    di.common_spatial_pass.number_of_passes = 1;
    di.common_spatial_pass.reuse_radius = 16;
    di.common_spatial_pass.auto_reuse_radius = 1;
    di.common_spatial_pass.reuse_neighbor_count = 5;
    di.common_spatial_pass.disocclusion_reuse_count = 5;
    di.common_spatial_pass.spatial_neighbors_rng_seed = 42;
    di.common_spatial_pass.coalesced_spatial_reuse = 1;
    di.common_spatial_pass.converged_neighbor_reuse_probability = 0.5f;
    di.common_spatial_pass.do_visibility_only_last_pass = 1;
    di.common_spatial_pass.neighbor_visibility_count = 5;
    di.common_spatial_pass.minimum_per_pixel_reuse_radius = 3;
    di.neighbor_similarity_settings.use_normal_similarity_heuristic = 1;
    di.neighbor_similarity_settings.normal_similarity_angle_degrees = 37.5f;
    di.neighbor_similarity_settings.normal_similarity_angle_precomp = 0.906307787f;
    di.neighbor_similarity_settings.reject_using_geometric_normals = 1;
    di.neighbor_similarity_settings.use_plane_distance_heuristic = 1;
    di.neighbor_similarity_settings.plane_distance_threshold = 0.1f;
    di.neighbor_similarity_settings.use_roughness_similarity_heuristic = 0;
    di.neighbor_similarity_settings.roughness_similarity_threshold = 0.25f;
    di.m_cap = 3;
    di.use_confidence_weights = 1;
    di.symmetric_ratio_mis_weights_beta_exponent = 2.0f;
    di.initial_candidates.number_of_initial_light_candidates = 4;
    di.initial_candidates.number_of_initial_bsdf_candidates = 1;
    di.initial_candidates.envmap_candidate_probability = 0.5f;
    di.light_presampling.number_of_subsets = 128;
    di.light_presampling.subset_size = 1024;
    di.light_presampling.tile_size = 8;
    di.do_fused_spatiotemporal = 0;
    di.do_final_shading_visibility = 1;

    b2pt_world_settings& w = rd.world_settings;
    w.ambient_light_type = 1; /* UNIFORM */
    w.envmap_intensity = 1.0f;
    for (int i = 0; i < 3; i++)
        w.envmap_to_world_matrix[i][i] = w.world_to_envmap_matrix[i][i] = 1.0f;

    rd.current_camera.do_jittering = 1;
    rd.prev_camera.do_jittering = 1;
}

struct B2FrameLoop
{
    b2pt_render_data rd;
    Rng host_rng;
    bool restir_di = false;          /* DirectLightSamplingStrategy == LSS_RESTIR_DI */
    bool restir_di_presampling = false;
    bool odd_frame = false;
    void* initial_reservoirs = nullptr;
    void* spatial_reservoirs_1 = nullptr;
    void* spatial_reservoirs_2 = nullptr;
    b2pt_camera camera;              /* what Camera::to_hiprt() returns every sample (static camera) */

    B2FrameLoop()
    {
        b2pt_render_data_set_defaults(rd);
        host_rng = make_rng(42);
        std::memset(&camera, 0, sizeof(camera));
    }

    /* one sample of every pixel, in CPURenderer::render order (CPURenderer.cpp:541-566) */
    template <class Launch>
    void render_one_sample(Launch&& launch)
    {
        rd.render_settings.do_update_status_buffers = 1;
        /* pre_render_update (:574-584) resets the two status words: done by the caller-provided pass 'launch'
         * implementations before B2_PASS_CAMERA_RAYS, since where they live (host / device) is theirs to know */
        rd.prev_camera = rd.current_camera;
        rd.current_camera = camera;

        rd.random_number = host_rng.next_u32();
        launch(B2_PASS_CAMERA_RAYS);

        if (restir_di)
            restir_di_pass(launch);

        rd.random_number = host_rng.next_u32();
        launch(B2_PASS_MEGAKERNEL);

        if (rd.render_settings.accumulate)
            rd.render_settings.sample_number++;
        rd.random_number = host_rng.next_u32();
        rd.render_settings.need_to_reset = 0;
    }

    template <class Launch>
    void restir_di_pass(Launch&& launch)
    {
        b2pt_restir_di_settings& di = rd.render_settings.restir_di_settings;
        if (restir_di_presampling)
        {
            rd.random_number = host_rng.next_u32();
            launch(B2_PASS_RESTIR_DI_PRESAMPLING);
        }

        /* ReSTIRDIRenderPass::compute_optimal_spatial_reuse_radii (ReSTIRDIRenderPass.cpp:391-415): a prepass at sample 0 of an
         * accumulating, non-interactive render; no host RNG draw. (The CPURenderer defines the equivalent function but never
         * calls it for ReSTIR DI, CPURenderer.cpp:797-818,863-872; the test oracle calls it at this point on request.) */
        if (di.common_spatial_pass.use_adaptive_directional_spatial_reuse && rd.render_settings.accumulate && rd.render_settings.sample_number == 0 &&
            !rd.render_settings.wants_render_low_resolution)
            launch(B2_PASS_RESTIR_DIRECTIONAL_REUSE);

        rd.random_number = host_rng.next_u32();
        di.initial_candidates.output_reservoirs = initial_reservoirs;
        launch(B2_PASS_RESTIR_DI_INITIAL_CANDIDATES);

        if (di.do_fused_spatiotemporal)
        {
            rd.random_number = host_rng.next_u32();
            di.common_temporal_pass.permutation_sampling_random_bits = (int)host_rng.next_u32();
            di.temporal_pass.input_reservoirs = odd_frame ? spatial_reservoirs_1 : spatial_reservoirs_2;
            di.temporal_pass.output_reservoirs = nullptr;
            di.spatial_pass.input_reservoirs = di.temporal_pass.input_reservoirs;
            di.spatial_pass.output_reservoirs = odd_frame ? spatial_reservoirs_2 : spatial_reservoirs_1;
            launch(B2_PASS_RESTIR_DI_SPATIOTEMPORAL);
        }
        else
        {
            if (di.common_temporal_pass.do_temporal_reuse_pass)
            {
                rd.random_number = host_rng.next_u32();
                di.common_temporal_pass.permutation_sampling_random_bits = (int)host_rng.next_u32();
                di.temporal_pass.input_reservoirs = di.restir_output_reservoirs;
                if (di.common_spatial_pass.do_spatial_reuse_pass)
                    di.temporal_pass.output_reservoirs = initial_reservoirs;
                else
                    di.temporal_pass.output_reservoirs = odd_frame ? spatial_reservoirs_1 : spatial_reservoirs_2;
                launch(B2_PASS_RESTIR_DI_TEMPORAL);
            }
            if (di.common_spatial_pass.do_spatial_reuse_pass)
            {
                for (int pass = 0; pass < di.common_spatial_pass.number_of_passes; pass++)
                {
                    rd.random_number = host_rng.next_u32();
                    if (pass == 0)
                    {
                        di.spatial_pass.input_reservoirs = di.common_temporal_pass.do_temporal_reuse_pass ? di.temporal_pass.output_reservoirs : di.initial_candidates.output_reservoirs;
                        di.spatial_pass.output_reservoirs = spatial_reservoirs_1;
                    }
                    else if ((pass & 1) == 0)
                    {
                        di.spatial_pass.input_reservoirs = spatial_reservoirs_2;
                        di.spatial_pass.output_reservoirs = spatial_reservoirs_1;
                    }
                    else
                    {
                        di.spatial_pass.input_reservoirs = spatial_reservoirs_1;
                        di.spatial_pass.output_reservoirs = spatial_reservoirs_2;
                    }
                    launch(B2_PASS_RESTIR_DI_SPATIAL);
                }
            }
        }

        if (di.common_spatial_pass.do_spatial_reuse_pass)
            di.restir_output_reservoirs = di.spatial_pass.output_reservoirs;
        else if (di.common_temporal_pass.do_temporal_reuse_pass)
            di.restir_output_reservoirs = di.temporal_pass.output_reservoirs;
        else
            di.restir_output_reservoirs = di.initial_candidates.output_reservoirs;
        odd_frame = !odd_frame;
    }
};

#endif
