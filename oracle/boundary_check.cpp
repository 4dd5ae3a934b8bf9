/*
 * oracle/boundary_check.cpp -- TEST INFRASTRUCTURE ONLY. The drop-in boundary, shown rather than asserted (SURVEY.md 8f rank 1,
 * minimum slice): a host program compiled against the REFERENCE's own headers -- HostDeviceCommon/RenderData.h, not
 * include/b2pt_render_data.h -- fills a HIPRTRenderData the way the reference's host does (scene packing, material SoA, alias
 * table, settings: the oracle harness, i.e. reference code), moves every buffer to the GPU through libb2pt.so's DEVICE LAYER
 * (b2pt_malloc / b2pt_memcpy_h2d = OrochiBuffer, b2pt_lut_create = OrochiTexture(3D), b2pt_bvh_build = HIPRTGeometry::build_bvh),
 * looks the kernels up by the reference's macro names (b2pt_kernel_get = GPUKernel::compile with GPUKernelCompilerOptions) and
 * launches them with the reference's launch signature, b2pt_kernel_launch(block_x, block_y, width, height, { &render_data })
 * (GPUKernel::launch, src/Compiler/GPUKernel.cpp:151-165), in CPURenderer's pass order. The framebuffer it reads back must be
 * bit-identical to the one the reference's CPU path renders from the very same HIPRTRenderData.
 *
 *   boundary_check_<variant> --scene X.b2scene --luts L.b2luts --width W --height H --spp N --bounces B
 */
#define main ref_harness_main
#include "ref_harness.cpp"
#undef main

#include "b2pt.h"

#ifdef B2_BOUNDARY_GPUKERNEL
/* boundary_gpukernel_<variant>: the same program with every kernel held, "compiled" and launched as a reference GPUKernel object
 * (src/Compiler/GPUKernel.h, bound to this library by hiprt-path-tracer_b200/binding/GPUKernel_b2pt.cpp) and the macros set through the
 * reference's GPUKernelCompilerOptions, the way GPURenderer and its render passes do (GPURenderer.cpp:482-501,
 * MegaKernelRenderPass.cpp:20-60, ReSTIRDIRenderPass.cpp:21-140) */
#include "Compiler/GPUKernel.h"
#endif

#define CHECK(call)                                                                                     \
    do                                                                                                  \
    {                                                                                                   \
        int rc_ = (call);                                                                               \
        if (rc_ != B2PT_OK)                                                                             \
        {                                                                                               \
            std::fprintf(stderr, "%s failed (%d): %s\n", #call, rc_, b2pt_last_error_string());         \
            std::exit(1);                                                                               \
        }                                                                                               \
    } while (0)

static b2pt_ctx* g_ctx = nullptr;

/* OrochiBuffer<T>::upload_data: device allocation + copy; the host pointer in the render data is replaced by the device one */
template <class T>
static void to_device(T*& pointer, size_t count)
{
    if (!pointer || count == 0)
    {
        pointer = nullptr;
        return;
    }
    void* d = nullptr;
    CHECK(b2pt_malloc(g_ctx, count * sizeof(T), &d));
    CHECK(b2pt_memcpy_h2d(g_ctx, d, pointer, count * sizeof(T), nullptr));
    pointer = reinterpret_cast<T*>(d);
}

template <class T>
static T* device_zeros(size_t count)
{
    void* d = nullptr;
    CHECK(b2pt_malloc(g_ctx, count * sizeof(T), &d));
    CHECK(b2pt_memset(g_ctx, d, 0, count * sizeof(T), nullptr));
    return reinterpret_cast<T*>(d);
}

static void* lut(const LutFile& luts, const char* name)
{
    const LutFile::Entry& e = luts.entries.at(name);
    void* handle = nullptr;
    CHECK(b2pt_lut_create(g_ctx, e.data.data(), e.w, e.h, e.d, 1, &handle));
    return handle;
}

int main(int argc, char** argv)
{
    Args args = parse_args(argc, argv);
    if (args.threads > 0)
        omp_set_num_threads(args.threads);
    if (args.scene.empty() || args.luts.empty())
    {
        std::fprintf(stderr, "usage: %s --scene X.b2scene --luts brdf_luts.b2luts --width W --height H --spp N --bounces B\n", argv[0]);
        return 2;
    }
    static_assert(sizeof(HIPRTRenderData) == sizeof(b2pt_render_data), "the reference's render data and the ABI's mirror differ in size");

    /* ---- the reference side: one Oracle renders on the CPU, a second one only PREPARES the same render data ---------------- */
    auto prepare = [&](Oracle& o) {
        o.rd.render_settings.nb_bounces = args.bounces;
        o.rd.render_settings.samples_per_frame = args.spp;
        o.rd.render_settings.output_debug_sample_N = args.spp - 1;
        o.rd.render_settings.restir_di_settings.do_fused_spatiotemporal = args.fused;
        o.setup_brdfs_data(args.luts.c_str());
        o.set_envmap(args.envmap.c_str());
        o.set_scene(args.scene.c_str());
    };
    Oracle cpu(args);
    prepare(cpu);
    cpu.render();

    Oracle host(args);
    prepare(host);

    /* ---- the device side, through the C ABI only ------------------------------------------------------------------------------ */
    CHECK(b2pt_ctx_create(0, &g_ctx));
    b2pt_stream stream = nullptr;
    CHECK(b2pt_stream_create(g_ctx, &stream));

    HIPRTRenderData rd = host.rd; /* by-value copy: what GPURenderer::update_render_data assembles every frame */
    const size_t n = (size_t)args.width * args.height;
    const size_t n_vert = host.scene.n_vertices(), n_tri = host.scene.n_triangles(), n_mat = host.scene.n_materials(), n_emissive = host.emissive_prims.size();

    /* scene buffers (GPURenderer::set_scene, GPURenderer.cpp:769-939) */
    to_device(rd.buffers.triangles_indices, 3 * n_tri);
    to_device(rd.buffers.vertices_positions, n_vert);
    to_device(rd.buffers.vertex_normals, n_vert);
    to_device(rd.buffers.texcoords, n_vert);
    to_device(rd.buffers.has_vertex_normals, n_vert);
    to_device(rd.buffers.material_indices, n_tri);
    to_device(rd.buffers.material_opaque, n_mat);
    to_device(rd.buffers.triangles_areas, n_tri);
    to_device(rd.buffers.emissive_triangles_primitive_indices, n_emissive);
    rd.buffers.emissive_triangles_primitive_indices_and_emissive_textures = rd.buffers.emissive_triangles_primitive_indices;
    to_device(rd.buffers.emissives_power_alias_table.alias_table_alias, n_emissive);
    to_device(rd.buffers.emissives_power_alias_table.alias_table_probas, n_emissive);
    {
        auto& m = rd.buffers.materials_buffer; /* DevicePackedTexturedMaterialSoA: one array per packed field */
        to_device(m.flags, n_mat);
        to_device(m.emission, n_mat);
        to_device(m.base_color_roughness, n_mat);
        to_device(m.oren_nayar_sigma, n_mat);
        to_device(m.metallic_F90_and_metallic, n_mat);
        to_device(m.metallic_F82_packed_and_diffuse_transmission, n_mat);
        to_device(m.metallic_F90_falloff_exponent, n_mat);
        to_device(m.anisotropy_and_rotation_and_second_roughness, n_mat);
        to_device(m.specular_color_and_tint_factor, n_mat);
        to_device(m.specular_and_darkening_and_coat_roughness, n_mat);
        to_device(m.coat_medium_thickness, n_mat);
        to_device(m.coat_and_medium_absorption, n_mat);
        to_device(m.coat_roughening_darkening_anisotropy_and_rotation, n_mat);
        to_device(m.coat_ior, n_mat);
        to_device(m.sheen_and_color, n_mat);
        to_device(m.ior, n_mat);
        to_device(m.absorption_color_packed, n_mat);
        to_device(m.absorption_at_distance, n_mat);
        to_device(m.sheen_roughness_transmission_dispersion_thin_film, n_mat);
        to_device(m.dispersion_abbe_number, n_mat);
        to_device(m.thin_film_ior, n_mat);
        to_device(m.thin_film_thickness, n_mat);
        to_device(m.thin_film_kappa_3, n_mat);
        to_device(m.thin_film_base_ior_override, n_mat);
        to_device(m.alpha_thin_film_hue_dielectric_priority, n_mat);
        to_device(m.normal_map_emission_index, n_mat);
        to_device(m.base_color_roughness_metallic_index, n_mat);
        to_device(m.roughness_and_metallic_index, n_mat);
        to_device(m.anisotropic_specular_index, n_mat);
        to_device(m.coat_sheen_index, n_mat);
        to_device(m.specular_transmission_index, n_mat);
    }
    if (rd.world_settings.ambient_light_type == AmbientLightType::ENVMAP)
    {
        const size_t texels = (size_t)rd.world_settings.envmap_width * rd.world_settings.envmap_height;
        to_device(rd.world_settings.envmap, texels);
        to_device(rd.world_settings.envmap_alias_table.alias_table_alias, texels);
        to_device(rd.world_settings.envmap_alias_table.alias_table_probas, texels);
    }

    /* per-pixel buffers (GPURenderer's OrochiBuffers / the interop framebuffer) */
    rd.buffers.accumulated_ray_colors = device_zeros<ColorRGB32F>(n);
    rd.aux_buffers.pixel_active = device_zeros<unsigned char>(n);
    rd.aux_buffers.denoiser_albedo = device_zeros<ColorRGB32F>(n);
    rd.aux_buffers.denoiser_normals = device_zeros<float3>(n);
    rd.aux_buffers.pixel_sample_count = device_zeros<int>(n);
    rd.aux_buffers.pixel_converged_sample_count = device_zeros<int>(n);
    rd.aux_buffers.pixel_squared_luminance = device_zeros<float>(n);
    unsigned char* status = device_zeros<unsigned char>(16);
    rd.aux_buffers.still_one_ray_active = status;
    rd.aux_buffers.pixel_count_converged_so_far = reinterpret_cast<AtomicType<unsigned int>*>(status + 4);
    for (GBufferDevice* g : { &rd.g_buffer, &rd.g_buffer_prev_frame })
    {
        g->materials = device_zeros<DevicePackedEffectiveMaterial>(n);
        g->geometric_normals = device_zeros<Octahedral24BitNormalPadded32b>(n);
        g->shading_normals = device_zeros<Octahedral24BitNormalPadded32b>(n);
        g->primary_hit_position = device_zeros<float3>(n);
        g->first_hit_prim_index = device_zeros<int>(n);
    }
    /* ReSTIR DI reservoir buffers: same byte size as the reference's (sizeof(ReSTIRDIReservoir) per pixel), default-constructed
     * reservoirs (the ABI documents its SoA plane layout: 9 planes of 4-byte words, emissive triangle index = plane 6 = -1) */
    ReSTIRDIReservoir* reservoirs[3] = { nullptr, nullptr, nullptr };
#if DirectLightSamplingStrategy == LSS_RESTIR_DI
    static_assert(sizeof(ReSTIRDIReservoir) == 36, "reservoir size");
    for (int i = 0; i < 3; i++)
    {
        reservoirs[i] = device_zeros<ReSTIRDIReservoir>(n);
        CHECK(b2pt_memset(g_ctx, reinterpret_cast<char*>(reservoirs[i]) + 6 * n * 4, 0xFF, n * 4, nullptr));
    }
    rd.render_settings.restir_di_settings.common_spatial_pass.per_pixel_spatial_reuse_directions_mask_ull = device_zeros<unsigned long long int>(n);
    rd.render_settings.restir_di_settings.common_spatial_pass.per_pixel_spatial_reuse_radius = device_zeros<unsigned char>(n);
    rd.render_settings.restir_di_settings.common_spatial_pass.per_pixel_spatial_reuse_directions_mask_u = nullptr;
    rd.render_settings.restir_di_settings.common_spatial_pass.spatial_reuse_hit_rate_total = nullptr;
    rd.render_settings.restir_di_settings.common_spatial_pass.spatial_reuse_hit_rate_hits = nullptr;
    rd.render_settings.restir_di_settings.light_presampling.light_samples = nullptr;
    rd.render_settings.restir_di_settings.initial_candidates.output_reservoirs = reservoirs[0];
    rd.render_settings.restir_di_settings.restir_output_reservoirs = reservoirs[1];
#endif

    /* BRDF LUTs: OrochiTexture / OrochiTexture3D handles in BRDFsData (GPURenderer.cpp:188-243) */
    LutFile luts;
    if (!luts.load(args.luts.c_str())) { std::fprintf(stderr, "cannot read LUT blob\n"); return 1; }
    {
        void* handle = nullptr;
        CHECK(b2pt_lut_create(g_ctx, reinterpret_cast<const float*>(ltc_parameters_table_approximation.data()), 32, 32, 1, 3, &handle));
        rd.bsdfs_data.sheen_ltc_parameters_texture = handle;
    }
    rd.bsdfs_data.GGX_conductor_directional_albedo = lut(luts, "ggx_conductor");
    rd.bsdfs_data.glossy_dielectric_directional_albedo = lut(luts, "glossy_dielectric");
    rd.bsdfs_data.GGX_glass_directional_albedo = lut(luts, "ggx_glass");
    rd.bsdfs_data.GGX_glass_directional_albedo_inverse = lut(luts, "ggx_glass_inv");
    rd.bsdfs_data.GGX_thin_glass_directional_albedo = lut(luts, "ggx_thin_glass");

    /* acceleration structure: HIPRTGeometry::build_bvh -> hiprtGeometry handle in the render data */
    b2pt_bvh* bvh = nullptr;
    CHECK(b2pt_bvh_build(g_ctx, host.scene.indices.data(), (int)n_tri, host.scene.positions.data(), (int)n_vert, 0, nullptr, &bvh));
    rd.GPU_BVH = reinterpret_cast<hiprtGeometry>(b2pt_bvh_device_handle(bvh));
    rd.light_GPU_BVH = nullptr;
    rd.cpu_only = decltype(rd.cpu_only)();

#ifdef B2_BOUNDARY_GPUKERNEL
    /* the renderer's global compiler options, shared with every kernel (GPURenderer::m_global_compiler_options); the values of this
     * build's strategy macros are set the way the ImGui panels set them */
    auto hiprt_orochi_ctx = std::make_shared<HIPRTOrochiCtx>();
    hiprt_orochi_ctx->orochi_ctx = reinterpret_cast<oroCtx>(g_ctx); /* the binding's convention: the context slot carries the b2pt_ctx */
    auto global_options = std::make_shared<GPUKernelCompilerOptions>();
    global_options->set_macro_value(GPUKernelCompilerOptions::BSDF_OVERRIDE, (int)BSDFOverride);
    global_options->set_macro_value(GPUKernelCompilerOptions::DIRECT_LIGHT_SAMPLING_STRATEGY, (int)DirectLightSamplingStrategy);
    global_options->set_macro_value(GPUKernelCompilerOptions::DIRECT_LIGHT_SAMPLING_BASE_STRATEGY, (int)DirectLightSamplingBaseStrategy);
    global_options->set_macro_value(GPUKernelCompilerOptions::RESTIR_DI_LATER_BOUNCES_SAMPLING_STRATEGY, (int)ReSTIR_DI_LaterBouncesSamplingStrategy);
    global_options->set_macro_value(GPUKernelCompilerOptions::ENVMAP_SAMPLING_STRATEGY, (int)EnvmapSamplingStrategy);
    global_options->set_macro_value(GPUKernelCompilerOptions::RESTIR_DI_BIAS_CORRECTION_WEIGHTS, (int)ReSTIR_DI_BiasCorrectionWeights);
    global_options->set_macro_value(GPUKernelCompilerOptions::RESTIR_DI_DO_LIGHT_PRESAMPLING, (int)ReSTIR_DI_DoLightPresampling);
    GPUKernel camera_rays_kernel(std::string(DEVICE_KERNELS_DIRECTORY) + "/CameraRays.h", "CameraRays");
    GPUKernel megakernel_kernel(std::string(DEVICE_KERNELS_DIRECTORY) + "/Megakernel.h", "MegaKernel");
    GPUKernel initial_candidates_kernel(std::string(DEVICE_KERNELS_DIRECTORY) + "/ReSTIR/DI/InitialCandidates.h", "ReSTIR_DI_InitialCandidates");
    GPUKernel spatiotemporal_kernel(std::string(DEVICE_KERNELS_DIRECTORY) + "/ReSTIR/DI/FusedSpatiotemporalReuse.h", "ReSTIR_DI_SpatiotemporalReuse");
    GPUKernel temporal_kernel(std::string(DEVICE_KERNELS_DIRECTORY) + "/ReSTIR/DI/TemporalReuse.h", "ReSTIR_DI_TemporalReuse");
    GPUKernel spatial_kernel(std::string(DEVICE_KERNELS_DIRECTORY) + "/ReSTIR/DI/SpatialReuse.h", "ReSTIR_DI_SpatialReuse");
    auto prepare_kernel = [&](GPUKernel& k) {
        k.synchronize_options_with(global_options);
        k.compile(hiprt_orochi_ctx, {}, /* use_cache */ true, /* silent */ false);
        if (!k.has_been_compiled())
            std::exit(1);
    };
    GPUKernel *camera_rays = &camera_rays_kernel, *megakernel = &megakernel_kernel, *initial_candidates = nullptr, *spatiotemporal = nullptr, *temporal = nullptr, *spatial = nullptr;
    prepare_kernel(camera_rays_kernel);
    prepare_kernel(megakernel_kernel);
#if DirectLightSamplingStrategy == LSS_RESTIR_DI
    initial_candidates = &initial_candidates_kernel;
    spatiotemporal = &spatiotemporal_kernel;
    temporal = &temporal_kernel;
    spatial = &spatial_kernel;
    prepare_kernel(initial_candidates_kernel);
    prepare_kernel(spatiotemporal_kernel);
    prepare_kernel(temporal_kernel);
    prepare_kernel(spatial_kernel);
#endif
    Xorshift32Generator rng(args.seed);
    void* launch_args[] = { &rd };
    auto launch = [&](GPUKernel* k) { k->launch_asynchronous(8, 8, args.width, args.height, launch_args, reinterpret_cast<oroStream_t>(stream)); };
#else
    /* kernels, by the reference's macro names (GPUKernelCompilerOptions.cpp:19-105) */
    b2pt_option options[] = {
        { "BSDFOverride", (int)BSDFOverride }, { "DirectLightSamplingStrategy", (int)DirectLightSamplingStrategy },
        { "DirectLightSamplingBaseStrategy", (int)DirectLightSamplingBaseStrategy },
        { "ReSTIR_DI_LaterBouncesSamplingStrategy", (int)ReSTIR_DI_LaterBouncesSamplingStrategy }, { "EnvmapSamplingStrategy", (int)EnvmapSamplingStrategy },
        { "PrincipledBSDFDoEnergyCompensation", (int)PrincipledBSDFDoEnergyCompensation }, { "NestedDielectricsStackSize", (int)NestedDielectricsStackSize },
        { "ReSTIR_DI_BiasCorrectionWeights", (int)ReSTIR_DI_BiasCorrectionWeights }, { "ReSTIR_DI_DoVisibilityReuse", (int)ReSTIR_DI_DoVisibilityReuse },
    };
    const int option_count = (int)(sizeof(options) / sizeof(options[0]));
    b2pt_kernel *camera_rays = nullptr, *megakernel = nullptr, *initial_candidates = nullptr, *spatiotemporal = nullptr, *temporal = nullptr, *spatial = nullptr;
    CHECK(b2pt_kernel_get(g_ctx, "CameraRays", options, option_count, &camera_rays));
    CHECK(b2pt_kernel_get(g_ctx, "MegaKernel", options, option_count, &megakernel));
#if DirectLightSamplingStrategy == LSS_RESTIR_DI
    CHECK(b2pt_kernel_get(g_ctx, "ReSTIR_DI_InitialCandidates", options, option_count, &initial_candidates));
    CHECK(b2pt_kernel_get(g_ctx, "ReSTIR_DI_SpatiotemporalReuse", options, option_count, &spatiotemporal));
    CHECK(b2pt_kernel_get(g_ctx, "ReSTIR_DI_TemporalReuse", options, option_count, &temporal));
    CHECK(b2pt_kernel_get(g_ctx, "ReSTIR_DI_SpatialReuse", options, option_count, &spatial));
#endif

    /* the frame loop of CPURenderer::render (CPURenderer.cpp:534-604), passes launched like GPUKernel::launch */
    Xorshift32Generator rng(args.seed);
    void* launch_args[] = { &rd };
    auto launch = [&](b2pt_kernel* k) { CHECK(b2pt_kernel_launch(k, 8, 8, args.width, args.height, launch_args, stream)); };
#endif
    bool odd_frame = false;
    for (int frame = 1; frame <= args.spp; frame++)
    {
        rd.render_settings.do_update_status_buffers = true;
        CHECK(b2pt_memset(g_ctx, status, 0, 16, stream));
        rd.prev_camera = rd.current_camera;
        rd.current_camera = host.rd.current_camera;
        rd.random_number = rng.xorshift32();
        launch(camera_rays);
#if DirectLightSamplingStrategy == LSS_RESTIR_DI
        {
            /* ReSTIRDIRenderPass::launch, fused or split (CPURenderer.cpp:860-1032) */
            ReSTIRDISettings& di = rd.render_settings.restir_di_settings;
            rd.random_number = rng.xorshift32();
            di.initial_candidates.output_reservoirs = reservoirs[0];
            launch(initial_candidates);
            if (di.do_fused_spatiotemporal)
            {
                rd.random_number = rng.xorshift32();
                di.common_temporal_pass.permutation_sampling_random_bits = rng.xorshift32();
                di.temporal_pass.input_reservoirs = odd_frame ? reservoirs[1] : reservoirs[2];
                di.temporal_pass.output_reservoirs = nullptr;
                di.spatial_pass.input_reservoirs = di.temporal_pass.input_reservoirs;
                di.spatial_pass.output_reservoirs = odd_frame ? reservoirs[2] : reservoirs[1];
                launch(spatiotemporal);
            }
            else
            {
                rd.random_number = rng.xorshift32();
                di.common_temporal_pass.permutation_sampling_random_bits = rng.xorshift32();
                di.temporal_pass.input_reservoirs = di.restir_output_reservoirs;
                di.temporal_pass.output_reservoirs = reservoirs[0];
                launch(temporal);
                rd.random_number = rng.xorshift32();
                di.spatial_pass.input_reservoirs = di.temporal_pass.output_reservoirs;
                di.spatial_pass.output_reservoirs = reservoirs[1];
                launch(spatial);
            }
            di.restir_output_reservoirs = di.spatial_pass.output_reservoirs;
            odd_frame = !odd_frame;
        }
#endif
        rd.random_number = rng.xorshift32();
        launch(megakernel);
        if (rd.render_settings.accumulate)
            rd.render_settings.sample_number++;
        rd.random_number = rng.xorshift32();
        rd.render_settings.need_to_reset = false;
    }
    CHECK(b2pt_stream_synchronize(g_ctx, stream));

    std::vector<float> device_framebuffer(n * 3);
    CHECK(b2pt_memcpy_d2h(g_ctx, device_framebuffer.data(), rd.buffers.accumulated_ray_colors, n * 12, nullptr));
    const float* cpu_framebuffer = cpu.framebuffer.data().data();
    size_t differing = 0;
    for (size_t i = 0; i < n; i++)
        if (std::memcmp(&device_framebuffer[3 * i], &cpu_framebuffer[3 * i], 12) != 0)
            differing++;
    double mean = 0.0;
    for (size_t i = 0; i < n * 3; i++)
        mean += device_framebuffer[i];
#ifdef B2_BOUNDARY_GPUKERNEL
    if (megakernel->compute_execution_time() <= 0.0f)
    {
        std::fprintf(stderr, "GPUKernel::compute_execution_time reports no time for a kernel that ran\n");
        return 1;
    }
    std::printf("{\"through_gpukernel_class\": true, \"megakernel_last_ms\": %.4f}\n", megakernel->get_last_execution_time());
#endif
    std::printf("{\"boundary_check\": true, \"render_data_bytes\": %zu, \"width\": %d, \"height\": %d, \"spp\": %d, \"pixels\": %zu, \"pixels_not_bit_identical\": %zu, "
                "\"mean\": %.6g, \"direct_light_strategy\": %d}\n",
                sizeof(HIPRTRenderData), args.width, args.height, args.spp, n, differing, mean / (n * 3), (int)DirectLightSamplingStrategy);
    return differing == 0 ? 0 : 1;
}
