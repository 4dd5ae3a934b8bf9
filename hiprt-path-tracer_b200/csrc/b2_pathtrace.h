/*
 * b2_pathtrace.h -- the two per-pixel passes every configuration runs: G-buffer fill ("CameraRays") and the
 * bounce loop ("MegaKernel"), written as per-pixel functions so the CUDA kernels (b2_kernels.cu) decide the
 * thread <-> pixel mapping (persistent blocks, tile order) independently of the math.
 *
 * Restates /root/reference/src/Device/kernels/CameraRays.h:20-213, kernels/Megakernel.h:22-144,
 * includes/PathTracing.h:18-217, RussianRoulette.h:20-55, AdaptiveSampling.h:18-102, SanityCheck.h:30-105,
 * LightSampling/Lights.h:228-423 and HostDeviceCommon/HIPRTCamera.h:31-52.
 */
#ifndef B2_PATHTRACE_H
#define B2_PATHTRACE_H

#include "b2_envmap.h"
#include "b2_principled.h"
#include "b2_ris.h"
#include "b2_restir_di_shading.h"
#include "b2_nee.h"

#if defined(__CUDA_ARCH__)
#define B2_ATOMIC_ADD_U32(ptr, v) atomicAdd((ptr), (v))
#else
#define B2_ATOMIC_ADD_U32(ptr, v) __atomic_fetch_add((ptr), (v), __ATOMIC_RELAXED)
#endif

B2_HD bool has_access_to_adaptive_sampling_buffers(const b2pt_render_settings& s)
{
    bool has_access = false;
    has_access |= s.stop_pixel_noise_threshold > 0.0f;
    has_access |= s.enable_adaptive_sampling != 0;
    has_access &= s.accumulate != 0;
    return has_access;
}

template <class V>
B2_HD bool use_prev_frame_g_buffer(const b2pt_render_settings& s)
{
    return V::DIRECT == B2_LSS_RESTIR_DI && s.restir_di_settings.common_temporal_pass.do_temporal_reuse_pass;
}


B2_HD void store_packed_material(void* buffer, uint32_t pixel_index, const PackedMat& p)
{
    uint32_t* dst = reinterpret_cast<uint32_t*>(buffer) + (size_t)pixel_index * PM_WORDS;
#pragma unroll
    for (int i = 0; i < PM_WORDS; i++)
        dst[i] = p.w[i];
}
B2_HD PackedMat load_packed_material(const void* buffer, uint32_t pixel_index)
{
    const uint32_t* src = reinterpret_cast<const uint32_t*>(buffer) + (size_t)pixel_index * PM_WORDS;
    PackedMat p;
#pragma unroll
    for (int i = 0; i < PM_WORDS; i++)
        p.w[i] = src[i];
    return p;
}

/* CameraRays.h:20-59 */
template <class V>
B2_HD void reset_render(const b2pt_render_data& rd, uint32_t pixel_index)
{
    if (has_access_to_adaptive_sampling_buffers(rd.render_settings))
    {
        rd.aux_buffers.pixel_sample_count[pixel_index] = 0;
        rd.aux_buffers.pixel_squared_luminance[pixel_index] = 0;
        rd.aux_buffers.pixel_converged_sample_count[pixel_index] = -1;
    }
    const uint32_t zero_normal = pack_normal_oct24(v3(0.0f));
    const PackedMat default_material = pack_material<V::BSDF_OVERRIDE>(mat_defaults());
    rd.g_buffer.first_hit_prim_index[pixel_index] = -1;
    rd.g_buffer.geometric_normals[pixel_index] = zero_normal;
    rd.g_buffer.shading_normals[pixel_index] = zero_normal;
    rd.g_buffer.primary_hit_position[pixel_index] = b2pt_float3{ 0.0f, 0.0f, 0.0f };
    store_packed_material(rd.g_buffer.materials, pixel_index, default_material);
    if (use_prev_frame_g_buffer<V>(rd.render_settings))
    {
        rd.g_buffer_prev_frame.first_hit_prim_index[pixel_index] = -1;
        rd.g_buffer_prev_frame.geometric_normals[pixel_index] = zero_normal;
        rd.g_buffer_prev_frame.shading_normals[pixel_index] = zero_normal;
        rd.g_buffer_prev_frame.primary_hit_position[pixel_index] = b2pt_float3{ 0.0f, 0.0f, 0.0f };
        store_packed_material(rd.g_buffer_prev_frame.materials, pixel_index, default_material);
    }
}

/* AdaptiveSampling.h:18-102 */
B2_HD float pixel_confidence_interval(const b2pt_render_data& rd, int pixel_index, int pixel_sample_count, float& average_luminance)
{
    float lum = luminance(load_rgb(rd.buffers.accumulated_ray_colors, pixel_index));
    average_luminance = lum / (pixel_sample_count + 1);
    float squared_luminance = rd.aux_buffers.pixel_squared_luminance[pixel_index];
    float pixel_variance = (squared_luminance - lum * average_luminance) / (pixel_sample_count + 1);
    return 1.96f * sqrtf(pixel_variance) / sqrtf((float)(pixel_sample_count + 1));
}

B2_HD bool adaptive_sampling(const b2pt_render_data& rd, int pixel_index, bool& pixel_converged)
{
    const b2pt_render_settings& s = rd.render_settings;
    const b2pt_aux_buffers& aux = rd.aux_buffers;
    if (!has_access_to_adaptive_sampling_buffers(s))
        return true;

    if (s.enable_adaptive_sampling)
    {
        if (aux.pixel_converged_sample_count[pixel_index] != -1)
            return false;
        int pixel_sample_count = aux.pixel_sample_count[pixel_index];
        if (pixel_sample_count > s.adaptive_sampling_min_samples)
        {
            float average_luminance;
            float confidence_interval = pixel_confidence_interval(rd, pixel_index, pixel_sample_count, average_luminance);
            bool pixel_needs_sampling = confidence_interval > s.adaptive_sampling_noise_threshold * average_luminance;
            if (!pixel_needs_sampling)
            {
                if (aux.pixel_converged_sample_count[pixel_index] == -1)
                    aux.pixel_converged_sample_count[pixel_index] = pixel_sample_count;
                return false;
            }
        }
        return true;
    }
    else if (s.stop_pixel_noise_threshold > 0.0f && s.use_pixel_stop_noise_threshold)
    {
        int pixel_sample_count = aux.pixel_sample_count[pixel_index];
        float average_luminance;
        float confidence_interval = pixel_confidence_interval(rd, pixel_index, pixel_sample_count, average_luminance);
        pixel_converged = (confidence_interval <= s.stop_pixel_noise_threshold * average_luminance) && (s.sample_number > 1);
        int current_converged_count = aux.pixel_converged_sample_count[pixel_index];
        if (pixel_converged && current_converged_count == -1)
            aux.pixel_converged_sample_count[pixel_index] = pixel_sample_count;
        else if (!pixel_converged)
            aux.pixel_converged_sample_count[pixel_index] = -1;
    }
    return true;
}

/* HIPRTCamera::get_camera_ray, HIPRTCamera.h:31-52 */
B2_HD void camera_ray(const b2pt_camera& cam, float x, float y, b2pt_int2 res, V3& origin, V3& direction)
{
    float x_ndc = x / res.x * 2 - 1;
    float y_ndc = y / res.y * 2 - 1;
    origin = mat4_mul_point(cam.inverse_view, v3(0.0f, 0.0f, 0.0f));
    V3 point_view_space = mat4_mul_point(cam.inverse_projection, v3(x_ndc, y_ndc, -1.0f));
    V3 point_world_space = mat4_mul_point(cam.inverse_view, point_view_space);
    direction = normalize(point_world_space - origin);
}

B2_HD uint32_t pixel_seed(const b2pt_render_data& rd, uint32_t pixel_index)
{
    if (rd.render_settings.freeze_random)
        return wang_hash(pixel_index + 1);
    return wang_hash((pixel_index + 1) * (rd.render_settings.sample_number + 1) * rd.random_number);
}

/* rescale_samples, CameraRays.h:61-80 (GMoN part out of scope) */
B2_HD void rescale_samples(const b2pt_render_data& rd, uint32_t pixel_index)
{
    float n = (float)rd.render_settings.sample_number;
    Rgb c = load_rgb(rd.buffers.accumulated_ray_colors, pixel_index) / n * (float)(rd.render_settings.sample_number + 1);
    rd.buffers.accumulated_ray_colors[pixel_index] = b2pt_rgb{ c.r, c.g, c.b };
}

/* The CameraRays pass for pixel (x, y): CameraRays.h:82-213 */
template <class V>
B2_HD void camera_rays_pixel(const Ctx& ctx, uint32_t x, uint32_t y)
{
    const b2pt_render_data& rd = *ctx.rd;
    const b2pt_int2 res = rd.render_settings.render_resolution;
    if (x >= (uint32_t)res.x || y >= (uint32_t)res.y)
        return;
    uint32_t pixel_index = x + y * res.x;

    if (rd.render_settings.need_to_reset)
        reset_render<V>(rd, pixel_index);

    if (do_render_low_resolution(rd.render_settings))
    {
        int res_scaling = rd.render_settings.render_low_resolution_scaling;
        if (x % res_scaling != 0 || y % res_scaling != 0)
        {
            rd.aux_buffers.pixel_active[pixel_index] = false;
            return;
        }
        pixel_index /= res_scaling;
    }

    if (use_prev_frame_g_buffer<V>(rd.render_settings))
    {
        rd.g_buffer_prev_frame.geometric_normals[pixel_index] = rd.g_buffer.geometric_normals[pixel_index];
        rd.g_buffer_prev_frame.shading_normals[pixel_index] = rd.g_buffer.shading_normals[pixel_index];
        store_packed_material(rd.g_buffer_prev_frame.materials, pixel_index, load_packed_material(rd.g_buffer.materials, pixel_index));
        rd.g_buffer_prev_frame.primary_hit_position[pixel_index] = rd.g_buffer.primary_hit_position[pixel_index];
        rd.g_buffer_prev_frame.first_hit_prim_index[pixel_index] = rd.g_buffer.first_hit_prim_index[pixel_index];
    }

    bool pixel_converged = false;
    bool sampling_needed = adaptive_sampling(rd, pixel_index, pixel_converged);
    if (pixel_converged || !sampling_needed)
    {
        if (rd.render_settings.do_update_status_buffers)
            B2_ATOMIC_ADD_U32(rd.aux_buffers.pixel_count_converged_so_far, 1u);
    }

    if (has_access_to_adaptive_sampling_buffers(rd.render_settings))
    {
        if (!sampling_needed)
        {
            rescale_samples(rd, pixel_index);
            rd.aux_buffers.pixel_active[pixel_index] = false;
            return;
        }
        else
            rd.aux_buffers.pixel_sample_count[pixel_index]++;
    }

    Rng rng = make_rng(pixel_seed(rd, pixel_index));

    float x_ray_point = (x + 0.5f);
    float y_ray_point = (y + 0.5f);
    if (rd.current_camera.do_jittering)
    {
        x_ray_point += rng.next_float() - 0.5f;
        y_ray_point += rng.next_float() - 0.5f;
    }

    V3 origin, direction;
    camera_ray(rd.current_camera, x_ray_point, y_ray_point, res, origin, direction);

    Payload payload;
    payload.init();
    HitInfo hit = hit_info_defaults();
    bool intersection_found = trace_main_path_ray<V>(ctx, origin, direction, payload, hit, -1, rng);

    if (intersection_found)
    {
        rd.g_buffer.geometric_normals[pixel_index] = pack_normal_oct24(hit.geometric_normal);
        rd.g_buffer.shading_normals[pixel_index] = pack_normal_oct24(hit.shading_normal);
        store_packed_material(rd.g_buffer.materials, pixel_index, pack_material<V::BSDF_OVERRIDE>(payload.material));
        rd.g_buffer.primary_hit_position[pixel_index] = b2pt_float3{ hit.inter_point.x, hit.inter_point.y, hit.inter_point.z };
    }
    else
    {
        V3 p = origin + direction;
        rd.g_buffer.primary_hit_position[pixel_index] = b2pt_float3{ p.x, p.y, p.z };
    }
    rd.g_buffer.first_hit_prim_index[pixel_index] = intersection_found ? hit.primitive_index : -1;
    rd.aux_buffers.pixel_active[pixel_index] = true;

    if (rd.render_settings.do_update_status_buffers)
        rd.aux_buffers.still_one_ray_active[0] = 1;
}

/* ---------------------------------------------------------------------------------------------------------- */

/* RussianRoulette.h:20-47 */
B2_HD bool do_russian_roulette(const b2pt_render_settings& s, int bounce, Rgb& ray_throughput, Rgb current_weight, Rng& rng)
{
    if (bounce >= s.russian_roulette_min_depth && s.do_russian_roulette)
    {
        float survive_probability = 0.0f;
        if (s.path_russian_roulette_method == 0)
            survive_probability = max_component(ray_throughput);
        else if (s.path_russian_roulette_method == 1)
        {
            survive_probability = max_component(ray_throughput * current_weight) / max_component(ray_throughput);
            survive_probability = sqrtf(survive_probability);
        }
        survive_probability = b2_min(survive_probability, 1.0f);
        if (rng.next_float() > survive_probability)
            return false;
        float throughput_scaling = 1.0f / survive_probability;
        if (s.russian_roulette_throughput_clamp > 0.0f)
            throughput_scaling = b2_min(throughput_scaling, s.russian_roulette_throughput_clamp);
        ray_throughput *= throughput_scaling;
    }
    return true;
}

/* PathTracing.h:44-62 */
B2_HD Rgb update_ray_throughput(const b2pt_render_data& rd, Payload& payload, const HitInfo& hit, Rgb current_throughput, Rgb bsdf_color, V3 bounce_direction, float bsdf_pdf_value, Rng& rng)
{
    Rgb throughput_attenuation = bsdf_color * fabsf(dot(bounce_direction, hit.shading_normal)) / bsdf_pdf_value;
    if (!do_russian_roulette(rd.render_settings, payload.bounce, current_throughput, throughput_attenuation, rng))
        return rgb(0.0f);
    current_throughput *= get_dispersion_ray_color(payload.volume_state.sampled_wavelength, payload.material.dispersion_scale);
    current_throughput *= throughput_attenuation;
    current_throughput = rgb_max(current_throughput, rgb(1.0e-5f, 1.0e-5f, 1.0e-5f));
    payload.next_ray_state = B2_RAY_BOUNCE;
    return current_throughput;
}

/* PathTracing.h:24-39,81-101 */
template <class V>
B2_HD bool compute_next_indirect_bounce(const Ctx& ctx, Payload& payload, HitInfo& hit, V3 view_direction, V3& out_origin, V3& out_direction, Rng& rng)
{
    const b2pt_render_data& rd = *ctx.rd;
    Rgb bsdf_color;
    V3 bounce_direction;
    float bsdf_pdf_value;
    int sampled_light_info = B2_LOBE_NO_INFO;
    BsdfCtx bctx = make_bsdf_ctx(view_direction, hit.shading_normal, hit.geometric_normal, v3(0.0f), &sampled_light_info, &payload.volume_state, true,
                                 &payload.material, payload.bounce, payload.accumulated_roughness);
    bsdf_color = bsdf_sample<V>(rd, bctx, bounce_direction, bsdf_pdf_value, rng);
    payload.accumulate_roughness(sampled_light_info);

    if (bsdf_pdf_value <= 0.0f)
        return false;

    payload.throughput = update_ray_throughput(rd, payload, hit, payload.throughput, bsdf_color, bounce_direction, bsdf_pdf_value, rng);
    if (is_black(payload.throughput))
        return false;

    out_origin = hit.inter_point;
    out_direction = bounce_direction;
    return true;
}

/* Lights.h:228-352: strategy dispatch for emissive geometry */
template <class V>
B2_HD Rgb sample_emissive_geometry(const Ctx& ctx, Payload& payload, const HitInfo& hit, V3 view_direction, Rng& rng, int x, int y)
{
    const b2pt_render_data& rd = *ctx.rd;
    if (rd.buffers.emissive_triangles_count == 0 && !(rd.world_settings.ambient_light_type == 2 && V::DIRECT == B2_LSS_RESTIR_DI))
        return rgb(0.0f);
    if (rd.bsdfs_data.white_furnace_mode && rd.bsdfs_data.white_furnace_mode_turn_off_emissives)
        return rgb(0.0f);

    Rgb material_self_textured_emission = rgb(0.0f);
    if (payload.material.emissive_texture_used)
        material_self_textured_emission = payload.material.emission;

    Rgb direct = rgb(0.0f);
    /* DirectLightSamplingNEESampleCount = 1: "sum of one / 1" keeps the value bit-identical */
    if constexpr (V::DIRECT == B2_LSS_ONE_LIGHT)
        direct = sample_one_light_no_MIS<V>(ctx, payload, hit, view_direction, rng) / 1.0f;
    else if constexpr (V::DIRECT == B2_LSS_BSDF)
        direct = sample_one_light_bsdf<V>(ctx, payload, hit, view_direction, rng) / 1.0f;
    else if constexpr (V::DIRECT == B2_LSS_MIS)
        direct = sample_one_light_MIS<V>(ctx, payload, hit, view_direction, rng) / 1.0f;
    else if constexpr (V::DIRECT == B2_LSS_RIS)
        direct = sample_lights_RIS<V>(ctx, payload, hit, view_direction, rng) / 1.0f;
    else if constexpr (V::DIRECT == B2_LSS_RESTIR_DI)
    {
        if (payload.bounce == 0)
            direct = sample_light_ReSTIR_DI<V>(ctx, payload, hit, view_direction, rng, x, y);
        else
        {
            if constexpr (V::LATER == B2_LATER_ONE_LIGHT)
                direct = sample_one_light_no_MIS<V>(ctx, payload, hit, view_direction, rng);
            else if constexpr (V::LATER == B2_LATER_BSDF)
                direct = sample_one_light_bsdf<V>(ctx, payload, hit, view_direction, rng);
            else if constexpr (V::LATER == B2_LATER_MIS)
                direct = sample_one_light_MIS<V>(ctx, payload, hit, view_direction, rng);
            else
                direct = sample_lights_RIS<V>(ctx, payload, hit, view_direction, rng);
            direct /= 1.0f;
        }
    }
    return direct + material_self_textured_emission;
}

/* Lights.h:364-423 */
template <class V>
B2_HD Rgb estimate_direct_lighting_from(const Ctx& ctx, Payload& payload, HitInfo& hit, V3 view_direction, Rng& rng, Rgb emissive_contribution);

template <class V>
B2_HD Rgb estimate_direct_lighting(const Ctx& ctx, Payload& payload, HitInfo& hit, V3 view_direction, int x, int y, Rng& rng)
{
    Rgb emissive_contribution = sample_emissive_geometry<V>(ctx, payload, hit, view_direction, rng, x, y);
    return estimate_direct_lighting_from<V>(ctx, payload, hit, view_direction, rng, emissive_contribution);
}

/* the arithmetic tail of estimate_direct_lighting (Lights.h:398-423), given both contributions and the vertex's state */
template <class V>
B2_HD Rgb direct_lighting_combine(const b2pt_render_data& rd, Rgb emissive_contribution, Rgb envmap_contribution, Rgb throughput, int bounce, Rgb material_emission)
{
    Rgb total = rgb(0.0f);
    emissive_contribution = clamp_light_contribution(emissive_contribution, rd.render_settings.direct_contribution_clamp, bounce == 0);
    envmap_contribution = clamp_light_contribution(envmap_contribution, rd.render_settings.envmap_contribution_clamp, bounce == 0);

    if constexpr (V::DIRECT == B2_LSS_NONE)
    {
        Rgb hit_emission = clamp_light_contribution(material_emission, rd.render_settings.indirect_contribution_clamp, bounce > 0);
        total += hit_emission * throughput;
    }
    else
    {
        if (bounce == 0)
            total += material_emission;
        total += (emissive_contribution + envmap_contribution) * throughput;
    }
    return clamp_light_contribution(total, rd.render_settings.indirect_contribution_clamp, bounce > 0);
}

/* everything of estimate_direct_lighting that follows the emissive-geometry sample */
template <class V>
B2_HD Rgb estimate_direct_lighting_from(const Ctx& ctx, Payload& payload, HitInfo& hit, V3 view_direction, Rng& rng, Rgb emissive_contribution)
{
    Rgb envmap_contribution = sample_environment_map<V>(ctx, payload, hit, view_direction, rng);
    return direct_lighting_combine<V>(*ctx.rd, emissive_contribution, envmap_contribution, payload.throughput, payload.bounce, payload.material.emission);
}

/* PathTracing.h:122-174 (denoiser AOVs are written by store_denoiser_aovs below) */
B2_HD void store_denoiser_aovs(const b2pt_render_data& rd, uint32_t pixel_index, V3 shading_normal, Rgb base_color)
{
    const int counter = rd.render_settings.denoiser_AOV_accumulation_counter;
    if (rd.render_settings.sample_number == 0)
        rd.aux_buffers.denoiser_albedo[pixel_index] = b2pt_rgb{ base_color.r, base_color.g, base_color.b };
    else
    {
        Rgb a = (load_rgb(rd.aux_buffers.denoiser_albedo, pixel_index) * (float)counter + base_color) / (counter + 1.0f);
        rd.aux_buffers.denoiser_albedo[pixel_index] = b2pt_rgb{ a.r, a.g, a.b };
    }
    if (rd.render_settings.sample_number == 0)
        rd.aux_buffers.denoiser_normals[pixel_index] = b2pt_float3{ shading_normal.x, shading_normal.y, shading_normal.z };
    else
    {
        V3 acc = (load_v3(rd.aux_buffers.denoiser_normals, pixel_index) * (float)counter + shading_normal) / (counter + 1.0f);
        float len = length(acc);
        if (!b2_is_zero(len))
        {
            acc = acc / len;
            rd.aux_buffers.denoiser_normals[pixel_index] = b2pt_float3{ acc.x, acc.y, acc.z };
        }
    }
}

template <class V>
B2_HD Rgb miss_gather_envmap(const b2pt_render_data& rd, Rgb ray_throughput, V3 ray_direction, int bounce, uint32_t pixel_index)
{
    const b2pt_world_settings& ws = rd.world_settings;
    Rgb skysphere_color = rgb(0.0f);
    if (ws.ambient_light_type == 1 || rd.bsdfs_data.white_furnace_mode)
        skysphere_color = rgb(ws.uniform_light_color.r, ws.uniform_light_color.g, ws.uniform_light_color.b);
    else if (ws.ambient_light_type == 2 && ws.envmap_intensity == 0.0f)
        return rgb(0.0f);
    else if (ws.ambient_light_type == 2)
    {
        if (V::ENVMAP == B2_ESS_NONE || bounce == 0)
        {
            skysphere_color = eval_envmap_no_pdf(ws, ray_direction);
            bool unscale = V::ENVMAP == B2_ESS_NONE ? (!ws.envmap_scale_background_intensity && bounce == 0) : !ws.envmap_scale_background_intensity;
            if (unscale)
                skysphere_color /= ws.envmap_intensity;
        }
    }
    skysphere_color = clamp_light_contribution(skysphere_color, rd.render_settings.envmap_contribution_clamp, true);
    Rgb contribution = skysphere_color * ray_throughput;
    Rgb clamped = clamp_light_contribution(contribution, rd.render_settings.indirect_contribution_clamp, bounce > 0);
    if (bounce == 0)
        store_denoiser_aovs(rd, pixel_index, v3(0.0f, 0.0f, 0.0f), skysphere_color);
    return clamped;
}

/* SanityCheck.h:78-105: returns false for NaN / Inf / negative samples. The reference's host build also runs this
 * check after every direct-lighting estimate (sanity_check<true>), which is what the parity target does. */
B2_HD bool sanity_check(const b2pt_render_data& rd, Rgb& color, int x, int y)
{
    bool invalid = (color.r < 0 || color.g < 0 || color.b < 0) || isnan(color.r) || isnan(color.g) || isnan(color.b) || isinf(color.r) || isinf(color.g) || isinf(color.b);
    if (invalid)
    {
        if (rd.render_settings.display_NaNs)
        {
            Rgb pink = rgb(1.0e30f, 0.0f, 1.0e30f);
            if (rd.render_settings.sample_number != 0)
                pink = pink * (float)rd.render_settings.sample_number;
            rd.buffers.accumulated_ray_colors[y * rd.render_settings.render_resolution.x + x] = b2pt_rgb{ pink.r, pink.g, pink.b };
        }
        else
            color = rgb(0.0f);
    }
    return !invalid;
}

/* PathTracing.h:176-217 (GMoN sets out of scope) */
B2_HD void accumulate_color(const b2pt_render_data& rd, Rgb ray_color, uint32_t pixel_index)
{
    if (has_access_to_adaptive_sampling_buffers(rd.render_settings))
    {
        float squared_luminance_of_samples = luminance(ray_color) * luminance(ray_color);
        rd.aux_buffers.pixel_squared_luminance[pixel_index] += squared_luminance_of_samples;
    }
    if (rd.render_settings.sample_number == 0)
        rd.buffers.accumulated_ray_colors[pixel_index] = b2pt_rgb{ ray_color.r, ray_color.g, ray_color.b };
    else
    {
        Rgb c = load_rgb(rd.buffers.accumulated_ray_colors, pixel_index) + ray_color;
        rd.buffers.accumulated_ray_colors[pixel_index] = b2pt_rgb{ c.r, c.g, c.b };
    }
}

/* RayVolumeState::reconstruct_first_hit, RayVolumeState.h:29-58 */
template <class V>
B2_HD void reconstruct_first_hit(VolumeState& vs, const Mat& material, const int* material_indices, int primitive_index, Rng& rng)
{
    if (primitive_index == -1)
        return;
    int mat_index = material_indices[primitive_index];
    vs.interior_stack.push(vs.incident_mat_index, vs.outgoing_mat_index, vs.inside_material, mat_index, effective_dielectric_priority<V>(material));
    if (material.dispersion_scale > 0.0f && material.specular_transmission > 0.0f && vs.sampled_wavelength == 0.0f)
        vs.sampled_wavelength = -sample_wavelength_uniformly(rng);
}

/* The MegaKernel pass for pixel (x, y): Megakernel.h:22-144, cut into begin / one bounce / end so that the kernel can
 * decide how threads and paths meet: the plain per-pixel loop (megakernel_pixel below: host port, tests) or the
 * persistent warps of b2_kernels.cu that hand a finished lane the next pixel ("path regeneration") instead of idling
 * it until the longest path of the warp ends. Per-pixel arithmetic and RNG order are identical either way. */
struct PathState
{
    Payload payload;
    HitInfo hit;
    V3 ray_origin, ray_direction;
    Rng rng;
    uint32_t pixel_index;
    int x, y;
    int nb_bounces;
    bool intersection_found;
};

/* Megakernel.h:33-79. Returns false when the pixel takes no sample (outside the frame / inactive). */
template <class V>
B2_HD bool megakernel_begin(const Ctx& ctx, uint32_t x, uint32_t y, PathState& s)
{
    const b2pt_render_data& rd = *ctx.rd;
    const b2pt_int2 res = rd.render_settings.render_resolution;
    if (x >= (uint32_t)res.x || y >= (uint32_t)res.y)
        return false;
    uint32_t pixel_index = x + y * res.x;
    if (!rd.aux_buffers.pixel_active[pixel_index])
        return false;

    s.pixel_index = pixel_index;
    s.x = (int)x;
    s.y = (int)y;
    s.nb_bounces = rd.render_settings.nb_bounces;
    if (do_render_low_resolution(rd.render_settings))
        s.nb_bounces = b2_min(3, s.nb_bounces);

    s.rng = make_rng(pixel_seed(rd, pixel_index));

    s.hit = hit_info_defaults();
    s.hit.inter_point = load_v3(rd.g_buffer.primary_hit_position, pixel_index);
    s.hit.geometric_normal = normalize(unpack_normal_oct24(rd.g_buffer.geometric_normals[pixel_index]));
    s.hit.shading_normal = normalize(unpack_normal_oct24(rd.g_buffer.shading_normals[pixel_index]));
    s.hit.primitive_index = rd.g_buffer.first_hit_prim_index[pixel_index];

    V3 cam_pos = v3(rd.current_camera.position.x, rd.current_camera.position.y, rd.current_camera.position.z);
    s.ray_origin = v3(0.0f);
    s.ray_direction = normalize(-normalize(cam_pos - s.hit.inter_point));

    s.payload.init();
    s.payload.material = unpack_material(load_packed_material(rd.g_buffer.materials, pixel_index));
    reconstruct_first_hit<V>(s.payload.volume_state, s.payload.material, rd.buffers.material_indices, s.hit.primitive_index, s.rng);

    s.intersection_found = s.hit.primitive_index != -1;
    return true;
}

/* One iteration of the bounce loop (Megakernel.h:81-135). Returns true when the loop goes on to the next bounce. */
template <class V>
B2_HD bool megakernel_bounce(const Ctx& ctx, PathState& s)
{
    const b2pt_render_data& rd = *ctx.rd;
    Payload& payload = s.payload;
    HitInfo& hit = s.hit;
    const int bounce = payload.bounce;
    if (!(bounce < s.nb_bounces + 1))
        return false;
    if (payload.next_ray_state == B2_RAY_MISSED)
        return false;
    B2_SET_RAY_BOUNCE(ctx, bounce); /* every ray of this iteration is traced with FilterFunctionPayload::bounce = bounce (Megakernel.h:81-131) */

    if (bounce > 0)
        s.intersection_found = trace_main_path_ray<V>(ctx, s.ray_origin, s.ray_direction, payload, hit, hit.primitive_index, s.rng);

    if (s.intersection_found)
    {
        if (bounce == 0)
            store_denoiser_aovs(rd, s.pixel_index, hit.shading_normal, payload.material.base_color);

        if (bounce > 0 || rd.render_settings.enable_direct)
        {
            payload.ray_color += estimate_direct_lighting<V>(ctx, payload, hit, -s.ray_direction, s.x, s.y, s.rng);
            sanity_check(rd, payload.ray_color, s.x, s.y);
        }

        bool valid_indirect_bounce = compute_next_indirect_bounce<V>(ctx, payload, hit, -s.ray_direction, s.ray_origin, s.ray_direction, s.rng);
        if (!valid_indirect_bounce)
            return false;
    }
    else
    {
        payload.ray_color += miss_gather_envmap<V>(rd, payload.throughput, s.ray_direction, payload.bounce, s.pixel_index);
        payload.next_ray_state = B2_RAY_MISSED;
        sanity_check(rd, payload.ray_color, s.x, s.y);
    }
    payload.bounce++;
    return true;
}

/* Megakernel.h:137-144 */
B2_HD void megakernel_end(const Ctx& ctx, PathState& s)
{
    const b2pt_render_data& rd = *ctx.rd;
    if (!sanity_check(rd, s.payload.ray_color, s.x, s.y))
        return;
    rd.aux_buffers.still_one_ray_active[0] = 1;
    accumulate_color(rd, s.payload.ray_color, s.pixel_index);
}

/* What begin + bounce + end compute for an active pixel whose primary ray missed (first_hit_prim_index == -1): no RNG
 * draw happens on that path (reconstruct_first_hit returns at once), the first loop iteration gathers the background
 * with throughput 1 and the path ends. The persistent kernel runs this for the miss pixels of a freshly fetched tile
 * with all of their lanes together instead of giving them a slot in the bounce loop. */
template <class V>
B2_HD void megakernel_primary_miss(const Ctx& ctx, uint32_t x, uint32_t y, uint32_t pixel_index)
{
    const b2pt_render_data& rd = *ctx.rd;
    V3 inter_point = load_v3(rd.g_buffer.primary_hit_position, pixel_index);
    V3 cam_pos = v3(rd.current_camera.position.x, rd.current_camera.position.y, rd.current_camera.position.z);
    V3 ray_direction = normalize(-normalize(cam_pos - inter_point));
    Rgb ray_color = rgb(0.0f);
    ray_color += miss_gather_envmap<V>(rd, rgb(1.0f), ray_direction, 0, pixel_index);
    sanity_check(rd, ray_color, (int)x, (int)y);
    if (!sanity_check(rd, ray_color, (int)x, (int)y))
        return;
    rd.aux_buffers.still_one_ray_active[0] = 1;
    accumulate_color(rd, ray_color, pixel_index);
}

/* ---- the bounce cut at its rays (b2_nee.h) for the phased MegaKernel -----------------------------------------------------
 * One iteration of a path:
 *     trace A: [main ray, deferred X, deferred Y, envmap light ray, envmap BSDF ray of the previous vertex]
 *     stage 1: resolve the previous vertex's NEE; end the path if it has no main ray; take the main hit; NEE begin
 *     trace B: [sequential ray]                       (only variants with RIS)
 *     stage 2: rest of the NEE sampling -> deferred X / Y; envmap NEE; next direction -> main ray
 * Same operations and RNG draws, in the same order, as megakernel_bounce; what moves is WHEN the NEE radiance of a vertex
 * is added to the pixel (one trace later), which no RNG draw and no decision of the path depends on. */
template <class V>
struct BouncePhase
{
    Nee<V> nee;
    bool nee_on;
    /* a vertex whose NEE waits for its deferred rays */
    bool pending;
    Rgb pending_throughput, pending_material_emission, pending_self_emission;
    EnvmapNee envmap; /* the vertex's envmap NEE, waiting for its two deferred shadow rays like `nee` for its own */
    int pending_bounce;
    /* the path has no next main ray: it ends as soon as its pending NEE is resolved */
    bool ending;

    B2_HD void reset()
    {
        nee_on = false;
        pending = false;
        ending = false;
    }
};

/* Returns false when the path ended here. `main_hit`: closest hit of the main ray when bounce > 0 (null = trace here). */
template <class V>
B2_HD bool megakernel_stage1(const Ctx& ctx, PathState& s, BouncePhase<V>& ph, const B2Hit* main_hit, const B2Hit& result_x, const B2Hit& result_y, const B2Hit& result_env_light,
                             const B2Hit& result_env_bsdf, RayRequest& sequential_ray)
{
    const b2pt_render_data& rd = *ctx.rd;
    Payload& payload = s.payload;
    HitInfo& hit = s.hit;
    request_none(sequential_ray);

    if (ph.pending)
    {
        /* sample_emissive_geometry's value for the previous vertex (s.hit still is that vertex), then the tail of
         * estimate_direct_lighting and the sanity check that follows it in the bounce loop */
        Rgb emissive_contribution = rgb(0.0f);
        if (ph.nee.active)
            emissive_contribution = ph.nee.resolve(rd, hit.shading_normal, result_x, result_y) / 1.0f + ph.pending_self_emission;
        payload.ray_color += direct_lighting_combine<V>(rd, emissive_contribution, ph.envmap.resolve(result_env_light, result_env_bsdf), ph.pending_throughput, ph.pending_bounce,
                                                        ph.pending_material_emission);
        sanity_check(rd, payload.ray_color, s.x, s.y);
        ph.pending = false;
    }
    if (ph.ending)
        return false;

    const int bounce = payload.bounce;
    ph.nee_on = false;
    B2_SET_RAY_BOUNCE(ctx, bounce);
    if (bounce > 0)
        s.intersection_found = trace_main_path_ray<V>(ctx, s.ray_origin, s.ray_direction, payload, hit, hit.primitive_index, s.rng, main_hit);

    if (!s.intersection_found)
    {
        payload.ray_color += miss_gather_envmap<V>(rd, payload.throughput, s.ray_direction, payload.bounce, s.pixel_index);
        payload.next_ray_state = B2_RAY_MISSED;
        sanity_check(rd, payload.ray_color, s.x, s.y);
        return false;
    }
    if (bounce == 0)
        store_denoiser_aovs(rd, s.pixel_index, hit.shading_normal, payload.material.base_color);
    if (bounce > 0 || rd.render_settings.enable_direct)
    {
        ph.nee_on = true;
        ph.nee.begin(ctx, payload, hit, -s.ray_direction, s.rng, sequential_ray);
    }
    return true;
}

/* Returns false when the path ended here (no next main ray and nothing pending). */
template <class V>
B2_HD bool megakernel_stage2(const Ctx& ctx, PathState& s, BouncePhase<V>& ph, const B2Hit& sequential_result, RayRequest& main_ray, RayRequest& deferred_x, RayRequest& deferred_y,
                             RayRequest& env_light_ray, RayRequest& env_bsdf_ray)
{
    Payload& payload = s.payload;
    HitInfo& hit = s.hit;
    request_none(main_ray);
    request_none(deferred_x);
    request_none(deferred_y);
    request_none(env_light_ray);
    request_none(env_bsdf_ray);
    B2_SET_RAY_BOUNCE(ctx, payload.bounce);
    if (ph.nee_on)
    {
        ph.nee.finish_sampling(ctx, payload, hit, -s.ray_direction, s.rng, s.x, s.y, sequential_result, deferred_x, deferred_y);
        ph.envmap.template sample<V>(ctx, payload, hit, -s.ray_direction, s.rng, env_light_ray, env_bsdf_ray);
        ph.pending_throughput = payload.throughput;
        ph.pending_bounce = payload.bounce;
        ph.pending_material_emission = payload.material.emission;
        ph.pending_self_emission = payload.material.emissive_texture_used ? payload.material.emission : rgb(0.0f);
        ph.pending = true;
    }
    bool valid_indirect_bounce = compute_next_indirect_bounce<V>(ctx, payload, hit, -s.ray_direction, s.ray_origin, s.ray_direction, s.rng);
    if (valid_indirect_bounce)
    {
        payload.bounce++;
        valid_indirect_bounce = payload.bounce < s.nb_bounces + 1;
    }
    if (valid_indirect_bounce)
        request_closest(main_ray, s.ray_origin, s.ray_direction, hit.primitive_index, /* main_path_ray = */ true);
    else
        ph.ending = true;
    return valid_indirect_bounce || ph.pending;
}

template <class V>
B2_HD void megakernel_pixel(const Ctx& ctx, uint32_t x, uint32_t y)
{
    PathState s;
    if (!megakernel_begin<V>(ctx, x, y, s))
        return;
    while (megakernel_bounce<V>(ctx, s))
    {
    }
    megakernel_end(ctx, s);
}

#endif
