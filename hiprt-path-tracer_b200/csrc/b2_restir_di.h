/*
 * b2_restir_di.h -- the ReSTIR DI passes as per-pixel functions: initial candidates, temporal reuse, spatial reuse
 * and the fused spatiotemporal reuse, with pairwise-MIS bias correction (confidence weights), visibility in the
 * bias-correction target functions and visibility reuse -- the reference's default macro set.
 *
 * Restates /root/reference/src/Device/kernels/ReSTIR/DI/InitialCandidates.h:92-538, TemporalReuse.h:49-296,
 * SpatialReuse.h:42-322, FusedSpatiotemporalReuse.h:43-564 and their helpers:
 *   includes/ReSTIR/Utils.h:25-65                (visibility reuse: kill / mark-unoccluded)
 *   includes/ReSTIR/UtilsTemporal.h:11-96        (reprojection, permutation sampling, neighbour search)
 *   includes/ReSTIR/UtilsSpatial.h:17-339        (disk neighbours, directional masks, RNG replay, valid-neighbour count)
 *   includes/ReSTIR/NeighborSimilarity.h:16-107  (plane distance / normal / roughness heuristics)
 *   includes/ReSTIR/TemporalMISWeight.h:163-254, SpatialMISWeight.h:119-235, SpatiotemporalMISWeight.h:140-225
 * The reference kernels take HIPRTRenderData BY VALUE and a few of them scribble per-pixel values into their private
 * copy (spatial_neighbors_rng_seed, reuse_radius, reuse_neighbor_count, the directions mask); here the render data is a
 * shared __grid_constant__, so those four live in `SpatialLocal`.
 * Quirks kept on purpose (the oracle has them): the DI spatial kernel keys its seed on the ReSTIR *GI* setting
 * `coalesced_spatial_reuse` (SpatialReuse.h:66); the fused kernel XORs instead of multiplying its seed (:173) and draws
 * the neighbour seed twice (:177 then UtilsSpatial.h:22); similarity heuristics read the neighbour's normal from the
 * CURRENT G-buffer even when comparing against the previous frame (NeighborSimilarity.h:100).
 */
#ifndef B2_RESTIR_DI_H
#define B2_RESTIR_DI_H

#include "b2_pathtrace.h"

#define B2_LIGHT_DOESNT_CONTRIBUTE_ENOUGH -42.0f
#define B2_DIRECTIONAL_REUSE_BIT_COUNT 64

/* ReSTIR GI's common_spatial_pass.coalesced_spatial_reuse inside the opaque GI settings block: ReSTIRGISettings starts
 * with the same ReSTIRCommonSettings prefix as ReSTIRDISettings (oracle/layout_check.cpp asserts the offset) */
B2_HD bool restir_gi_coalesced_spatial_reuse(const b2pt_render_data& rd)
{
    const uint8_t* gi = reinterpret_cast<const uint8_t*>(rd.render_settings.restir_gi_settings_opaque);
    const size_t offset = (size_t)(reinterpret_cast<const uint8_t*>(&rd.render_settings.restir_di_settings.common_spatial_pass.coalesced_spatial_reuse) -
                                   reinterpret_cast<const uint8_t*>(&rd.render_settings.restir_di_settings));
    return gi[offset] != 0;
}

/* ReSTIR_DI_BiasCorrectionWeights (HostDeviceCommon/KernelOptions/ReSTIRDIOptions.h:9-16) */
enum { B2_BIAS_1_OVER_M = 0, B2_BIAS_1_OVER_Z = 1, B2_BIAS_MIS_LIKE = 2, B2_BIAS_MIS_GBH = 3, B2_BIAS_PAIRWISE_MIS = 4, B2_BIAS_PAIRWISE_MIS_DEFENSIVE = 5,
       B2_BIAS_SYMMETRIC_RATIO = 6, B2_BIAS_ASYMMETRIC_RATIO = 7 };

/* The four visibility macros of ReSTIRDIOptions.h:50-84 as launch-time switches (Ctx::restir_options). The reference's defaults:
 * visibility reuse and visibility in the bias-correction target functions on, the other two off. The phase-cut (wavefront) forms of
 * InitialCandidates and of the fused pass are built for the defaults; any other setting runs the passes in one piece. */
enum { B2_DI_OPT_INITIAL_VISIBILITY = 1, B2_DI_OPT_SPATIAL_VISIBILITY = 2, B2_DI_OPT_VISIBILITY_REUSE = 4, B2_DI_OPT_BIAS_VISIBILITY = 8,
       B2_DI_OPT_DEFAULTS = B2_DI_OPT_VISIBILITY_REUSE | B2_DI_OPT_BIAS_VISIBILITY,
       /* ReSTIR_DI_DoOptimalVisibilitySampling (ReSTIRDIOptions.h:176): read by the spatial pass only, so it does not take the other passes
        * off their phase-cut forms (B2_DI_OPT_PHASE_CUT_MASK = the bits those forms depend on) */
       B2_DI_OPT_OPTIMAL_VISIBILITY_SAMPLING = 16, B2_DI_OPT_PHASE_CUT_MASK = 15 };

/* ReSTIR_DI_evaluate_target_function<ReSTIR_DI_BiasCorrectionUseVisibility> */
template <class V>
B2_HD float di_target_function_bias(const Ctx& ctx, const DiSample& sample, DiSurface& surface)
{
    if (ctx.restir_options & B2_DI_OPT_BIAS_VISIBILITY)
        return di_target_function<V, true>(ctx, sample, surface);
    return di_target_function<V, false>(ctx, sample, surface);
}

/* the visibility-reuse pass at the end of the spatial / fused pass exists when the bias-correction target functions see visibility and
 * the reservoirs that come in do too (SpatialReuse.h:264-270, FusedSpatiotemporalReuse.h:507-513) */
B2_HD bool di_end_of_pass_visibility_reuse(const Ctx& ctx)
{
    const int o = ctx.restir_options;
    return (o & B2_DI_OPT_BIAS_VISIBILITY) && ((o & B2_DI_OPT_VISIBILITY_REUSE) || ((o & B2_DI_OPT_INITIAL_VISIBILITY) && (o & B2_DI_OPT_SPATIAL_VISIBILITY)));
}

/* MISWeightsCommon.h:49-64: min(a / b, b / a) ^ beta, 0 when either target function is */
B2_HD float ratio_mis_difference_function(float target_function_at_center, float target_function_from_i, float exponent)
{
    if (target_function_at_center == 0.0f || target_function_from_i == 0.0f)
        return 0.0f;
    float ratio = b2_min(target_function_at_center / target_function_from_i, target_function_from_i / target_function_at_center);
    if (exponent == 2.0f)
        return ratio * ratio;
    if (exponent == 3.0f)
        return ratio * ratio * ratio;
    if (exponent == 4.0f)
    {
        float ratio2 = ratio * ratio;
        return ratio2 * ratio2;
    }
    return b2f_pow(ratio, exponent);
}

struct SpatialLocal
{
    unsigned int rng_seed;
    int reuse_radius;
    int reuse_neighbor_count;
    unsigned long long directions_mask;
};

B2_HD SpatialLocal spatial_local_from(const b2pt_render_data& rd)
{
    const b2pt_restir_spatial_common& s = rd.render_settings.restir_di_settings.common_spatial_pass;
    SpatialLocal l;
    l.rng_seed = s.spatial_neighbors_rng_seed;
    l.reuse_radius = s.reuse_radius;
    l.reuse_neighbor_count = s.reuse_neighbor_count;
    l.directions_mask = s.current_pixel_directions_reuse_mask;
    return l;
}

B2_HD bool do_adaptive_directional_spatial_reuse(const b2pt_render_data& rd)
{
    return rd.render_settings.restir_di_settings.common_spatial_pass.use_adaptive_directional_spatial_reuse && rd.render_settings.accumulate;
}

/* ---------------------------------------------------------------------------------------------------------- */
/* visibility reuse (ReSTIR/Utils.h:25-65) */
B2_HD bool di_visibility_test_kill_reservoir(const Ctx& ctx, DiReservoir& reservoir, V3 shading_point, int last_primitive_hit_index)
{
    if (reservoir.UCW <= 0.0f && reservoir.weight_sum <= 0.0f)
        return false;
    else if (reservoir.sample.flags & B2_DI_FLAGS_UNOCCLUDED)
        return false;

    float distance_to_light;
    V3 sample_direction = di_light_sample_direction(*ctx.rd, reservoir.sample, shading_point, distance_to_light);
    bool visible = !shadow_ray_occluded(ctx, shading_point, sample_direction, distance_to_light, last_primitive_hit_index);
    if (!visible)
    {
        reservoir.UCW = -1.0f;
        return true;
    }
    reservoir.sample.or_flags(B2_DI_FLAGS_UNOCCLUDED);
    return false;
}

/* ---------------------------------------------------------------------------------------------------------- */
/* neighbour similarity (ReSTIR/NeighborSimilarity.h) */
B2_HD bool check_neighbor_similarity_heuristics(const b2pt_render_data& rd, int neighbor_pixel_index, int center_pixel_index, V3 current_shading_point, V3 current_normal,
                                                bool previous_frame)
{
    if (neighbor_pixel_index == center_pixel_index)
        return true;
    const b2pt_gbuffer& g = previous_frame ? rd.g_buffer_prev_frame : rd.g_buffer;
    if (g.first_hit_prim_index[neighbor_pixel_index] == -1)
        return false;

    const b2pt_restir_neighbor_similarity& s = rd.render_settings.restir_di_settings.neighbor_similarity_settings;
    V3 neighbor_world_space_point = v3(0.0f);
    float neighbor_roughness = 0.0f;
    float current_material_roughness = 0.0f;
    if (previous_frame)
    {
        if (s.use_plane_distance_heuristic)
            neighbor_world_space_point = load_v3(g.primary_hit_position, neighbor_pixel_index);
        if (s.use_roughness_similarity_heuristic)
            neighbor_roughness = unpack_hi8(reinterpret_cast<const uint32_t*>(g.materials)[(size_t)neighbor_pixel_index * PM_WORDS + PM_BASE_COLOR_ROUGHNESS]);
    }
    else
    {
        neighbor_world_space_point = load_v3(g.primary_hit_position, neighbor_pixel_index);
        neighbor_roughness = unpack_hi8(reinterpret_cast<const uint32_t*>(g.materials)[(size_t)neighbor_pixel_index * PM_WORDS + PM_BASE_COLOR_ROUGHNESS]);
    }
    if (s.use_roughness_similarity_heuristic)
        current_material_roughness = unpack_hi8(reinterpret_cast<const uint32_t*>(rd.g_buffer.materials)[(size_t)center_pixel_index * PM_WORDS + PM_BASE_COLOR_ROUGHNESS]);

    V3 neighbor_normal = s.reject_using_geometric_normals ? unpack_normal_oct24(rd.g_buffer.geometric_normals[neighbor_pixel_index]) : unpack_normal_oct24(rd.g_buffer.shading_normals[neighbor_pixel_index]);

    bool plane_distance_passed = true;
    if (s.use_plane_distance_heuristic)
    {
        V3 direction_between_points = neighbor_world_space_point - current_shading_point;
        float distance_to_plane = fabsf(dot(direction_between_points, current_normal));
        plane_distance_passed = distance_to_plane < s.plane_distance_threshold;
    }
    bool normal_similarity_passed = true;
    if (s.use_normal_similarity_heuristic)
        normal_similarity_passed = dot(current_normal, neighbor_normal) > s.normal_similarity_angle_precomp;
    bool roughness_similarity_passed = true;
    if (s.use_roughness_similarity_heuristic)
        roughness_similarity_passed = fabsf(neighbor_roughness - current_material_roughness) < s.roughness_similarity_threshold;
    return plane_distance_passed && normal_similarity_passed && roughness_similarity_passed;
}

B2_HD V3 normal_for_rejection_heuristic(const b2pt_render_data& rd, const DiSurface& surface)
{
    return rd.render_settings.restir_di_settings.neighbor_similarity_settings.reject_using_geometric_normals ? surface.geometric_normal : surface.shading_normal;
}

/* ---------------------------------------------------------------------------------------------------------- */
/* temporal neighbour (ReSTIR/UtilsTemporal.h) */
struct Int3 { int x, y, z; };

template <class V>
B2_HD Int3 find_temporal_neighbor_index(const b2pt_render_data& rd, V3 current_shading_point, V3 current_normal, int center_pixel_index, Rng& rng)
{
    const b2pt_int2 resolution = rd.render_settings.render_resolution;
    if (rd.render_settings.accumulate)
        return Int3{ center_pixel_index, center_pixel_index % resolution.x, center_pixel_index / resolution.x };

    const b2pt_restir_temporal_common& t = rd.render_settings.restir_di_settings.common_temporal_pass;
    V3 prev_xyz = mat4_mul_point(rd.prev_camera.view_projection, current_shading_point);
    V2 prev_ss = v2(prev_xyz.x, prev_xyz.y);
    prev_ss = prev_ss + v2(1.0f, 1.0f);
    prev_ss = prev_ss * v2(0.5f, 0.5f);
    V2 prev_pixel_float = v2(prev_ss.x * resolution.x, prev_ss.y * resolution.y);
    prev_pixel_float = prev_pixel_float - v2(0.5f, 0.5f);

    int temporal_neighbor_index = -1;
    for (int i = 0; i < t.max_neighbor_search_count + 1; i++)
    {
        V2 offset = v2(0.0f, 0.0f);
        if (i > 0)
        {
            /* make_float2(rng() - 0.5f, rng() - 0.5f): right-to-left argument evaluation under g++, see get_spatial_neighbor_pixel_index */
            float oy = rng.next_float() - 0.5f;
            float ox = rng.next_float() - 0.5f;
            offset = v2(ox, oy) * (float)t.neighbor_search_radius;
        }
        int px = (int)roundf(prev_pixel_float.x + offset.x), py = (int)roundf(prev_pixel_float.y + offset.y);
        if (t.use_permutation_sampling && i == 0)
        {
            int random_bits = t.permutation_sampling_random_bits;
            int offx = random_bits & 3, offy = (random_bits >> 2) & 3;
            px += offx;
            py += offy;
            px ^= 3;
            py ^= 3;
            px -= offx;
            py -= offy;
        }
        if (px < 0 || px >= resolution.x || py < 0 || py >= resolution.y)
            continue;
        temporal_neighbor_index = px + py * resolution.x;
        if (check_neighbor_similarity_heuristics(rd, temporal_neighbor_index, center_pixel_index, current_shading_point, current_normal, use_prev_frame_g_buffer<V>(rd.render_settings)))
            break;
        temporal_neighbor_index = -1;
    }
    return Int3{ temporal_neighbor_index, (int)roundf(prev_pixel_float.x), (int)roundf(prev_pixel_float.y) };
}

/* ---------------------------------------------------------------------------------------------------------- */
/* spatial neighbours (ReSTIR/UtilsSpatial.h) */
B2_HD int b2_popcll(unsigned long long v)
{
#if defined(__CUDA_ARCH__)
    return __popcll(v);
#else
    return __builtin_popcountll(v);
#endif
}

B2_HD void setup_adaptive_directional_spatial_reuse(const b2pt_render_data& rd, SpatialLocal& local, unsigned int center_pixel_index, Rng& rng)
{
    const b2pt_restir_spatial_common& s = rd.render_settings.restir_di_settings.common_spatial_pass;
    local.rng_seed = rng.next_u32();
    if (do_adaptive_directional_spatial_reuse(rd))
    {
        local.reuse_radius = s.per_pixel_spatial_reuse_radius[center_pixel_index];
        local.directions_mask = s.per_pixel_spatial_reuse_directions_mask_ull[center_pixel_index];
        if (local.reuse_radius == 0)
            local.reuse_neighbor_count = 0;
    }
}

B2_HD V2 sample_spatial_neighbor_from_allowed_directions(const SpatialLocal& local, Rng& rng)
{
    unsigned long long directions_mask = local.directions_mask;
    int number_of_allowed_sectors = b2_popcll(directions_mask);
    unsigned char random_sector_index = (unsigned char)rng.next_index(number_of_allowed_sectors);
    unsigned char count_left_to_go = (unsigned char)(random_sector_index + 1);
    int sector_index = 0;
    unsigned char bit_count_so_far = 0;
    if (b2_popcll(directions_mask) == B2_DIRECTIONAL_REUSE_BIT_COUNT)
        sector_index = random_sector_index;
    else
    {
        while (count_left_to_go > 0)
        {
            unsigned char mask_length = count_left_to_go;
            unsigned long long mask = ((1ull << mask_length) - 1ull) << bit_count_so_far;
            int count_mask = b2_popcll(directions_mask & mask);
            count_left_to_go = (unsigned char)(count_left_to_go - count_mask);
            bit_count_so_far = (unsigned char)(bit_count_so_far + mask_length);
        }
        sector_index = --bit_count_so_far;
    }
    float theta_start = sector_index / (float)B2_DIRECTIONAL_REUSE_BIT_COUNT;
    float random_theta = theta_start + rng.next_float() * (1.0f / (float)B2_DIRECTIONAL_REUSE_BIT_COUNT);
    float second = rng.next_float();
    return v2(random_theta, second);
}

B2_HD int get_spatial_neighbor_pixel_index(const b2pt_render_data& rd, const SpatialLocal& local, int neighbor_index, int cx, int cy, Rng& rng)
{
    const b2pt_restir_spatial_common& s = rd.render_settings.restir_di_settings.common_spatial_pass;
    const b2pt_int2 res = rd.render_settings.render_resolution;
    if (neighbor_index == local.reuse_neighbor_count)
        return cx + cy * res.x;

    V2 uv;
    if (do_adaptive_directional_spatial_reuse(rd))
        uv = sample_spatial_neighbor_from_allowed_directions(local, rng);
    else
    {
        /* `make_float2(rng(), rng())` in the reference: the order of the two draws is unspecified in C++ and g++ (the
         * compiler of the parity oracle) evaluates call arguments right to left, so .y gets the first draw */
        float u1 = rng.next_float();
        float u0 = rng.next_float();
        uv = v2(u0, u1);
    }
    V2 neighbor_offset_in_disk = sample_in_disk_uv((float)local.reuse_radius, uv);
    int ox = (int)roundf(neighbor_offset_in_disk.x), oy = (int)roundf(neighbor_offset_in_disk.y);

    int nx, ny;
    if (s.debug_neighbor_location)
    {
        int dx, dy;
        if (s.debug_neighbor_location_direction == 0) { dx = local.reuse_radius; dy = 0; }
        else if (s.debug_neighbor_location_direction == 1) { dx = 0; dy = local.reuse_radius; }
        else { dx = local.reuse_radius; dy = local.reuse_radius; }
        nx = cx + dx;
        ny = cy + dy;
    }
    else
    {
        nx = cx + ox;
        ny = cy + oy;
    }
    if (nx < 0 || nx >= res.x || ny < 0 || ny >= res.y)
        return -1;

    int neighbor_pixel_index = nx + ny * res.x;
    if (rd.render_settings.enable_adaptive_sampling && rd.render_settings.sample_number >= (unsigned int)rd.render_settings.adaptive_sampling_min_samples)
    {
        if (s.allow_converged_neighbors_reuse)
        {
            Rng rng_converged = make_rng(rd.random_number);
            if (rng_converged.next_float() > s.converged_neighbor_reuse_probability)
                if (rd.aux_buffers.pixel_converged_sample_count[neighbor_pixel_index] != -1)
                    return -1;
        }
        else if (rd.aux_buffers.pixel_converged_sample_count[neighbor_pixel_index] != -1)
            return -1;
    }
    return neighbor_pixel_index;
}

B2_HD void spatial_neighbor_advance_rng(const b2pt_render_data& rd, Rng& rng)
{
    rng.next_u32();
    rng.next_u32();
    if (do_adaptive_directional_spatial_reuse(rd))
        rng.next_u32();
}

/* count_valid_spatial_neighbors / count_valid_spatiotemporal_neighbors: replay the neighbour RNG, apply the
 * heuristics, sum the neighbours' M and remember who passed in a bit cache */
B2_HD void count_valid_neighbors(const b2pt_render_data& rd, const SpatialLocal& local, const DiSurface& center_pixel_surface, int center_pixel_index, int around_x, int around_y,
                                 bool previous_frame, const void* input_reservoirs, int& out_valid_neighbor_count, int& out_valid_neighbor_M_sum, int& out_neighbor_heuristics_cache)
{
    out_valid_neighbor_count = 0;
    Rng spatial_neighbors_rng = make_rng(local.rng_seed);
    for (int neighbor_index = 0; neighbor_index < local.reuse_neighbor_count; neighbor_index++)
    {
        int neighbor_pixel_index = get_spatial_neighbor_pixel_index(rd, local, neighbor_index, around_x, around_y, spatial_neighbors_rng);
        if (neighbor_pixel_index == -1)
            continue;
        if (!check_neighbor_similarity_heuristics(rd, neighbor_pixel_index, center_pixel_index, center_pixel_surface.shading_point, normal_for_rejection_heuristic(rd, center_pixel_surface), previous_frame))
            continue;
        out_valid_neighbor_M_sum += di_load_M(rd, input_reservoirs, neighbor_pixel_index);
        out_valid_neighbor_count++;
        out_neighbor_heuristics_cache |= (1 << neighbor_index);
    }
}

B2_HD bool do_include_spatial_visibility_term(const Ctx& ctx, const SpatialLocal& local, int current_neighbor_index)
{
    /* do_include_visibility_term_or_not (UtilsSpatial.h:40-58) = do_include_spatial_visibility_term_or_not (FusedSpatiotemporalReuse.h:44-63);
     * ReSTIR_DI_SpatialTargetFunctionVisibility = FALSE (the reference's default) makes the whole expression false */
    const b2pt_restir_spatial_common& s = ctx.rd->render_settings.restir_di_settings.common_spatial_pass;
    bool is_last_pass = s.spatial_pass_index == s.number_of_passes - 1;
    bool include_target_function_visibility = s.do_visibility_only_last_pass && is_last_pass;
    include_target_function_visibility |= !s.do_visibility_only_last_pass;
    include_target_function_visibility &= current_neighbor_index < s.neighbor_visibility_count;
    include_target_function_visibility &= (ctx.restir_options & B2_DI_OPT_SPATIAL_VISIBILITY) != 0;
    include_target_function_visibility &= current_neighbor_index != local.reuse_neighbor_count;
    return include_target_function_visibility;
}

/* ---------------------------------------------------------------------------------------------------------- */
/* pairwise MIS (confidence weights): shared arithmetic of SpatialMISWeight.h:119-235 and SpatiotemporalMISWeight.h:140-225 */
struct PairwiseMis
{
    float mc;

    /* the weight of resampling one neighbour and the running canonical weight mc, given the target function of the CENTER pixel's
     * canonical sample evaluated at that neighbour's surface (with visibility: ReSTIR_DI_BiasCorrectionUseVisibility); only read
     * when update_mc */
    B2_HD float weight_given(const b2pt_render_data& rd, int reservoir_being_resampled_M, float reservoir_being_resampled_target_function, int center_pixel_reservoir_M,
                             float center_pixel_reservoir_target_function, float target_function_at_center, float target_function_center_sample_at_neighbor,
                             int valid_neighbors_count, int valid_neighbors_M_sum, bool update_mc, bool resampling_canonical, bool guard_zero_confidence, bool defensive = false)
    {
        if (defensive)
        {
            /* RESTIR_DI_BIAS_CORRECTION_PAIRWISE_MIS_DEFENSIVE, SpatialMISWeight.h:236-368: the canonical sample keeps a share of its own */
            const bool confidence = rd.render_settings.restir_di_settings.use_confidence_weights != 0;
            if (resampling_canonical)
            {
                if (mc == 0.0f)
                    return 1.0f;
                if (confidence)
                    return mc + (float)center_pixel_reservoir_M / (float)(center_pixel_reservoir_M + valid_neighbors_M_sum);
                return (1.0f + mc) / ((float)valid_neighbors_count + 1.0f);
            }
            float target_function_at_neighbor = reservoir_being_resampled_target_function;
            float reservoir_resampled_M = confidence ? (float)reservoir_being_resampled_M : 1.0f;
            float center_reservoir_M = confidence ? (float)center_pixel_reservoir_M : 1.0f;
            float neighbors_confidence_sum = confidence ? (float)valid_neighbors_M_sum : 1.0f;
            float valid_neighbor_division_term = confidence ? 1.0f : (float)valid_neighbors_count;
            float nume = target_function_at_neighbor * reservoir_resampled_M;
            float denom = target_function_at_neighbor * neighbors_confidence_sum + target_function_at_center / valid_neighbor_division_term * center_reservoir_M;
            float mi = denom == 0.0f ? 0.0f : (nume / denom);
            if (confidence)
                mi *= neighbors_confidence_sum / (neighbors_confidence_sum + center_reservoir_M);
            if (update_mc)
            {
                float target_function_center_sample_at_center = center_pixel_reservoir_target_function;
                float nume_mc = target_function_center_sample_at_center / valid_neighbor_division_term * center_reservoir_M;
                float denom_mc = target_function_center_sample_at_neighbor * neighbors_confidence_sum + target_function_center_sample_at_center / valid_neighbor_division_term * center_reservoir_M;
                float confidence_multiplier = 1.0f;
                if (confidence)
                    confidence_multiplier = reservoir_resampled_M / (center_reservoir_M + neighbors_confidence_sum);
                if (denom_mc != 0.0f)
                    mc += nume_mc / denom_mc * confidence_multiplier;
            }
            return confidence ? mi : mi / ((float)valid_neighbors_count + 1.0f);
        }
        if (resampling_canonical)
            return mc == 0.0f ? 1.0f : mc;

        float target_function_at_neighbor = reservoir_being_resampled_target_function;
        float target_function_center_sample_at_center = center_pixel_reservoir_target_function;
        bool use_confidence_weights = rd.render_settings.restir_di_settings.use_confidence_weights != 0;
        float reservoir_resampled_M = use_confidence_weights ? (float)reservoir_being_resampled_M : 1.0f;
        float center_reservoir_M = use_confidence_weights ? (float)center_pixel_reservoir_M : 1.0f;
        float neighbors_confidence_sum = use_confidence_weights ? (float)valid_neighbors_M_sum : 1.0f;
        float valid_neighbor_division_term = use_confidence_weights ? 1.0f : (float)valid_neighbors_count;

        float nume = target_function_at_neighbor * reservoir_resampled_M;
        float denom = target_function_at_neighbor * neighbors_confidence_sum + target_function_at_center / valid_neighbor_division_term * center_reservoir_M;
        float mi = denom == 0.0f ? 0.0f : (nume / denom);

        if (update_mc)
        {
            float nume_mc = target_function_center_sample_at_center / valid_neighbor_division_term * center_reservoir_M;
            float denom_mc = target_function_center_sample_at_neighbor * neighbors_confidence_sum + target_function_center_sample_at_center / valid_neighbor_division_term * center_reservoir_M;

            float confidence_weights_multiplier;
            if (use_confidence_weights)
            {
                if (guard_zero_confidence && neighbors_confidence_sum == 0.0f)
                    confidence_weights_multiplier = 0.0f;
                else
                    confidence_weights_multiplier = reservoir_resampled_M / neighbors_confidence_sum;
            }
            else
                confidence_weights_multiplier = 1.0f;

            if (denom_mc != 0.0f)
                mc += nume_mc / denom_mc / valid_neighbor_division_term * confidence_weights_multiplier;
        }
        return mi / valid_neighbor_division_term;
    }

    template <class V>
    B2_HD float weight(const Ctx& ctx, int reservoir_being_resampled_M, float reservoir_being_resampled_target_function, int center_pixel_reservoir_M,
                       float center_pixel_reservoir_target_function, const DiSample& center_pixel_reservoir_sample, float target_function_at_center, int neighbor_pixel_index,
                       bool neighbor_from_previous_frame, int valid_neighbors_count, int valid_neighbors_M_sum, bool update_mc, bool resampling_canonical, bool guard_zero_confidence, Rng& rng,
                       bool defensive = false)
    {
        const b2pt_render_data& rd = *ctx.rd;
        float target_function_center_sample_at_neighbor = 0.0f;
        if (!resampling_canonical && update_mc)
        {
            DiSurface neighbor_pixel_surface = get_pixel_surface<V>(rd, neighbor_pixel_index, neighbor_from_previous_frame, rng);
            target_function_center_sample_at_neighbor = di_target_function_bias<V>(ctx, center_pixel_reservoir_sample, neighbor_pixel_surface);
        }
        return weight_given(rd, reservoir_being_resampled_M, reservoir_being_resampled_target_function, center_pixel_reservoir_M, center_pixel_reservoir_target_function,
                            target_function_at_center, target_function_center_sample_at_neighbor, valid_neighbors_count, valid_neighbors_M_sum, update_mc, resampling_canonical,
                            guard_zero_confidence, defensive);
    }
};

/* ---------------------------------------------------------------------------------------------------------- */
/* ReSTIR_Directional_Reuse_Compute (kernels/ReSTIR/DirectionalReuseCompute.h:30-155, DI instantiation: 64 direction bits):
 * for 32 radii between the minimum and the configured reuse radius, test 64 directions ON the circle of that radius with
 * the neighbour-similarity heuristics (a direction must have passed at the previous radius); keep the radius whose valid arc
 * covers the largest area and the mask of its valid directions. The RNG the reference constructs is never drawn from. */
B2_HD void restir_directional_reuse_compute_pixel(const b2pt_render_data& rd, uint32_t x, uint32_t y, unsigned long long* out_masks_ull, uint8_t* out_radius)
{
    const b2pt_int2 res = rd.render_settings.render_resolution;
    if (x >= (uint32_t)res.x || y >= (uint32_t)res.y)
        return;
    const uint32_t center_pixel_index = x + y * res.x;
    if (!rd.aux_buffers.pixel_active[center_pixel_index])
        return;
    out_masks_ull[center_pixel_index] = 0;
    out_radius[center_pixel_index] = 0;

    const b2pt_restir_spatial_common& sp = rd.render_settings.restir_di_settings.common_spatial_pass;
    V3 center_shading_point = load_v3(rd.g_buffer.primary_hit_position, center_pixel_index);
    V3 center_normal = rd.render_settings.restir_di_settings.neighbor_similarity_settings.reject_using_geometric_normals
                           ? unpack_normal_oct24(rd.g_buffer.geometric_normals[center_pixel_index])
                           : unpack_normal_oct24(rd.g_buffer.shading_normals[center_pixel_index]);
    const int NB_RADIUS = 32, NB_SAMPLES_PER_RADIUS = 64;
    float best_area = 0.0f;
    int best_radius_index = 0;
    unsigned long long best_mask = 0, previous_mask = 0;
    for (int radius_index = 0; radius_index < NB_RADIUS; radius_index++)
    {
        float current_radius = sp.minimum_per_pixel_reuse_radius + (radius_index / (float)NB_RADIUS) * (sp.reuse_radius - sp.minimum_per_pixel_reuse_radius);
        float current_radius_circle_area = B2_PI * current_radius * current_radius;
        float area_at_current_radius = 0.0f;
        unsigned long long mask = 0;
        for (int sample_index = 0; sample_index < NB_SAMPLES_PER_RADIUS; sample_index++)
        {
            if (radius_index > 0 && !(previous_mask & (1ull << sample_index)))
                continue;
            float theta = sample_index / (float)NB_SAMPLES_PER_RADIUS * B2_TWO_PI;
            float x_circle = current_radius * b2f_cos(theta);
            float y_circle = current_radius * b2f_sin(theta);
            int nx = (int)x + (int)roundf(x_circle), ny = (int)y + (int)roundf(y_circle);
            if (nx < 0 || nx >= res.x || ny < 0 || ny >= res.y)
                continue;
            int neighbor_index = nx + ny * res.x;
            if (!check_neighbor_similarity_heuristics(rd, neighbor_index, (int)center_pixel_index, center_shading_point, center_normal, false))
                continue;
            mask |= 1ull << sample_index;
            area_at_current_radius += current_radius_circle_area * (1.0f / NB_SAMPLES_PER_RADIUS);
        }
        if (best_area < area_at_current_radius)
        {
            best_area = area_at_current_radius;
            best_radius_index = radius_index;
            best_mask = mask;
        }
        previous_mask = mask;
    }
    float best_radius = sp.minimum_per_pixel_reuse_radius + (best_radius_index / (float)NB_RADIUS) * (sp.reuse_radius - sp.minimum_per_pixel_reuse_radius);
    if (best_area == 0.0f)
        best_radius = 0.0f;
    out_radius[center_pixel_index] = (uint8_t)best_radius;
    /* = valid_samples_per_radius[best_radius_index]; when no radius has a valid direction that is index 0, an empty mask */
    out_masks_ull[center_pixel_index] = best_mask;
}

/* ---------------------------------------------------------------------------------------------------------- */
/* initial candidates (InitialCandidates.h) */
/* ---- light presampling (kernels/ReSTIR/DI/LightsPresampling.h, includes/ReSTIR/DI/PresampledLight.h) ----------------------------
 * ReSTIRDIPresampledLight is 48 bytes in the reference: tri index, point (envmap samples: direction in envmap space), light
 * normal, radiance, pdf, one flags byte + 3 bytes of padding (written here as one word, flags in the low byte). */
struct B2PresampledLight
{
    int emissive_triangle_index;
    float px, py, pz;
    float nx, ny, nz;
    float r, g, b;
    float pdf;
    uint32_t flags;
};
static_assert(sizeof(B2PresampledLight) == 48, "ReSTIRDIPresampledLight is 48 bytes");

/* LightPresamplingParameters (Device/kernel_parameters/ReSTIR/DI/LightPresamplingParameters.h): first kernel argument of
 * ReSTIR_DI_LightsPresampling */
struct B2LightPresamplingParameters
{
    int number_of_subsets;
    int subset_size;
    B2PresampledLight* out_light_samples;
    float envmap_sampling_probability;
};
static_assert(sizeof(B2LightPresamplingParameters) == 24, "LightPresamplingParameters is 24 bytes");

/* One thread per presampled light: LightsPresampling.h:78-110 (presample_envmap :27-41, presample_emissive_triangle :43-64) */
/* ReSTIR_DI_LightPresamplingStrategy (ReSTIRDIOptions.h:162) = LSS_BASE_UNIFORM: the variant with that base strategy, for the two
 * places that read the macro (LightsPresampling.h:44, InitialCandidates.h:364) */
template <class V>
using B2UniformBase = B2Variant<V::BSDF_OVERRIDE, V::DIRECT, B2_BASE_UNIFORM, V::LATER, V::ENVMAP, V::COUNTERS>;

template <class V>
B2_HD void restir_di_presample_light(const b2pt_render_data& rd, const B2LightPresamplingParameters& params, uint32_t x, bool uniform_strategy = false)
{
    if (rd.buffers.emissive_triangles_count == 0 && rd.world_settings.ambient_light_type != 2)
        return;
    if (x >= (uint32_t)(params.subset_size * params.number_of_subsets))
        return;
    uint32_t seed = rd.render_settings.freeze_random ? wang_hash(x + 1) : wang_hash((x + 1) * (rd.render_settings.sample_number + 1) * rd.random_number);
    Rng rng = make_rng(seed);

    float envmap_candidate_probability = 0.0f;
    if (rd.world_settings.ambient_light_type == 2)
        envmap_candidate_probability = rd.buffers.emissive_triangles_count == 0 ? 1.0f : params.envmap_sampling_probability;

    B2PresampledLight light;
    light.emissive_triangle_index = -1;
    light.px = light.py = light.pz = 0.0f;
    light.nx = light.ny = light.nz = 0.0f;
    light.r = light.g = light.b = 0.0f;
    light.pdf = 0.0f;
    light.flags = B2_DI_FLAGS_NONE;
    if (rng.next_float() < envmap_candidate_probability)
    {
        light.flags |= B2_DI_FLAGS_ENVMAP_SAMPLE;
        V3 direction = v3(0.0f, 1.0f, 0.0f);
        float pdf = 0.0f;
        Rgb radiance = rd.world_settings.envmap ? envmap_sample<V>(rd.world_settings, direction, pdf, rng) : rgb(0.0f);
        V3 envmap_space = mat3_mul_vec(rd.world_settings.world_to_envmap_matrix, direction);
        light.px = envmap_space.x; light.py = envmap_space.y; light.pz = envmap_space.z;
        light.r = radiance.r; light.g = radiance.g; light.b = radiance.b;
        light.pdf = pdf * envmap_candidate_probability;
    }
    else
    {
        LightSample sample = uniform_strategy ? sample_one_emissive_triangle<B2UniformBase<V>>(rd, rng) : sample_one_emissive_triangle<V>(rd, rng);
        if (sample.area_measure_pdf > 0.0f)
        {
            light.px = sample.point_on_light.x; light.py = sample.point_on_light.y; light.pz = sample.point_on_light.z;
            light.nx = sample.light_source_normal.x; light.ny = sample.light_source_normal.y; light.nz = sample.light_source_normal.z;
            light.emissive_triangle_index = sample.emissive_triangle_index;
            light.pdf = sample.area_measure_pdf * (1.0f - envmap_candidate_probability);
            light.r = sample.emission.r; light.g = sample.emission.g; light.b = sample.emission.b;
        }
    }
    params.out_light_samples[x] = light;
}

/* use_presampled_light_candidate, InitialCandidates.h:33-90: every pixel of a tile_size x tile_size tile draws from the same
 * random subset of the presampled lights */
B2_HD DiSample use_presampled_light_candidate(const b2pt_render_data& rd, int px, int py, V3 evaluated_point, V3 shading_normal, Rgb& out_sample_radiance,
                                              float& out_sample_cosine_term, float& out_sample_pdf, Rng& rng)
{
    const b2pt_restir_di_settings& di = rd.render_settings.restir_di_settings;
    const int tx = px / di.light_presampling.tile_size, ty = py / di.light_presampling.tile_size;
    const int tile_index_seed = (tx + ty + 1) * (tx + ty) / 2 + ty; /* cantor_pairing_function */
    Rng subset_rng = make_rng(rd.random_number * (uint32_t)(tile_index_seed + 1));
    int random_subset_index = subset_rng.next_index(di.light_presampling.number_of_subsets);
    int random_light_index_in_subset = rng.next_index(di.light_presampling.subset_size);
    int light_sample_index = random_subset_index * di.light_presampling.subset_size + random_light_index_in_subset;
    const B2PresampledLight presampled = reinterpret_cast<const B2PresampledLight*>(di.light_presampling.light_samples)[light_sample_index];

    DiSample light_sample = di_sample_defaults();
    light_sample.emissive_triangle_index = presampled.emissive_triangle_index;
    light_sample.point_on_light_source = v3(presampled.px, presampled.py, presampled.pz);
    light_sample.flags = presampled.flags & 0xFFu;
    out_sample_radiance = rgb(presampled.r, presampled.g, presampled.b);
    out_sample_pdf = presampled.pdf;

    float distance_to_light;
    V3 to_light_direction = di_light_sample_direction(rd, light_sample, evaluated_point, distance_to_light);
    out_sample_cosine_term = dot(shading_normal, to_light_direction);
    if (!light_sample.is_envmap_sample())
    {
        bool contributes_enough = check_minimum_light_contribution(rd.render_settings.minimum_light_contribution, out_sample_radiance * out_sample_cosine_term / out_sample_pdf);
        if (!contributes_enough)
            light_sample.target_function = B2_LIGHT_DOESNT_CONTRIBUTE_ENOUGH;
    }
    return light_sample;
}

template <class V>
B2_HD DiSample sample_fresh_light_candidate(const Ctx& ctx, float envmap_candidate_probability, const HitInfo& hit, Rgb& out_sample_radiance, float& out_sample_cosine_term,
                                            float& out_sample_pdf, Rng& rng)
{
    const b2pt_render_data& rd = *ctx.rd;
    DiSample light_sample = di_sample_defaults();
    V3 evaluated_point = hit.inter_point;

    if (rng.next_float() > envmap_candidate_probability)
    {
        LightSample light_sample_info = sample_one_emissive_triangle<V>(rd, rng);
        light_sample.emissive_triangle_index = light_sample_info.emissive_triangle_index;
        light_sample.point_on_light_source = light_sample_info.point_on_light;
        out_sample_pdf = light_sample_info.area_measure_pdf;
        if (out_sample_pdf > 0.0f)
        {
            V3 to_light_direction = light_sample.point_on_light_source - evaluated_point;
            float distance_to_light = length(to_light_direction);
            to_light_direction = to_light_direction / distance_to_light;
            out_sample_cosine_term = b2_max(0.0f, dot(hit.shading_normal, to_light_direction));
            bool contributes_enough = check_minimum_light_contribution(rd.render_settings.minimum_light_contribution, light_sample_info.emission * out_sample_cosine_term / out_sample_pdf);
            if (!contributes_enough)
            {
                light_sample.target_function = B2_LIGHT_DOESNT_CONTRIBUTE_ENOUGH;
                return light_sample;
            }
            out_sample_pdf *= (1.0f - envmap_candidate_probability);
            out_sample_radiance = light_sample_info.emission;
        }
    }
    else
    {
        V3 envmap_sampled_direction = v3(0.0f, 1.0f, 0.0f);
        out_sample_radiance = rd.world_settings.envmap ? envmap_sample<V>(rd.world_settings, envmap_sampled_direction, out_sample_pdf, rng) : rgb(0.0f);
        out_sample_cosine_term = b2_max(0.0f, dot(envmap_sampled_direction, hit.shading_normal));
        bool contributes_enough = check_minimum_light_contribution(rd.render_settings.minimum_light_contribution, out_sample_radiance * out_sample_cosine_term / out_sample_pdf);
        if (!contributes_enough)
        {
            light_sample.target_function = B2_LIGHT_DOESNT_CONTRIBUTE_ENOUGH;
            return light_sample;
        }
        out_sample_pdf *= envmap_candidate_probability;
        light_sample.emissive_triangle_index = -1;
        light_sample.point_on_light_source = mat3_mul_vec(rd.world_settings.world_to_envmap_matrix, envmap_sampled_direction);
        light_sample.or_flags(B2_DI_FLAGS_ENVMAP_SAMPLE);
    }
    return light_sample;
}

B2_HD DiSurface surface_from_hit(const HitInfo& hit, const Payload& payload, V3 view_direction)
{
    DiSurface surface;
    surface.geometric_normal = hit.geometric_normal;
    surface.primitive_index = hit.primitive_index;
    surface.material = payload.material;
    surface.ray_volume_state = payload.volume_state;
    surface.shading_normal = hit.shading_normal;
    surface.shading_point = hit.inter_point;
    surface.view_direction = view_direction;
    return surface;
}

template <class V>
B2_HD bool restir_pass_prologue(const b2pt_render_data& rd, uint32_t x, uint32_t y, uint32_t& pixel_index)
{
    const b2pt_int2 res = rd.render_settings.render_resolution;
    if (x >= (uint32_t)res.x || y >= (uint32_t)res.y)
        return false;
    pixel_index = x + y * res.x;
    return true;
}

/* ---- InitialCandidates cut at its rays -------------------------------------------------------------------------------------
 * The pass traces two rays per pixel, and what follows each depends on its result: the closest hit along BSDF candidate 0 (it
 * becomes a reservoir candidate) and the visibility of the reservoir's final sample (ReSTIR_DI_DoVisibilityReuse). The per-pixel
 * code is therefore three phases with a ray request between them:
 *     phase A: light candidates, BSDF candidate 0 sampled            -> closest-hit ray
 *     phase B: BSDF candidate 0 resampled, reservoir.end()           -> shadow ray towards the reservoir's sample
 *     phase C: visibility reuse, store
 * restir_di_initial_candidates_pixel below runs A, trace, B, trace, C for one pixel (the form the host port, the one-thread-per-pixel
 * kernel and the goldens go through); the wavefront kernels of b2_kernels.cu run each phase over the whole frame and trace the
 * frame's rays together (b2_wavefront.h). Same operations and RNG draws in the same order either way. */
struct DiInitialState
{
    DiReservoir reservoir;
    Rng rng;
    V3 bsdf_sampled_direction;
    float bsdf_sample_pdf_solid_angle;
    int sampled_lobe_info;
    int nb_light_candidates, nb_bsdf_candidates;
    float envmap_candidate_probability;
    int active;        /* 0: the pixel takes no part in the pass (nothing is stored) */
    int pending_final; /* phase B requested the visibility ray */
};

/* what the three phases share: the pixel's primary hit rebuilt from the G-buffer (InitialCandidates.h:430-470) */
template <class V>
B2_HD void di_initial_surface(const b2pt_render_data& rd, uint32_t pixel_index, Rng& rng, HitInfo& hit, Payload& payload, V3& view_direction)
{
    hit = hit_info_defaults();
    hit.geometric_normal = unpack_normal_oct24(rd.g_buffer.geometric_normals[pixel_index]);
    hit.shading_normal = unpack_normal_oct24(rd.g_buffer.shading_normals[pixel_index]);
    hit.inter_point = load_v3(rd.g_buffer.primary_hit_position, pixel_index);
    hit.primitive_index = rd.g_buffer.first_hit_prim_index[pixel_index];

    payload.init();
    payload.material = unpack_material(load_packed_material(rd.g_buffer.materials, pixel_index));
    reconstruct_first_hit<V>(payload.volume_state, payload.material, rd.buffers.material_indices, hit.primitive_index, rng);

    V3 cam = v3(rd.current_camera.position.x, rd.current_camera.position.y, rd.current_camera.position.z);
    view_direction = normalize(cam - hit.inter_point);
}

template <class V, bool IN_ONE_PIECE = false>
B2_HD void restir_di_initial_phase_a(const Ctx& ctx, uint32_t x, uint32_t y, DiInitialState& st, RayRequest& candidate_ray)
{
    const b2pt_render_data& rd = *ctx.rd;
    const b2pt_restir_di_settings& di = rd.render_settings.restir_di_settings;
    request_none(candidate_ray);
    st.active = 0;
    st.pending_final = 0;
    if (rd.buffers.emissive_triangles_count == 0 && rd.world_settings.ambient_light_type != 2)
        return;
    uint32_t pixel_index;
    if (!restir_pass_prologue<V>(rd, x, y, pixel_index))
        return;

    Rng rng = make_rng(pixel_seed(rd, pixel_index));
    if (!rd.aux_buffers.pixel_active[pixel_index] || rd.g_buffer.first_hit_prim_index[pixel_index] == -1)
        return;
    st.active = 1;

    HitInfo hit;
    Payload payload;
    V3 view_direction;
    di_initial_surface<V>(rd, pixel_index, rng, hit, payload, view_direction);

    /* sample_initial_candidates (InitialCandidates.h:150-426) */
    int initial_nb_light_cand = di.initial_candidates.number_of_initial_light_candidates;
    int initial_nb_bsdf_cand = di.initial_candidates.number_of_initial_bsdf_candidates;
    const bool low_res = do_render_low_resolution(rd.render_settings);
    int nb_light_candidates = low_res ? b2_min(1, initial_nb_light_cand) : initial_nb_light_cand;
    int nb_bsdf_candidates = low_res ? b2_min(1, initial_nb_bsdf_cand) : initial_nb_bsdf_cand;
    float envmap_candidate_probability = 0.0f;
    if (rd.world_settings.ambient_light_type == 2)
    {
        if (rd.buffers.emissive_triangles_count == 0)
            envmap_candidate_probability = 1.0f;
        else
            envmap_candidate_probability = di.initial_candidates.envmap_candidate_probability;
    }
    st.nb_light_candidates = nb_light_candidates;
    st.nb_bsdf_candidates = nb_bsdf_candidates;
    st.envmap_candidate_probability = envmap_candidate_probability;

    DiReservoir reservoir;
    reservoir.init();

    /* ---- light candidates ---- */
    for (int i = 0; i < nb_light_candidates; i++)
    {
        Rgb sample_radiance = rgb(0.0f);
        float sample_cosine_term = 0.0f;
        float light_pdf_area_measure = 0.0f;
        float distance_to_light = 0.0f;
        V3 to_light_direction = v3(0.0f);

        /* ReSTIR_DI_DoLightPresampling (a -D macro in the reference, InitialCandidates.h:192-211) is a launch-time switch here */
        DiSample light_sample = ctx.light_presampling
                                    ? use_presampled_light_candidate(rd, (int)x, (int)y, hit.inter_point, hit.shading_normal, sample_radiance, sample_cosine_term, light_pdf_area_measure, rng)
                                    : sample_fresh_light_candidate<V>(ctx, envmap_candidate_probability, hit, sample_radiance, sample_cosine_term, light_pdf_area_measure, rng);
        to_light_direction = di_light_sample_direction(rd, light_sample, hit.inter_point, distance_to_light);

        if (light_sample.target_function == B2_LIGHT_DOESNT_CONTRIBUTE_ENOUGH)
            continue;

        float candidate_weight = 0.0f;
        if (sample_cosine_term > 0.0f && light_pdf_area_measure > 0.0f)
        {
            float bsdf_pdf_solid_angle;
            int incident_light_info = B2_LOBE_NO_INFO; /* uninitialised in the reference (InitialCandidates.h:226) */
            BsdfCtx bctx = make_bsdf_ctx(view_direction, hit.shading_normal, hit.geometric_normal, to_light_direction, &incident_light_info, &payload.volume_state, false, &payload.material,
                                         payload.bounce, payload.accumulated_roughness);
            const Rgb candidate_bsdf_color = bsdf_eval<V>(rd, bctx, bsdf_pdf_solid_angle);

            DiSurface surface = surface_from_hit(hit, payload, view_direction);
            /* same surface, same direction (di_light_sample_direction both times), lobe info NO_INFO (a fresh or presampled light candidate
             * carries no lobe bits), bounce 0, accumulated roughness 0 at the primary hit: the target function's own evaluation
             * would return candidate_bsdf_color again */
            const bool same_context = payload.bounce == 0 && payload.accumulated_roughness == 0.0f && light_sample.lobe_info() == B2_LOBE_NO_INFO;
            float target_function = di_target_function<V, false>(ctx, light_sample, surface, same_context ? &candidate_bsdf_color : nullptr);
            if (bsdf_pdf_solid_angle <= 0.0f || !check_minimum_light_contribution(rd.render_settings.minimum_light_contribution, target_function / light_pdf_area_measure / bsdf_pdf_solid_angle))
                target_function = 0.0f;
            else
            {
                float light_pdf_solid_angle;
                if (light_sample.is_envmap_sample())
                    light_pdf_solid_angle = light_pdf_area_measure;
                else
                {
                    V3 light_normal = triangle_normal_not_normalized(rd, light_sample.emissive_triangle_index);
                    float normal_length = length(light_normal);
                    light_normal = light_normal / normal_length;
                    light_pdf_solid_angle = area_to_solid_angle_pdf(light_pdf_area_measure / (1.0f - envmap_candidate_probability), distance_to_light,
                                                                    cosine_term_at_light_source(light_normal, -to_light_direction));
                    light_pdf_solid_angle *= (1.0f - envmap_candidate_probability);
                }
                float mis_weight = balance_heuristic(light_pdf_solid_angle, (float)nb_light_candidates, bsdf_pdf_solid_angle, (float)nb_bsdf_candidates);
                candidate_weight = mis_weight * target_function / light_pdf_area_measure;
                light_sample.target_function = target_function;
            }
        }
        /* ReSTIR_DI_InitialTargetFunctionVisibility (InitialCandidates.h:266-291): a shadow ray per light candidate, traced on the spot --
         * the per-pixel form of the pass only, the wavefront form is launched for the default (off) */
        if constexpr (IN_ONE_PIECE)
        {
            if ((ctx.restir_options & B2_DI_OPT_INITIAL_VISIBILITY) && !do_render_low_resolution(rd.render_settings) && light_sample.target_function > 0.0f)
            {
                bool visible = !shadow_ray_occluded(ctx, hit.inter_point, to_light_direction, distance_to_light, hit.primitive_index);
                if (!visible)
                {
                    reservoir.M++;
                    continue;
                }
                light_sample.or_flags(B2_DI_FLAGS_UNOCCLUDED);
            }
        }
        reservoir.add_one_candidate(light_sample, candidate_weight, rng);
    }

    /* ---- BSDF candidate 0: the sampling half ---- */
    st.bsdf_sample_pdf_solid_angle = 0.0f;
    st.bsdf_sampled_direction = v3(0.0f);
    st.sampled_lobe_info = B2_LOBE_NO_INFO;
    if (nb_bsdf_candidates > 0)
    {
        BsdfCtx bctx = make_bsdf_ctx(view_direction, hit.shading_normal, hit.geometric_normal, v3(0.0f), &st.sampled_lobe_info, &payload.volume_state, false, &payload.material, 0,
                                     payload.accumulated_roughness);
        (void)bsdf_sample<V>(rd, bctx, st.bsdf_sampled_direction, st.bsdf_sample_pdf_solid_angle, rng);
        if (st.bsdf_sample_pdf_solid_angle > 0.0f)
            request_closest(candidate_ray, hit.inter_point, st.bsdf_sampled_direction, hit.primitive_index);
    }
    st.reservoir = reservoir;
    st.rng = rng;
}

/* one BSDF candidate, given the closest hit along its direction (InitialCandidates.h:296-420) */
template <class V>
B2_HD void di_initial_bsdf_candidate_finish(const Ctx& ctx, const Payload& payload, const HitInfo& hit, V3 view_direction, V3 bsdf_sampled_direction, float bsdf_sample_pdf_solid_angle,
                                            int sampled_lobe_info, bool hit_found, const LightRayHit& light_hit, int nb_light_candidates, int nb_bsdf_candidates,
                                            float envmap_candidate_probability, DiReservoir& reservoir, Rng& rng)
{
    const b2pt_render_data& rd = *ctx.rd;
    if (hit_found && !is_black(light_hit.emission))
    {
        DiSurface surface = surface_from_hit(hit, payload, view_direction);
        DiSample bsdf_RIS_sample = di_sample_defaults();
        bsdf_RIS_sample.emissive_triangle_index = light_hit.prim;
        bsdf_RIS_sample.point_on_light_source = hit.inter_point + bsdf_sampled_direction * light_hit.distance;
        bsdf_RIS_sample.or_flags(B2_DI_FLAGS_UNOCCLUDED);
        bsdf_RIS_sample.or_flags(sampled_lobe_info);
        bsdf_RIS_sample.target_function = di_target_function<V, false>(ctx, bsdf_RIS_sample, surface);

        float light_pdf_solid_angle = 0.0f;
        bool refraction_sampled = dot(bsdf_sampled_direction, hit.shading_normal) < 0.0f;
        if (!refraction_sampled)
            light_pdf_solid_angle = ctx.light_presampling == 2 ? pdf_of_emissive_triangle_hit_solid_angle<B2UniformBase<V>>(rd, light_hit, bsdf_sampled_direction)
                                                               : pdf_of_emissive_triangle_hit_solid_angle<V>(rd, light_hit, bsdf_sampled_direction);

        float target_function = bsdf_RIS_sample.target_function;
        if (bsdf_sample_pdf_solid_angle <= 0.0f || !check_minimum_light_contribution(rd.render_settings.minimum_light_contribution, target_function / light_pdf_solid_angle / bsdf_sample_pdf_solid_angle))
            return;

        light_pdf_solid_angle *= (1.0f - envmap_candidate_probability);
        float mis_weight = balance_heuristic(bsdf_sample_pdf_solid_angle, (float)nb_bsdf_candidates, light_pdf_solid_angle, (float)nb_light_candidates);

        float bsdf_sample_pdf_area_measure = bsdf_sample_pdf_solid_angle;
        bsdf_sample_pdf_area_measure /= (light_hit.distance * light_hit.distance);
        bsdf_sample_pdf_area_measure *= cosine_term_at_light_source(light_hit.geometric_normal, -bsdf_sampled_direction);

        float candidate_weight = mis_weight * target_function / bsdf_sample_pdf_area_measure;
        reservoir.add_one_candidate(bsdf_RIS_sample, candidate_weight, rng);
    }
    else if (!hit_found && rd.world_settings.ambient_light_type == 2)
    {
        if (dot(hit.shading_normal, bsdf_sampled_direction) > 0.0f)
        {
            float envmap_pdf;
            (void)envmap_eval<V>(rd.world_settings, bsdf_sampled_direction, envmap_pdf);

            DiSurface surface = surface_from_hit(hit, payload, view_direction);
            DiSample bsdf_RIS_sample = di_sample_defaults();
            bsdf_RIS_sample.emissive_triangle_index = -1;
            bsdf_RIS_sample.point_on_light_source = mat3_mul_vec(rd.world_settings.world_to_envmap_matrix, bsdf_sampled_direction);
            bsdf_RIS_sample.or_flags(B2_DI_FLAGS_UNOCCLUDED);
            bsdf_RIS_sample.or_flags(B2_DI_FLAGS_ENVMAP_SAMPLE);
            bsdf_RIS_sample.or_flags(sampled_lobe_info);
            bsdf_RIS_sample.target_function = di_target_function<V, false>(ctx, bsdf_RIS_sample, surface);

            float target_function = bsdf_RIS_sample.target_function;
            if (bsdf_sample_pdf_solid_angle <= 0.0f || !check_minimum_light_contribution(rd.render_settings.minimum_light_contribution, target_function / bsdf_sample_pdf_solid_angle))
                return;

            envmap_pdf *= envmap_candidate_probability;
            float mis_weight = balance_heuristic(bsdf_sample_pdf_solid_angle, (float)nb_bsdf_candidates, envmap_pdf, (float)nb_light_candidates);
            float candidate_weight = mis_weight * target_function / bsdf_sample_pdf_solid_angle;
            reservoir.add_one_candidate(bsdf_RIS_sample, candidate_weight, rng);
        }
    }
}

/* the shadow ray of the visibility-reuse test (ReSTIR/Utils.h:25-65), or none when the test would not trace */
B2_HD bool di_visibility_test_request(const b2pt_render_data& rd, const DiReservoir& reservoir, V3 shading_point, int last_primitive_hit_index, RayRequest& ray)
{
    request_none(ray);
    if (reservoir.UCW <= 0.0f && reservoir.weight_sum <= 0.0f)
        return false;
    else if (reservoir.sample.flags & B2_DI_FLAGS_UNOCCLUDED)
        return false;
    float distance_to_light;
    V3 sample_direction = di_light_sample_direction(rd, reservoir.sample, shading_point, distance_to_light);
    request_shadow(ray, shading_point, sample_direction, distance_to_light, last_primitive_hit_index);
    return true;
}
B2_HD void di_visibility_test_apply(DiReservoir& reservoir, bool occluded)
{
    if (occluded)
        reservoir.UCW = -1.0f;
    else
        reservoir.sample.or_flags(B2_DI_FLAGS_UNOCCLUDED);
}

template <class V>
B2_HD void restir_di_initial_phase_b(const Ctx& ctx, uint32_t x, uint32_t y, DiInitialState& st, const B2Hit& candidate_hit, RayRequest& visibility_ray)
{
    const b2pt_render_data& rd = *ctx.rd;
    request_none(visibility_ray);
    if (!st.active)
        return;
    const uint32_t pixel_index = x + y * (uint32_t)rd.render_settings.render_resolution.x;
    HitInfo hit;
    Payload payload;
    V3 view_direction;
    Rng surface_rng = make_rng(pixel_seed(rd, pixel_index)); /* reconstruct_first_hit's draw (dispersive glass only) already happened in phase A */
    di_initial_surface<V>(rd, pixel_index, surface_rng, hit, payload, view_direction);

    DiReservoir reservoir = st.reservoir;
    Rng rng = st.rng;
    if (st.nb_bsdf_candidates > 0 && st.bsdf_sample_pdf_solid_angle > 0.0f)
    {
        LightRayHit light_hit;
        bool hit_found = light_ray_hit_from(rd, candidate_hit, 1.0e35f, light_hit);
        di_initial_bsdf_candidate_finish<V>(ctx, payload, hit, view_direction, st.bsdf_sampled_direction, st.bsdf_sample_pdf_solid_angle, st.sampled_lobe_info, hit_found, light_hit,
                                            st.nb_light_candidates, st.nb_bsdf_candidates, st.envmap_candidate_probability, reservoir, rng);
    }
    /* BSDF candidates beyond the first (not the default): traced in place, one after the other */
    for (int i = 1; i < st.nb_bsdf_candidates; i++)
    {
        float bsdf_sample_pdf_solid_angle = 0.0f;
        V3 bsdf_sampled_direction;
        int sampled_lobe_info = B2_LOBE_NO_INFO;
        BsdfCtx bctx = make_bsdf_ctx(view_direction, hit.shading_normal, hit.geometric_normal, v3(0.0f), &sampled_lobe_info, &payload.volume_state, false, &payload.material, 0,
                                     payload.accumulated_roughness);
        (void)bsdf_sample<V>(rd, bctx, bsdf_sampled_direction, bsdf_sample_pdf_solid_angle, rng);
        if (bsdf_sample_pdf_solid_angle > 0.0f)
        {
            LightRayHit light_hit;
            bool hit_found = bsdf_light_sample_ray(ctx, hit.inter_point, bsdf_sampled_direction, 1.0e35f, light_hit, hit.primitive_index);
            di_initial_bsdf_candidate_finish<V>(ctx, payload, hit, view_direction, bsdf_sampled_direction, bsdf_sample_pdf_solid_angle, sampled_lobe_info, hit_found, light_hit,
                                                st.nb_light_candidates, st.nb_bsdf_candidates, st.envmap_candidate_probability, reservoir, rng);
        }
    }
    reservoir.end();
    reservoir.M = 1;

    st.reservoir = reservoir;
    st.rng = rng;
    /* ReSTIR_DI_DoVisibilityReuse (InitialCandidates.h:533-535) */
    st.pending_final = 0;
    if (ctx.restir_options & B2_DI_OPT_VISIBILITY_REUSE)
        st.pending_final = di_visibility_test_request(rd, reservoir, hit.inter_point, hit.primitive_index, visibility_ray) ? 1 : 0;
}

template <class V>
B2_HD void restir_di_initial_phase_c(const Ctx& ctx, uint32_t x, uint32_t y, DiInitialState& st, const B2Hit& visibility_hit)
{
    const b2pt_render_data& rd = *ctx.rd;
    if (!st.active)
        return;
    const uint32_t pixel_index = x + y * (uint32_t)rd.render_settings.render_resolution.x;
    if (st.pending_final)
        di_visibility_test_apply(st.reservoir, visibility_hit.prim >= 0);
    di_store(rd, rd.render_settings.restir_di_settings.initial_candidates.output_reservoirs, (int)pixel_index, st.reservoir);
}

/* a requested ray traced on the spot (the per-pixel form of the passes) */
B2_HD B2Hit di_trace_request(const Ctx& ctx, const RayRequest& ray)
{
    B2Hit none;
    none.t = -1.0f;
    none.u = none.v = 0.0f;
    none.prim = -1;
    none.ng = v3(0.0f);
    if (ray.kind == B2_RAY_NONE)
        return none;
    return b2_trace(ctx, ray.kind == B2_RAY_ANY_HIT, ray.origin, ray.direction, ray.t_max, ray.skip_prim);
}

template <class V>
B2_HD void restir_di_initial_candidates_pixel(const Ctx& ctx, uint32_t x, uint32_t y)
{
    DiInitialState st;
    RayRequest ray;
    restir_di_initial_phase_a<V, true>(ctx, x, y, st, ray);
    B2Hit result = di_trace_request(ctx, ray);
    restir_di_initial_phase_b<V>(ctx, x, y, st, result, ray);
    result = di_trace_request(ctx, ray);
    restir_di_initial_phase_c<V>(ctx, x, y, st, result);
}

/* ---------------------------------------------------------------------------------------------------------- */
/* temporal reuse (TemporalReuse.h) */
template <class V>
B2_HD void restir_di_temporal_pixel(const Ctx& ctx, uint32_t x, uint32_t y)
{
    const b2pt_render_data& rd = *ctx.rd;
    const b2pt_restir_di_settings& di = rd.render_settings.restir_di_settings;
    uint32_t center_pixel_index;
    if (!restir_pass_prologue<V>(rd, x, y, center_pixel_index))
        return;
    if (!rd.aux_buffers.pixel_active[center_pixel_index] || rd.g_buffer.first_hit_prim_index[center_pixel_index] == -1)
        return;

    Rng rng = make_rng(pixel_seed(rd, center_pixel_index));

    if (di.common_temporal_pass.temporal_buffer_clear_requested)
    {
        DiReservoir empty;
        empty.init();
        di_store(rd, di.temporal_pass.input_reservoirs, (int)center_pixel_index, empty);
    }

    DiSurface center_pixel_surface = get_pixel_surface<V>(rd, (int)center_pixel_index, false, rng);
    DiReservoir initial_candidates_reservoir = di_load(rd, di.initial_candidates.output_reservoirs, (int)center_pixel_index);

    int temporal_neighbor_pixel_index = find_temporal_neighbor_index<V>(rd, load_v3(rd.g_buffer.primary_hit_position, center_pixel_index), center_pixel_surface.shading_normal,
                                                                        (int)center_pixel_index, rng).x;
    if (temporal_neighbor_pixel_index == -1 || rd.render_settings.freeze_random)
    {
        di_store(rd, di.temporal_pass.output_reservoirs, (int)center_pixel_index, initial_candidates_reservoir);
        return;
    }

    DiReservoir temporal_neighbor_reservoir = di_load(rd, di.temporal_pass.input_reservoirs, temporal_neighbor_pixel_index);
    if (temporal_neighbor_reservoir.M == 0)
    {
        di_store(rd, di.temporal_pass.output_reservoirs, (int)center_pixel_index, initial_candidates_reservoir);
        return;
    }

    DiReservoir out;
    out.init();
    DiSurface temporal_neighbor_surface = get_pixel_surface<V>(rd, temporal_neighbor_pixel_index, use_prev_frame_g_buffer<V>(rd.render_settings), rng);
    if (mat_is_emissive(temporal_neighbor_surface.material))
    {
        di_store(rd, di.temporal_pass.output_reservoirs, (int)center_pixel_index, initial_candidates_reservoir);
        return;
    }

    /* ReSTIR_DI_BiasCorrectionWeights: how the two reservoirs are weighted against each other (TemporalMISWeight.h) and how the result is
     * normalised (TemporalNormalizationWeight.h); a -D macro in the reference, a launch-time switch here */
    const int bias = ctx.bias_correction;
    const bool use_confidence_weights = di.use_confidence_weights != 0;
    int selected_neighbor = 0; /* TEMPORAL_NEIGHBOR_ID = 0, INITIAL_CANDIDATES_ID = 1 */
    float mc = 0.0f;

    /* generalised balance heuristic (TemporalMISWeight.h:93-160). As in the reference, BOTH target functions are evaluated on the
     * temporal neighbour's surface (:129 passes temporal_neighbor_surface where the centre pixel's is meant) */
    auto gbh_weight = [&](const DiSample& sample, int current_neighbor_index) {
        float target_function_at_temporal_neighbor = 0.0f;
        if (temporal_neighbor_reservoir.M != 0)
            target_function_at_temporal_neighbor = di_target_function_bias<V>(ctx, sample, temporal_neighbor_surface);
        if (current_neighbor_index == 0 && target_function_at_temporal_neighbor == 0.0f)
            return 0.0f;
        float target_function_at_center = di_target_function_bias<V>(ctx, sample, temporal_neighbor_surface);
        int temporal_M = temporal_neighbor_reservoir.M;
        int center_reservoir_M = initial_candidates_reservoir.M;
        if (!use_confidence_weights)
        {
            temporal_M = 1;
            center_reservoir_M = 1;
        }
        float nume = current_neighbor_index == 0 ? target_function_at_temporal_neighbor * temporal_M : target_function_at_center * center_reservoir_M;
        float denom = target_function_at_temporal_neighbor * temporal_M + target_function_at_center * center_reservoir_M;
        return denom == 0.0f ? 0.0f : nume / denom;
    };

    if (temporal_neighbor_reservoir.M > 0)
    {
        float target_function_at_center = 0.0f;
        if (temporal_neighbor_reservoir.UCW > 0.0f)
            target_function_at_center = di_target_function_bias<V>(ctx, temporal_neighbor_reservoir.sample, center_pixel_surface);

        float temporal_neighbor_resampling_mis_weight;
        if (bias == B2_BIAS_1_OVER_Z || bias == B2_BIAS_1_OVER_M)
            temporal_neighbor_resampling_mis_weight = (float)temporal_neighbor_reservoir.M;
        else if (bias == B2_BIAS_MIS_LIKE)
            temporal_neighbor_resampling_mis_weight = use_confidence_weights ? (float)temporal_neighbor_reservoir.M : 1.0f;
        else if (bias == B2_BIAS_MIS_GBH)
            temporal_neighbor_resampling_mis_weight = gbh_weight(temporal_neighbor_reservoir.sample, 0);
        else if (bias == B2_BIAS_SYMMETRIC_RATIO || bias == B2_BIAS_ASYMMETRIC_RATIO)
        {
            /* symmetric ratio (TemporalMISWeight.h:367-462) and asymmetric ratio (:464-595); no guard on the divisions, as there */
            const bool asymmetric = bias == B2_BIAS_ASYMMETRIC_RATIO;
            const float beta = di.symmetric_ratio_mis_weights_beta_exponent;
            float target_function_neighbor_sample_at_neighbor = temporal_neighbor_reservoir.sample.target_function;
            float temporal_neighbor_M = use_confidence_weights ? (float)temporal_neighbor_reservoir.M : 1.0f;
            float center_reservoir_M = use_confidence_weights ? (float)initial_candidates_reservoir.M : 1.0f;
            float neighbors_confidence_sum = use_confidence_weights ? temporal_neighbor_M : 1.0f;

            float difference_function = ratio_mis_difference_function(target_function_at_center, target_function_neighbor_sample_at_neighbor, beta);
            float nume_mi = difference_function * temporal_neighbor_M;
            float denom_mi = center_reservoir_M + neighbors_confidence_sum * difference_function;
            if (asymmetric && !(target_function_at_center <= target_function_neighbor_sample_at_neighbor))
                denom_mi = center_reservoir_M + neighbors_confidence_sum;
            float mi = nume_mi / denom_mi;

            float target_function_center_sample_at_neighbor = di_target_function_bias<V>(ctx, initial_candidates_reservoir.sample, temporal_neighbor_surface);
            float target_function_center_sample_at_center = initial_candidates_reservoir.sample.target_function;
            float difference_function_mc = ratio_mis_difference_function(target_function_center_sample_at_neighbor, target_function_center_sample_at_center, beta);
            if (asymmetric)
            {
                float nume_mc = difference_function_mc * temporal_neighbor_M;
                float denom_mc = center_reservoir_M + neighbors_confidence_sum * difference_function_mc;
                if (!(target_function_center_sample_at_center <= target_function_center_sample_at_neighbor))
                    denom_mc = center_reservoir_M + neighbors_confidence_sum;
                mc += nume_mc / denom_mc;
            }
            else
            {
                float nume_mc = center_reservoir_M;
                float denom_mc = center_reservoir_M + neighbors_confidence_sum * difference_function_mc;
                float confidence_weights_multiplier = 1.0f;
                if (use_confidence_weights)
                    confidence_weights_multiplier = neighbors_confidence_sum == 0.0f ? 0.0f : temporal_neighbor_M / neighbors_confidence_sum;
                mc += confidence_weights_multiplier * nume_mc / denom_mc;
            }
            temporal_neighbor_resampling_mis_weight = mi;
        }
        else
        {
            /* pairwise MIS (TemporalMISWeight.h:163-254) and its defensive form (:256-365) */
            const bool defensive = bias == B2_BIAS_PAIRWISE_MIS_DEFENSIVE;
            float target_function_at_neighbor = temporal_neighbor_reservoir.sample.target_function;
            float temporal_neighbor_M = use_confidence_weights ? (float)temporal_neighbor_reservoir.M : 1.0f;
            float center_reservoir_M = use_confidence_weights ? (float)initial_candidates_reservoir.M : 1.0f;
            float neighbors_confidence_sum = use_confidence_weights ? (float)temporal_neighbor_reservoir.M : 1.0f;

            float nume = target_function_at_neighbor * temporal_neighbor_M;
            float denom = target_function_at_neighbor * neighbors_confidence_sum + target_function_at_center * center_reservoir_M;
            float mi = denom == 0.0f ? 0.0f : (nume / denom);
            if (defensive && use_confidence_weights)
                mi *= neighbors_confidence_sum / (neighbors_confidence_sum + center_reservoir_M);

            float target_function_center_sample_at_neighbor = di_target_function_bias<V>(ctx, initial_candidates_reservoir.sample, temporal_neighbor_surface);
            float target_function_center_sample_at_center = initial_candidates_reservoir.sample.target_function;
            float nume_mc = target_function_center_sample_at_center * center_reservoir_M;
            float denom_mc = target_function_center_sample_at_neighbor * neighbors_confidence_sum + target_function_center_sample_at_center * center_reservoir_M;
            float confidence_multiplier = 1.0f;
            if (use_confidence_weights)
                confidence_multiplier = defensive ? neighbors_confidence_sum / (neighbors_confidence_sum + center_reservoir_M) : temporal_neighbor_M / neighbors_confidence_sum;
            if (denom_mc != 0.0f)
                mc += nume_mc / denom_mc * confidence_multiplier;

            temporal_neighbor_resampling_mis_weight = (defensive && !use_confidence_weights) ? mi * 0.5f : mi;
        }

        if (out.combine_with(temporal_neighbor_reservoir, temporal_neighbor_resampling_mis_weight, target_function_at_center, 1.0f, rng))
        {
            selected_neighbor = 0;
            /* TemporalReuse.h:193-203: with visibility in the target function the neighbour could only be chosen unoccluded */
            if (ctx.restir_options & B2_DI_OPT_BIAS_VISIBILITY)
                out.sample.or_flags(B2_DI_FLAGS_UNOCCLUDED);
            else
                out.sample.flags &= ~(uint32_t)B2_DI_FLAGS_UNOCCLUDED;
        }
    }

    float initial_candidates_mis_weight;
    if (bias == B2_BIAS_1_OVER_Z || bias == B2_BIAS_1_OVER_M)
        initial_candidates_mis_weight = (float)initial_candidates_reservoir.M;
    else if (bias == B2_BIAS_MIS_LIKE)
        initial_candidates_mis_weight = use_confidence_weights ? (float)initial_candidates_reservoir.M : 1.0f;
    else if (bias == B2_BIAS_MIS_GBH)
        initial_candidates_mis_weight = gbh_weight(initial_candidates_reservoir.sample, 1);
    else if (bias == B2_BIAS_PAIRWISE_MIS_DEFENSIVE)
    {
        if (mc == 0.0f)
            initial_candidates_mis_weight = 1.0f;
        else if (use_confidence_weights)
            initial_candidates_mis_weight = mc + (float)initial_candidates_reservoir.M / (float)(initial_candidates_reservoir.M + temporal_neighbor_reservoir.M);
        else
            initial_candidates_mis_weight = (1.0f + mc) * 0.5f;
    }
    else if (bias == B2_BIAS_ASYMMETRIC_RATIO)
        initial_candidates_mis_weight = mc == 0.0f ? 1.0f : 1.0f - mc;
    else
        initial_candidates_mis_weight = mc == 0.0f ? 1.0f : mc;
    if (out.combine_with(initial_candidates_reservoir, initial_candidates_mis_weight, initial_candidates_reservoir.sample.target_function, 1.0f, rng))
    {
        selected_neighbor = 1;
        if (ctx.restir_options & B2_DI_OPT_BIAS_VISIBILITY) /* TemporalReuse.h:248-256 */
            out.sample.or_flags(B2_DI_FLAGS_UNOCCLUDED);
        else
            out.sample.or_flags((int)(initial_candidates_reservoir.sample.flags & B2_DI_FLAGS_UNOCCLUDED));
    }

    /* TemporalNormalizationWeight.h: 1/Z counts the confidence of the reservoirs that could have produced the chosen sample, MIS-like
     * weighs the chosen one against both; the others put everything into the resampling weights */
    float normalization_numerator = 1.0f, normalization_denominator = 1.0f;
    if (bias == B2_BIAS_1_OVER_M && !(out.weight_sum <= 0.0f))
        normalization_denominator = (float)(initial_candidates_reservoir.M + temporal_neighbor_reservoir.M); /* TemporalNormalizationWeight.h:38-54: every sample counted, biased */
    if ((bias == B2_BIAS_1_OVER_Z || bias == B2_BIAS_MIS_LIKE) && !(out.weight_sum <= 0.0f))
    {
        float center_pixel_target_function = di_target_function_bias<V>(ctx, out.sample, center_pixel_surface);
        float temporal_neighbor_target_function = 0.0f;
        if (bias == B2_BIAS_1_OVER_Z)
        {
            normalization_denominator = 0.0f;
            normalization_denominator += (float)((center_pixel_target_function > 0.0f) * initial_candidates_reservoir.M);
            if (temporal_neighbor_reservoir.M > 0)
            {
                temporal_neighbor_target_function = di_target_function_bias<V>(ctx, out.sample, temporal_neighbor_surface);
                normalization_denominator += (float)((temporal_neighbor_target_function > 0.0f) * temporal_neighbor_reservoir.M);
            }
        }
        else
        {
            if (temporal_neighbor_reservoir.M > 0)
                temporal_neighbor_target_function = di_target_function_bias<V>(ctx, out.sample, temporal_neighbor_surface);
            normalization_numerator = selected_neighbor == 1 ? center_pixel_target_function : temporal_neighbor_target_function;
            int initial_candidates_M = initial_candidates_reservoir.M, temporal_neighbor_M = temporal_neighbor_reservoir.M;
            if (!use_confidence_weights)
            {
                initial_candidates_M = 1;
                temporal_neighbor_M = 1;
            }
            normalization_denominator = center_pixel_target_function * initial_candidates_M + temporal_neighbor_target_function * temporal_neighbor_M;
        }
    }

    out.end_with_normalization(normalization_numerator, normalization_denominator);
    if (di.m_cap > 0)
        out.M = b2_min(out.M, di.m_cap);
    di_store(rd, di.temporal_pass.output_reservoirs, (int)center_pixel_index, out);
}

/* ---------------------------------------------------------------------------------------------------------- */
/* spatial reuse (SpatialReuse.h) */
template <class V>
B2_HD void restir_di_spatial_pixel(const Ctx& ctx, uint32_t x, uint32_t y)
{
    const b2pt_render_data& rd = *ctx.rd;
    const b2pt_restir_di_settings& di = rd.render_settings.restir_di_settings;
    uint32_t center_pixel_index;
    if (!restir_pass_prologue<V>(rd, x, y, center_pixel_index))
        return;
    if (!rd.aux_buffers.pixel_active[center_pixel_index] || rd.g_buffer.first_hit_prim_index[center_pixel_index] == -1)
        return;

    unsigned int seed;
    if (rd.render_settings.freeze_random)
        seed = wang_hash(center_pixel_index + 1);
    else if (restir_gi_coalesced_spatial_reuse(rd))
        seed = wang_hash((rd.render_settings.sample_number + 1) * rd.random_number);
    else
        seed = wang_hash(((center_pixel_index + 1) * (rd.render_settings.sample_number + 1)) * rd.random_number);
    Rng rng = make_rng(seed);

    const void* input_reservoir_buffer = di.spatial_pass.input_reservoirs;
    DiReservoir out;
    out.init();
    SpatialLocal local = spatial_local_from(rd);

    DiReservoir center_pixel_reservoir = di_load(rd, input_reservoir_buffer, (int)center_pixel_index);
    if ((center_pixel_reservoir.M <= 1) && di.common_spatial_pass.do_disocclusion_reuse_boost)
        local.reuse_neighbor_count = di.common_spatial_pass.disocclusion_reuse_count;

    DiSurface center_pixel_surface = get_pixel_surface<V>(rd, (int)center_pixel_index, false, rng);
    setup_adaptive_directional_spatial_reuse(rd, local, center_pixel_index, rng);

    int neighbor_heuristics_cache = 0;
    int valid_neighbors_count = 0;
    int valid_neighbors_M_sum = 0;
    count_valid_neighbors(rd, local, center_pixel_surface, (int)center_pixel_index, (int)x, (int)y, false, input_reservoir_buffer, valid_neighbors_count, valid_neighbors_M_sum,
                          neighbor_heuristics_cache);

    PairwiseMis mis;
    mis.mc = 0.0f;
    Rng spatial_neighbors_rng = make_rng(local.rng_seed);
    const int bias = ctx.bias_correction;
    const bool use_confidence_weights = di.use_confidence_weights != 0;
    int selected_neighbor = 0;

    /* The families that walk over ALL neighbours again -- the generalised balance heuristic for every resampled neighbour, 1/Z and
     * MIS-like for the normalisation -- regenerate the neighbour positions by re-seeding the kernel's MAIN generator with the
     * spatial-neighbours seed (SpatialMISWeight.h:66-67, SpatialNormalizationWeight.h:88,158); the balance heuristic's restore of
     * the old seed sits behind its return statements (:113-118) and never runs, so whatever the kernel draws next -- the reservoir's
     * combine_with -- comes from the re-seeded stream. Mirrored as written. `evaluate(j, neighbour pixel, its surface)` is called
     * for every neighbour that passes the similarity heuristics. */
    auto for_all_neighbors = [&](auto&& evaluate) {
        rng = make_rng(local.rng_seed);
        for (int j = 0; j < local.reuse_neighbor_count + 1; j++)
        {
            int neighbor_index_j = get_spatial_neighbor_pixel_index(rd, local, j, (int)x, (int)y, rng);
            if (neighbor_index_j == -1)
                continue;
            if (!check_neighbor_similarity_heuristics(rd, neighbor_index_j, (int)center_pixel_index, center_pixel_surface.shading_point, normal_for_rejection_heuristic(rd, center_pixel_surface), false))
                continue;
            DiSurface neighbor_surface = get_pixel_surface<V>(rd, neighbor_index_j, false, rng);
            evaluate(j, neighbor_index_j, neighbor_surface);
        }
    };
    /* SpatialMISWeight.h:46-117 */
    auto gbh_weight = [&](float reservoir_being_resampled_UCW, const DiSample& reservoir_being_resampled_sample, int current_neighbor_index) {
        if (reservoir_being_resampled_UCW <= 0.0f)
            return 1.0f;
        float nume = 0.0f, denom = 0.0f;
        for_all_neighbors([&](int j, int neighbor_index_j, DiSurface& neighbor_surface) {
            float target_function_at_j = di_target_function_bias<V>(ctx, reservoir_being_resampled_sample, neighbor_surface);
            int M = 1;
            if (use_confidence_weights)
                M = di_load(rd, input_reservoir_buffer, neighbor_index_j).M;
            denom += target_function_at_j * M;
            if (j == current_neighbor_index)
                nume = target_function_at_j * M;
        });
        return denom == 0.0f ? 0.0f : nume / denom;
    };

    int reused_neighbors_count = local.reuse_neighbor_count;
    int start_index = 0;
    if (valid_neighbors_M_sum == 0)
        start_index = reused_neighbors_count;
    for (int neighbor_index = start_index; neighbor_index < reused_neighbors_count + 1; neighbor_index++)
    {
        if (neighbor_index < reused_neighbors_count && reused_neighbors_count <= 32)
        {
            if ((neighbor_heuristics_cache & (1 << neighbor_index)) == 0)
            {
                spatial_neighbor_advance_rng(rd, spatial_neighbors_rng);
                continue;
            }
        }
        int neighbor_pixel_index = get_spatial_neighbor_pixel_index(rd, local, neighbor_index, (int)x, (int)y, spatial_neighbors_rng);
        if (neighbor_pixel_index == -1)
            continue;
        if (neighbor_index < reused_neighbors_count && reused_neighbors_count > 32)
            if (!check_neighbor_similarity_heuristics(rd, neighbor_pixel_index, (int)center_pixel_index, center_pixel_surface.shading_point, normal_for_rejection_heuristic(rd, center_pixel_surface), false))
                continue;

        DiReservoir neighbor_reservoir = di_load(rd, input_reservoir_buffer, neighbor_pixel_index);
        float target_function_at_center = 0.0f;
        bool do_neighbor_target_function_visibility = do_include_spatial_visibility_term(ctx, local, neighbor_index);
        if (neighbor_reservoir.UCW > 0.0f)
        {
            if (neighbor_index == reused_neighbors_count)
                target_function_at_center = neighbor_reservoir.sample.target_function;
            else if (do_neighbor_target_function_visibility)
                target_function_at_center = di_target_function<V, true>(ctx, neighbor_reservoir.sample, center_pixel_surface);
            else
                target_function_at_center = di_target_function<V, false>(ctx, neighbor_reservoir.sample, center_pixel_surface);
        }

        /* ReSTIR_DI_BiasCorrectionWeights (SpatialMISWeight.h): a -D macro in the reference, a launch-time switch here */
        float mis_weight;
        if (bias == B2_BIAS_1_OVER_Z)
            mis_weight = (float)neighbor_reservoir.M;
        else if (bias == B2_BIAS_MIS_LIKE)
            mis_weight = use_confidence_weights ? (float)neighbor_reservoir.M : 1.0f;
        else if (bias == B2_BIAS_MIS_GBH)
            mis_weight = gbh_weight(neighbor_reservoir.UCW, neighbor_reservoir.sample, neighbor_index);
        else if (bias == B2_BIAS_SYMMETRIC_RATIO || bias == B2_BIAS_ASYMMETRIC_RATIO)
        {
            /* symmetric ratio (SpatialMISWeight.h:370-475) and asymmetric ratio (:477-599); mis.mc is the canonical sample's running weight */
            const bool asymmetric = bias == B2_BIAS_ASYMMETRIC_RATIO;
            const float beta = di.symmetric_ratio_mis_weights_beta_exponent;
            bool update_mc = center_pixel_reservoir.M > 0 && center_pixel_reservoir.UCW > 0.0f;
            if (neighbor_index == reused_neighbors_count)
                mis_weight = mis.mc == 0.0f ? 1.0f : (asymmetric ? 1.0f - mis.mc : mis.mc);
            else
            {
                float target_function_neighbor_sample_at_neighbor = neighbor_reservoir.sample.target_function;
                float target_function_center_sample_at_center = center_pixel_reservoir.sample.target_function;
                float reservoir_resampled_M = use_confidence_weights ? (float)neighbor_reservoir.M : 1.0f;
                float center_reservoir_M = use_confidence_weights ? (float)center_pixel_reservoir.M : 1.0f;
                float neighbors_confidence_sum = use_confidence_weights ? (float)valid_neighbors_M_sum : (float)valid_neighbors_count;

                float difference_function = ratio_mis_difference_function(target_function_at_center, target_function_neighbor_sample_at_neighbor, beta);
                float nume_mi = difference_function * reservoir_resampled_M;
                float denom_mi = center_reservoir_M + neighbors_confidence_sum * difference_function;
                if (asymmetric && !(target_function_at_center <= target_function_neighbor_sample_at_neighbor))
                    denom_mi = center_reservoir_M + neighbors_confidence_sum;
                mis_weight = nume_mi / denom_mi;

                if (update_mc)
                {
                    DiSurface neighbor_pixel_surface = get_pixel_surface<V>(rd, neighbor_pixel_index, false, rng);
                    float target_function_center_sample_at_neighbor = di_target_function_bias<V>(ctx, center_pixel_reservoir.sample, neighbor_pixel_surface);
                    float difference_function_mc = ratio_mis_difference_function(target_function_center_sample_at_neighbor, target_function_center_sample_at_center, beta);
                    if (asymmetric)
                    {
                        float nume_mc = difference_function_mc * reservoir_resampled_M;
                        float denom_mc = center_reservoir_M + neighbors_confidence_sum * difference_function_mc;
                        if (!(target_function_center_sample_at_center <= target_function_center_sample_at_neighbor))
                            denom_mc = center_reservoir_M + neighbors_confidence_sum;
                        mis.mc += nume_mc / denom_mc;
                    }
                    else
                    {
                        float nume_mc = center_reservoir_M;
                        float denom_mc = center_reservoir_M + neighbors_confidence_sum * difference_function_mc;
                        float confidence_weights_multiplier;
                        if (use_confidence_weights)
                            confidence_weights_multiplier = neighbors_confidence_sum == 0.0f ? 0.0f : reservoir_resampled_M / neighbors_confidence_sum;
                        else
                            confidence_weights_multiplier = 1.0f / (float)valid_neighbors_count;
                        mis.mc += confidence_weights_multiplier * nume_mc / denom_mc;
                    }
                }
            }
        }
        else
        {
            bool update_mc = center_pixel_reservoir.M > 0 && center_pixel_reservoir.UCW > 0.0f;
            mis_weight = mis.weight<V>(ctx, neighbor_reservoir.M, neighbor_reservoir.sample.target_function, center_pixel_reservoir.M, center_pixel_reservoir.sample.target_function,
                                       center_pixel_reservoir.sample, target_function_at_center, neighbor_pixel_index, false, valid_neighbors_count, valid_neighbors_M_sum, update_mc,
                                       neighbor_index == reused_neighbors_count, true, rng, bias == B2_BIAS_PAIRWISE_MIS_DEFENSIVE);
        }

        if (out.combine_with(neighbor_reservoir, mis_weight, target_function_at_center, 1.0f, rng))
        {
            selected_neighbor = neighbor_index;
            if (do_neighbor_target_function_visibility)
                out.sample.or_flags(B2_DI_FLAGS_UNOCCLUDED);
            else if (neighbor_index == reused_neighbors_count)
                out.sample.or_flags((int)(neighbor_reservoir.sample.flags & B2_DI_FLAGS_UNOCCLUDED));
            else
                out.sample.flags &= ~(uint32_t)B2_DI_FLAGS_UNOCCLUDED;
        }

        /* ReSTIR_optimal_visibility_sampling (OptimalVisibilitySampling.h:17-46, called at SpatialReuse.h:225): once the last neighbour
         * before the canonical sample has been resampled, what the neighbours produced is tested for visibility and thrown away if it is
         * occluded, so that the canonical sample is not drowned by it. (The reference does not compile with the option on -- :34 names a
         * member ReSTIRSurface does not have; the oracle builds it with -Dlast_hit_primitive_index=primitive_index, the member meant.) */
        if ((ctx.restir_options & B2_DI_OPT_OPTIMAL_VISIBILITY_SAMPLING) && out.weight_sum > 0.0f && neighbor_index == reused_neighbors_count - 1)
            if (di_visibility_test_kill_reservoir(ctx, out, center_pixel_surface.shading_point, center_pixel_surface.primitive_index))
                out.weight_sum = 0.0f;
    }

    /* SpatialNormalizationWeight.h: 1/Z divides by the confidence of the neighbours that could have produced the chosen sample (:64-131),
     * MIS-like weighs the chosen neighbour against all of them (:133-198); the others put everything into the resampling weights */
    float normalization_numerator = 1.0f, normalization_denominator = 1.0f;
    if ((bias == B2_BIAS_1_OVER_Z || bias == B2_BIAS_MIS_LIKE) && !(out.weight_sum <= 0.0f))
    {
        const DiSample final_reservoir_sample = out.sample;
        normalization_denominator = 0.0f;
        normalization_numerator = bias == B2_BIAS_1_OVER_Z ? 1.0f : 0.0f;
        for_all_neighbors([&](int j, int neighbor_index_j, DiSurface& neighbor_surface) {
            float target_function_at_neighbor = di_target_function_bias<V>(ctx, final_reservoir_sample, neighbor_surface);
            if (target_function_at_neighbor > 0.0f)
            {
                if (bias == B2_BIAS_1_OVER_Z)
                    normalization_denominator += (float)di_load(rd, input_reservoir_buffer, neighbor_index_j).M;
                else
                {
                    int M = 1;
                    if (use_confidence_weights)
                        M = di_load(rd, input_reservoir_buffer, neighbor_index_j).M;
                    if (j == selected_neighbor)
                        normalization_numerator = target_function_at_neighbor;
                    normalization_denominator += target_function_at_neighbor * M;
                }
            }
        });
    }
    out.end_with_normalization(normalization_numerator, normalization_denominator);

    /* the visibility-reuse test at the end of the pass exists for 1/Z, pairwise MIS and its defensive form only (SpatialReuse.h:264-275) */
    bool not_last_spatial_pass = di.common_spatial_pass.number_of_passes - 1 != di.common_spatial_pass.spatial_pass_index;
    if (bias != B2_BIAS_MIS_LIKE && bias != B2_BIAS_MIS_GBH && di_end_of_pass_visibility_reuse(ctx) && (di.common_temporal_pass.do_temporal_reuse_pass || not_last_spatial_pass))
        di_visibility_test_kill_reservoir(ctx, out, center_pixel_surface.shading_point, center_pixel_surface.primitive_index);

    if (di.m_cap > 0)
        out.M = b2_min(out.M, di.m_cap);
    di_store(rd, di.spatial_pass.output_reservoirs, (int)center_pixel_index, out);
}

/* ---------------------------------------------------------------------------------------------------------- */
/* fused spatiotemporal reuse (FusedSpatiotemporalReuse.h) */
/* ---- the fused spatiotemporal pass cut at its rays ------------------------------------------------------------------------------
 * Up to 8 shadow rays per pixel decide target functions that are otherwise fully determined: the temporal neighbour's sample seen
 * from the center pixel, and the center pixel's canonical sample seen from the temporal and from every spatial neighbour (the
 * pairwise-MIS canonical weight, ReSTIR_DI_BiasCorrectionUseVisibility). They do not depend on one another, so:
 *     phase A: neighbours found, every target function evaluated AS IF VISIBLE              -> up to 8 shadow rays
 *     phase B: occluded values zeroed, MIS weights, the reservoirs combined (RNG draws in the reference's order),
 *              normalisation                                                                -> the visibility-reuse shadow ray
 *     phase C: visibility reuse, M cap, store
 * restir_di_spatiotemporal_pixel below = A, trace, B, trace, C for one pixel; the wavefront kernels run each phase frame-wide. */
#define B2_ST_MAX_ENTRIES 8 /* temporal neighbour + up to 6 spatial neighbours + the center pixel; more neighbours: the caller uses the in-place form */
enum { B2_ST_TEMPORAL = 1, B2_ST_CANONICAL = 2 };
struct DiStEntry
{
    int neighbor_pixel_index;
    float target_function_at_center;            /* as if visible */
    float target_function_canonical_at_neighbor; /* as if visible */
    int kind;
    float combine_draw; /* the random number of the entry's combine_with, drawn in phase A at the reference's stream position */
};
struct DiSpatiotemporalState
{
    Rng rng;
    int active, pending_final;
    int entry_count;
    int valid_neighbors_count, valid_neighbors_M_sum;
    int update_mc;
    DiStEntry entries[B2_ST_MAX_ENTRIES];
    DiReservoir out; /* between phases B and C */
};
/* ray slots of phase A: [2 k] = center-side ray of entry k (temporal entry only), [2 k + 1] = its canonical-at-neighbour ray */
#define B2_ST_RAYS (2 * B2_ST_MAX_ENTRIES)

/* di_target_function<V, true> split at its shadow ray */
template <class V>
B2_HD float di_target_function_deferred(const Ctx& ctx, const DiSample& sample, DiSurface& surface, RayRequest& ray)
{
    request_none(ray);
    const b2pt_render_data& rd = *ctx.rd;
    if (sample.emissive_triangle_index == -1 && !sample.is_envmap_sample())
        return 0.0f;
    float distance_to_light;
    V3 sample_direction = di_light_sample_direction(rd, sample, surface.shading_point, distance_to_light);
    float target_function = di_target_function<V, false>(ctx, sample, surface);
    if (target_function == 0.0f)
        return 0.0f;
    request_shadow(ray, surface.shading_point, sample_direction, distance_to_light, surface.primitive_index);
    return target_function;
}

template <class V>
B2_HD bool restir_di_spatiotemporal_supported(const b2pt_render_data& rd)
{
    const b2pt_restir_di_settings& di = rd.render_settings.restir_di_settings;
    int most_neighbors = b2_max(di.common_spatial_pass.reuse_neighbor_count, di.common_spatial_pass.do_disocclusion_reuse_boost ? di.common_spatial_pass.disocclusion_reuse_count : 0);
    return most_neighbors + 2 <= B2_ST_MAX_ENTRIES;
}

template <class V>
B2_HD void restir_di_spatiotemporal_phase_a(const Ctx& ctx, uint32_t x, uint32_t y, DiSpatiotemporalState& st, RayRequest (&rays)[B2_ST_RAYS])
{
    const b2pt_render_data& rd = *ctx.rd;
    const b2pt_restir_di_settings& di = rd.render_settings.restir_di_settings;
    for (int k = 0; k < B2_ST_RAYS; k++)
        request_none(rays[k]);
    st.active = 0;
    st.pending_final = 0;
    st.entry_count = 0;
    uint32_t center_pixel_index;
    if (!restir_pass_prologue<V>(rd, x, y, center_pixel_index))
        return;
    if (!rd.aux_buffers.pixel_active[center_pixel_index] || rd.g_buffer.first_hit_prim_index[center_pixel_index] == -1)
        return;
    st.active = 1;

    unsigned int seed = rd.render_settings.freeze_random ? wang_hash(center_pixel_index + 1) : wang_hash(((center_pixel_index + 1) * (rd.render_settings.sample_number + 1)) ^ rd.random_number);
    Rng rng = make_rng(seed);

    SpatialLocal local = spatial_local_from(rd);
    local.rng_seed = rng.next_u32();

    if (di.common_temporal_pass.temporal_buffer_clear_requested)
    {
        DiReservoir empty;
        empty.init();
        di_store(rd, di.temporal_pass.input_reservoirs, (int)center_pixel_index, empty);
    }

    DiSurface center_pixel_surface = get_pixel_surface<V>(rd, (int)center_pixel_index, false, rng);
    const bool prev_gbuffer = use_prev_frame_g_buffer<V>(rd.render_settings);

    /* load_spatiotemporal_neighbor_data (:73-104) */
    DiReservoir temporal_neighbor_reservoir;
    temporal_neighbor_reservoir.init();
    Int3 temporal = find_temporal_neighbor_index<V>(rd, load_v3(rd.g_buffer.primary_hit_position, center_pixel_index), center_pixel_surface.shading_normal, (int)center_pixel_index, rng);
    if (!(temporal.x == -1 || rd.render_settings.freeze_random))
        temporal_neighbor_reservoir = di_load(rd, di.temporal_pass.input_reservoirs, temporal.x);
    /* load_spatiotemporal_neighbor_data rebuilds the temporal neighbour's surface here (:97), which draws a wavelength when that pixel shows
     * dispersive glass. Pairwise MIS never reads that surface (its weight function rebuilds its own, SpatiotemporalMISWeight.h:183): only
     * the draw is kept. */
    if (!(temporal.x == -1 || rd.render_settings.freeze_random) && temporal_neighbor_reservoir.M != 0)
        (void)get_pixel_surface<V>(rd, temporal.x, prev_gbuffer, rng);
    if ((temporal.x == -1 || temporal_neighbor_reservoir.M <= 1) && di.common_spatial_pass.do_disocclusion_reuse_boost)
        local.reuse_neighbor_count = di.common_spatial_pass.disocclusion_reuse_count;

    setup_adaptive_directional_spatial_reuse(rd, local, center_pixel_index, rng);

    int neighbor_heuristics_cache = 0;
    int valid_neighbors_count = 0;
    int valid_neighbors_M_sum = 0;
    count_valid_neighbors(rd, local, center_pixel_surface, (int)center_pixel_index, temporal.y, temporal.z, prev_gbuffer, di.spatial_pass.input_reservoirs, valid_neighbors_count,
                          valid_neighbors_M_sum, neighbor_heuristics_cache);
    if (temporal.x != -1 && temporal_neighbor_reservoir.M > 0)
    {
        valid_neighbors_count++;
        valid_neighbors_M_sum += temporal_neighbor_reservoir.M;
    }
    st.valid_neighbors_count = valid_neighbors_count;
    st.valid_neighbors_M_sum = valid_neighbors_M_sum;

    DiReservoir initial_candidates_reservoir = di_load(rd, di.initial_candidates.output_reservoirs, (int)center_pixel_index);
    const bool update_mc = initial_candidates_reservoir.M > 0 && initial_candidates_reservoir.UCW > 0.0f;
    st.update_mc = update_mc ? 1 : 0;

    /* the canonical sample seen from a neighbour's surface (pairwise MIS, update_mc; SpatiotemporalMISWeight.h:181-191), as if visible */
    auto canonical_at = [&](int neighbor_pixel_index, RayRequest& ray) {
        if (!update_mc)
            return 0.0f;
        DiSurface neighbor_pixel_surface = get_pixel_surface<V>(rd, neighbor_pixel_index, prev_gbuffer, rng);
        return di_target_function_deferred<V>(ctx, initial_candidates_reservoir.sample, neighbor_pixel_surface, ray);
    };

    if (temporal.x != -1 && temporal_neighbor_reservoir.M > 0)
    {
        DiStEntry& e = st.entries[st.entry_count];
        e.neighbor_pixel_index = temporal.x;
        e.kind = B2_ST_TEMPORAL;
        e.target_function_at_center = 0.0f;
        if (temporal_neighbor_reservoir.UCW > 0.0f)
            e.target_function_at_center = di_target_function_deferred<V>(ctx, temporal_neighbor_reservoir.sample, center_pixel_surface, rays[2 * st.entry_count]);
        e.target_function_canonical_at_neighbor = canonical_at(temporal.x, rays[2 * st.entry_count + 1]);
        /* :281 combine_with draws only for a reservoir with UCW > 0 */
        e.combine_draw = temporal_neighbor_reservoir.UCW > 0.0f ? rng.next_float() : 0.0f;
        st.entry_count++;
    }

    const void* spatial_input_reservoir_buffer = di.spatial_pass.input_reservoirs;
    Rng spatial_neighbors_rng = make_rng(local.rng_seed);
    int reused_neighbors_count = local.reuse_neighbor_count;
    int start_index = 0;
    if (valid_neighbors_M_sum == 0)
        start_index = reused_neighbors_count;
    for (int spatial_neighbor_index = start_index; spatial_neighbor_index < reused_neighbors_count + 1; spatial_neighbor_index++)
    {
        if (spatial_neighbor_index < reused_neighbors_count && reused_neighbors_count <= 32)
            if ((neighbor_heuristics_cache & (1 << spatial_neighbor_index)) == 0)
            {
                spatial_neighbor_advance_rng(rd, spatial_neighbors_rng);
                continue;
            }

        int neighbor_pixel_index = -1;
        if (spatial_neighbor_index == reused_neighbors_count)
            neighbor_pixel_index = (int)center_pixel_index;
        else
            neighbor_pixel_index = get_spatial_neighbor_pixel_index(rd, local, spatial_neighbor_index, temporal.y, temporal.z, spatial_neighbors_rng);
        if (neighbor_pixel_index == -1)
            continue;

        if (spatial_neighbor_index < reused_neighbors_count && reused_neighbors_count > 32)
            if (!check_neighbor_similarity_heuristics(rd, neighbor_pixel_index, (int)center_pixel_index, center_pixel_surface.shading_point, normal_for_rejection_heuristic(rd, center_pixel_surface), prev_gbuffer))
                continue;

        const bool canonical = spatial_neighbor_index == reused_neighbors_count;
        DiReservoir neighbor_reservoir = canonical ? initial_candidates_reservoir : di_load(rd, spatial_input_reservoir_buffer, neighbor_pixel_index);

        DiStEntry& e = st.entries[st.entry_count];
        e.neighbor_pixel_index = neighbor_pixel_index;
        e.kind = canonical ? B2_ST_CANONICAL : 0;
        e.target_function_at_center = 0.0f;
        if (neighbor_reservoir.UCW > 0.0f)
        {
            if (canonical)
                e.target_function_at_center = neighbor_reservoir.sample.target_function;
            else
                e.target_function_at_center = di_target_function<V, false>(ctx, neighbor_reservoir.sample, center_pixel_surface);
        }
        /* :397 the neighbour's surface is rebuilt for every entry, the centre pixel's own included (a wavelength draw on dispersive glass;
         * pairwise MIS does not read it), then the weight function rebuilds it once more for the canonical sample, then :444 combine_with draws */
        (void)get_pixel_surface<V>(rd, neighbor_pixel_index, prev_gbuffer, rng);
        e.target_function_canonical_at_neighbor = canonical ? 0.0f : canonical_at(neighbor_pixel_index, rays[2 * st.entry_count + 1]);
        e.combine_draw = neighbor_reservoir.UCW > 0.0f ? rng.next_float() : 0.0f;
        st.entry_count++;
    }
    st.rng = rng;
}

template <class V>
B2_HD void restir_di_spatiotemporal_phase_b(const Ctx& ctx, uint32_t x, uint32_t y, DiSpatiotemporalState& st, const B2Hit (&results)[B2_ST_RAYS], RayRequest& visibility_ray)
{
    const b2pt_render_data& rd = *ctx.rd;
    const b2pt_restir_di_settings& di = rd.render_settings.restir_di_settings;
    request_none(visibility_ray);
    if (!st.active)
        return;
    const int center_pixel_index = (int)(x + y * (uint32_t)rd.render_settings.render_resolution.x);
    Rng rng = st.rng;
    DiReservoir initial_candidates_reservoir = di_load(rd, di.initial_candidates.output_reservoirs, center_pixel_index);
    const bool update_mc = st.update_mc != 0;

    DiReservoir out;
    out.init();
    PairwiseMis mis;
    mis.mc = 0.0f;
    for (int k = 0; k < st.entry_count; k++)
    {
        const DiStEntry& e = st.entries[k];
        /* an occluded ray zeroes its target function (target_function *= visible ? 1 : 0); unrequested rays come back as misses */
        float target_function_at_center = e.target_function_at_center;
        if (results[2 * k].prim >= 0)
            target_function_at_center *= 0.0f;
        float target_function_canonical_at_neighbor = e.target_function_canonical_at_neighbor;
        if (results[2 * k + 1].prim >= 0)
            target_function_canonical_at_neighbor *= 0.0f;

        if (e.kind == B2_ST_TEMPORAL)
        {
            DiReservoir temporal_neighbor_reservoir = di_load(rd, di.temporal_pass.input_reservoirs, e.neighbor_pixel_index);
            float w = mis.weight_given(rd, temporal_neighbor_reservoir.M, temporal_neighbor_reservoir.sample.target_function, initial_candidates_reservoir.M,
                                       initial_candidates_reservoir.sample.target_function, target_function_at_center, target_function_canonical_at_neighbor, st.valid_neighbors_count,
                                       st.valid_neighbors_M_sum, update_mc, false, false);
            if (out.combine_with_drawn(temporal_neighbor_reservoir, w, target_function_at_center, 1.0f, e.combine_draw))
                out.sample.or_flags(B2_DI_FLAGS_UNOCCLUDED);
        }
        else
        {
            const bool canonical = e.kind == B2_ST_CANONICAL;
            DiReservoir neighbor_reservoir = canonical ? initial_candidates_reservoir : di_load(rd, di.spatial_pass.input_reservoirs, e.neighbor_pixel_index);
            float mis_weight = mis.weight_given(rd, neighbor_reservoir.M, neighbor_reservoir.sample.target_function, initial_candidates_reservoir.M,
                                                initial_candidates_reservoir.sample.target_function, target_function_at_center, target_function_canonical_at_neighbor,
                                                st.valid_neighbors_count, st.valid_neighbors_M_sum, update_mc, canonical, false);
            if (out.combine_with_drawn(neighbor_reservoir, mis_weight, target_function_at_center, 1.0f, e.combine_draw))
            {
                /* ReSTIR_DI_SpatialTargetFunctionVisibility = FALSE: do_include_spatial_visibility_term is false for every neighbour */
                if (canonical)
                    out.sample.or_flags((int)(neighbor_reservoir.sample.flags & B2_DI_FLAGS_UNOCCLUDED));
                else
                    out.sample.flags &= ~(uint32_t)B2_DI_FLAGS_UNOCCLUDED;
            }
        }
    }

    out.end_with_normalization(1.0f, 1.0f);

    st.out = out;
    st.rng = rng;
    if (di.common_temporal_pass.do_temporal_reuse_pass || di.common_spatial_pass.number_of_passes - 1 != di.common_spatial_pass.spatial_pass_index)
        st.pending_final = di_visibility_test_request(rd, out, load_v3(rd.g_buffer.primary_hit_position, center_pixel_index), rd.g_buffer.first_hit_prim_index[center_pixel_index],
                                                      visibility_ray)
                               ? 1
                               : 0;
}

template <class V>
B2_HD void restir_di_spatiotemporal_phase_c(const Ctx& ctx, uint32_t x, uint32_t y, DiSpatiotemporalState& st, const B2Hit& visibility_hit)
{
    const b2pt_render_data& rd = *ctx.rd;
    const b2pt_restir_di_settings& di = rd.render_settings.restir_di_settings;
    if (!st.active)
        return;
    const int center_pixel_index = (int)(x + y * (uint32_t)rd.render_settings.render_resolution.x);
    DiReservoir out = st.out;
    if (st.pending_final)
        di_visibility_test_apply(out, visibility_hit.prim >= 0);
    if (di.m_cap > 0)
        out.M = b2_min(out.M, di.m_cap);
    di_store(rd, di.spatial_pass.output_reservoirs, center_pixel_index, out);
}

/* ---- the fused pass in one piece -------------------------------------------------------------------------------------------------
 * FusedSpatiotemporalReuse.h:169-563 followed top to bottom, rays traced where they occur. Takes any number of neighbours and every
 * ReSTIR_DI_BiasCorrectionWeights but 1/M; the phase-cut form above is the fast path of the default (pairwise MIS, <= 6 neighbours).
 * What the reference does here and is mirrored as written:
 *   - the weight / normalisation functions that walk over all neighbours again (generalised balance heuristic, 1/Z, MIS-like:
 *     SpatiotemporalMISWeight.h:78-121, SpatiotemporalNormalizationWeight.h:98-131,170-215) draw the neighbour positions from the
 *     kernel's main generator where it stands, not from the spatial-neighbours seed: they visit OTHER pixels than the ones resampled;
 *   - temporal_neighbor_surface stays default-constructed (zero normals: every target function on it is 0) when there is no temporal
 *     neighbour or its reservoir is empty (:73-104);
 *   - the asymmetric-ratio weights of this pass are a copy of the symmetric ones (SpatiotemporalMISWeight.h:449-555), and both rebuild
 *     the neighbour's surface from the CURRENT G-buffer (:389) where pairwise MIS takes the previous frame's when asked to (:183). */
template <class V>
B2_HD void restir_di_spatiotemporal_inplace_pixel(const Ctx& ctx, uint32_t x, uint32_t y)
{
    const b2pt_render_data& rd = *ctx.rd;
    const b2pt_restir_di_settings& di = rd.render_settings.restir_di_settings;
    uint32_t center_pixel_index;
    if (!restir_pass_prologue<V>(rd, x, y, center_pixel_index))
        return;
    if (!rd.aux_buffers.pixel_active[center_pixel_index] || rd.g_buffer.first_hit_prim_index[center_pixel_index] == -1)
        return;

    unsigned int seed = rd.render_settings.freeze_random ? wang_hash(center_pixel_index + 1) : wang_hash(((center_pixel_index + 1) * (rd.render_settings.sample_number + 1)) ^ rd.random_number);
    Rng rng = make_rng(seed);
    SpatialLocal local = spatial_local_from(rd);
    local.rng_seed = rng.next_u32();

    if (di.common_temporal_pass.temporal_buffer_clear_requested)
    {
        DiReservoir empty;
        empty.init();
        di_store(rd, di.temporal_pass.input_reservoirs, (int)center_pixel_index, empty);
    }

    DiSurface center_pixel_surface = get_pixel_surface<V>(rd, (int)center_pixel_index, false, rng);
    const bool prev_gbuffer = use_prev_frame_g_buffer<V>(rd.render_settings);
    const int bias = ctx.bias_correction;
    const bool use_confidence_weights = di.use_confidence_weights != 0;

    /* load_spatiotemporal_neighbor_data (:73-104) */
    DiReservoir temporal_neighbor_reservoir;
    temporal_neighbor_reservoir.init();
    DiSurface temporal_neighbor_surface;
    temporal_neighbor_surface.material = unpack_material(PackedMat{});
    temporal_neighbor_surface.ray_volume_state.init();
    temporal_neighbor_surface.primitive_index = -1;
    temporal_neighbor_surface.view_direction = temporal_neighbor_surface.shading_normal = temporal_neighbor_surface.geometric_normal = temporal_neighbor_surface.shading_point = v3(0.0f, 0.0f, 0.0f);
    Int3 temporal = find_temporal_neighbor_index<V>(rd, load_v3(rd.g_buffer.primary_hit_position, center_pixel_index), center_pixel_surface.shading_normal, (int)center_pixel_index, rng);
    if (!(temporal.x == -1 || rd.render_settings.freeze_random))
    {
        temporal_neighbor_reservoir = di_load(rd, di.temporal_pass.input_reservoirs, temporal.x);
        if (temporal_neighbor_reservoir.M != 0)
            temporal_neighbor_surface = get_pixel_surface<V>(rd, temporal.x, prev_gbuffer, rng);
    }
    if ((temporal.x == -1 || temporal_neighbor_reservoir.M <= 1) && di.common_spatial_pass.do_disocclusion_reuse_boost)
        local.reuse_neighbor_count = di.common_spatial_pass.disocclusion_reuse_count;

    setup_adaptive_directional_spatial_reuse(rd, local, center_pixel_index, rng);

    int selected_neighbor = 0;
    int neighbor_heuristics_cache = 0;
    int valid_neighbors_count = 0;
    int valid_neighbors_M_sum = 0;
    count_valid_neighbors(rd, local, center_pixel_surface, (int)center_pixel_index, temporal.y, temporal.z, prev_gbuffer, di.spatial_pass.input_reservoirs, valid_neighbors_count,
                          valid_neighbors_M_sum, neighbor_heuristics_cache);
    if (temporal.x != -1 && temporal_neighbor_reservoir.M > 0)
    {
        valid_neighbors_count++;
        valid_neighbors_M_sum += temporal_neighbor_reservoir.M;
    }

    DiReservoir out;
    out.init();
    DiReservoir initial_candidates_reservoir = di_load(rd, di.initial_candidates.output_reservoirs, (int)center_pixel_index);
    const bool update_mc = initial_candidates_reservoir.M > 0 && initial_candidates_reservoir.UCW > 0.0f;
    PairwiseMis mis;
    mis.mc = 0.0f;

    /* the "all neighbours again" walk of the balance heuristic and of the 1/Z / MIS-like normalisations; `evaluate(j, M of neighbour j,
     * its surface)` for every neighbour that is inside the viewport and similar enough, the centre pixel (j = neighbour count) last */
    auto for_all_spatial_neighbors = [&](bool surfaces_from_previous_frame, auto&& evaluate) {
        for (int j = 0; j < local.reuse_neighbor_count + 1; j++)
        {
            int neighbor_index_j = (int)center_pixel_index;
            if (j != local.reuse_neighbor_count)
            {
                neighbor_index_j = get_spatial_neighbor_pixel_index(rd, local, j, temporal.y, temporal.z, rng);
                if (neighbor_index_j == -1)
                    continue;
                if (!check_neighbor_similarity_heuristics(rd, neighbor_index_j, (int)center_pixel_index, center_pixel_surface.shading_point, normal_for_rejection_heuristic(rd, center_pixel_surface), prev_gbuffer))
                    continue;
            }
            DiSurface neighbor_surface = center_pixel_surface;
            if (j != local.reuse_neighbor_count)
                neighbor_surface = get_pixel_surface<V>(rd, neighbor_index_j, surfaces_from_previous_frame, rng);
            int M = j == local.reuse_neighbor_count ? initial_candidates_reservoir.M : di_load(rd, di.spatial_pass.input_reservoirs, neighbor_index_j).M;
            evaluate(j, M, neighbor_surface);
        }
    };
    /* SpatiotemporalMISWeight.h:59-137; current_neighbor: 0 = the temporal neighbour, j + 1 = spatial neighbour j */
    auto gbh_weight = [&](float reservoir_being_resampled_UCW, const DiSample& reservoir_being_resampled_sample, int current_neighbor) {
        if (reservoir_being_resampled_UCW <= 0.0f)
            return 1.0f;
        float nume = 0.0f, denom = 0.0f;
        for_all_spatial_neighbors(prev_gbuffer, [&](int j, int M_j, DiSurface& neighbor_surface) {
            float target_function_at_j = di_target_function_bias<V>(ctx, reservoir_being_resampled_sample, neighbor_surface);
            int M = use_confidence_weights ? M_j : 1;
            denom += target_function_at_j * M;
            if (j + 1 == current_neighbor)
                nume = target_function_at_j * M;
        });
        float target_function_at_temporal_neighbor = di_target_function_bias<V>(ctx, reservoir_being_resampled_sample, temporal_neighbor_surface);
        int M = use_confidence_weights ? temporal_neighbor_reservoir.M : 1;
        denom += target_function_at_temporal_neighbor * M;
        if (current_neighbor == 0)
            nume = target_function_at_temporal_neighbor * M;
        return denom == 0.0f ? 0.0f : nume / denom;
    };
    /* one resampling weight: `resampled` seen from the centre pixel has target_function_at_center */
    auto resampling_weight = [&](const DiReservoir& resampled, float target_function_at_center, int neighbor_pixel_index, int current_neighbor, bool canonical) {
        if (bias == B2_BIAS_1_OVER_Z || bias == B2_BIAS_1_OVER_M)
            return (float)resampled.M;
        if (bias == B2_BIAS_MIS_LIKE)
            return use_confidence_weights ? (float)resampled.M : 1.0f;
        if (bias == B2_BIAS_MIS_GBH)
            return gbh_weight(resampled.UCW, resampled.sample, current_neighbor);
        if (bias == B2_BIAS_SYMMETRIC_RATIO || bias == B2_BIAS_ASYMMETRIC_RATIO)
        {
            /* SpatiotemporalMISWeight.h:341-447 (= :449-555) */
            if (canonical)
                return mis.mc == 0.0f ? 1.0f : mis.mc;
            const float beta = di.symmetric_ratio_mis_weights_beta_exponent;
            float target_function_neighbor_sample_at_neighbor = resampled.sample.target_function;
            float target_function_center_sample_at_center = initial_candidates_reservoir.sample.target_function;
            float reservoir_resampled_M = use_confidence_weights ? (float)resampled.M : 1.0f;
            float center_reservoir_M = use_confidence_weights ? (float)initial_candidates_reservoir.M : 1.0f;
            float neighbors_confidence_sum = use_confidence_weights ? (float)valid_neighbors_M_sum : (float)valid_neighbors_count;
            float difference_function = ratio_mis_difference_function(target_function_at_center, target_function_neighbor_sample_at_neighbor, beta);
            float nume_mi = difference_function * reservoir_resampled_M;
            float denom_mi = center_reservoir_M + neighbors_confidence_sum * difference_function;
            float mi = nume_mi / denom_mi;
            if (update_mc)
            {
                DiSurface neighbor_pixel_surface = get_pixel_surface<V>(rd, neighbor_pixel_index, false, rng);
                float target_function_center_sample_at_neighbor = di_target_function_bias<V>(ctx, initial_candidates_reservoir.sample, neighbor_pixel_surface);
                float nume_mc = center_reservoir_M;
                float denom_mc = center_reservoir_M + neighbors_confidence_sum * ratio_mis_difference_function(target_function_center_sample_at_neighbor, target_function_center_sample_at_center, beta);
                float confidence_weights_multiplier;
                if (use_confidence_weights)
                    confidence_weights_multiplier = neighbors_confidence_sum == 0.0f ? 0.0f : reservoir_resampled_M / neighbors_confidence_sum;
                else
                    confidence_weights_multiplier = 1.0f / (float)valid_neighbors_count;
                mis.mc += confidence_weights_multiplier * nume_mc / denom_mc;
            }
            return mi;
        }
        return mis.weight<V>(ctx, resampled.M, resampled.sample.target_function, initial_candidates_reservoir.M, initial_candidates_reservoir.sample.target_function,
                             initial_candidates_reservoir.sample, target_function_at_center, neighbor_pixel_index, prev_gbuffer, valid_neighbors_count, valid_neighbors_M_sum, update_mc,
                             canonical, false, rng, bias == B2_BIAS_PAIRWISE_MIS_DEFENSIVE);
    };

    if (temporal.x != -1 && temporal_neighbor_reservoir.M > 0)
    {
        float target_function_at_center = 0.0f;
        if (temporal_neighbor_reservoir.UCW > 0.0f)
            target_function_at_center = di_target_function_bias<V>(ctx, temporal_neighbor_reservoir.sample, center_pixel_surface);
        float w = resampling_weight(temporal_neighbor_reservoir, target_function_at_center, temporal.x, 0, false);
        if (out.combine_with(temporal_neighbor_reservoir, w, target_function_at_center, 1.0f, rng))
        {
            selected_neighbor = 0;
            if (ctx.restir_options & B2_DI_OPT_BIAS_VISIBILITY) /* :284-296 */
                out.sample.or_flags(B2_DI_FLAGS_UNOCCLUDED);
            else
                out.sample.flags &= ~(uint32_t)B2_DI_FLAGS_UNOCCLUDED;
        }
    }

    Rng spatial_neighbors_rng = make_rng(local.rng_seed);
    int reused_neighbors_count = local.reuse_neighbor_count;
    int start_index = 0;
    if (valid_neighbors_M_sum == 0)
        start_index = reused_neighbors_count;
    for (int spatial_neighbor_index = start_index; spatial_neighbor_index < reused_neighbors_count + 1; spatial_neighbor_index++)
    {
        if (spatial_neighbor_index < reused_neighbors_count && reused_neighbors_count <= 32)
            if ((neighbor_heuristics_cache & (1 << spatial_neighbor_index)) == 0)
            {
                spatial_neighbor_advance_rng(rd, spatial_neighbors_rng);
                continue;
            }

        int neighbor_pixel_index = -1;
        if (spatial_neighbor_index == reused_neighbors_count)
            neighbor_pixel_index = (int)center_pixel_index;
        else
            neighbor_pixel_index = get_spatial_neighbor_pixel_index(rd, local, spatial_neighbor_index, temporal.y, temporal.z, spatial_neighbors_rng);
        if (neighbor_pixel_index == -1)
            continue;

        if (spatial_neighbor_index < reused_neighbors_count && reused_neighbors_count > 32)
            if (!check_neighbor_similarity_heuristics(rd, neighbor_pixel_index, (int)center_pixel_index, center_pixel_surface.shading_point, normal_for_rejection_heuristic(rd, center_pixel_surface), prev_gbuffer))
                continue;

        const bool canonical = spatial_neighbor_index == reused_neighbors_count;
        DiReservoir neighbor_reservoir = canonical ? initial_candidates_reservoir : di_load(rd, di.spatial_pass.input_reservoirs, neighbor_pixel_index);

        float target_function_at_center = 0.0f;
        const bool do_neighbor_target_function_visibility = do_include_spatial_visibility_term(ctx, local, spatial_neighbor_index);
        if (neighbor_reservoir.UCW > 0.0f)
        {
            if (canonical)
                target_function_at_center = neighbor_reservoir.sample.target_function;
            else if (do_neighbor_target_function_visibility)
                target_function_at_center = di_target_function<V, true>(ctx, neighbor_reservoir.sample, center_pixel_surface);
            else
                target_function_at_center = di_target_function<V, false>(ctx, neighbor_reservoir.sample, center_pixel_surface);
        }
        (void)get_pixel_surface<V>(rd, neighbor_pixel_index, prev_gbuffer, rng); /* :397, read by ReSTIR GI only; its wavelength draw stays */

        float mis_weight = resampling_weight(neighbor_reservoir, target_function_at_center, neighbor_pixel_index, spatial_neighbor_index + 1, canonical);
        if (out.combine_with(neighbor_reservoir, mis_weight, target_function_at_center, 1.0f, rng))
        {
            selected_neighbor = spatial_neighbor_index + 1;
            if (do_neighbor_target_function_visibility)
                out.sample.or_flags(B2_DI_FLAGS_UNOCCLUDED);
            else if (canonical)
                out.sample.or_flags((int)(neighbor_reservoir.sample.flags & B2_DI_FLAGS_UNOCCLUDED));
            else
                out.sample.flags &= ~(uint32_t)B2_DI_FLAGS_UNOCCLUDED;
        }
    }

    /* SpatiotemporalNormalizationWeight.h: 1/Z (:84-147), MIS-like (:149-229); MIS-like rebuilds the surfaces from the current G-buffer (:189) */
    float normalization_numerator = 1.0f, normalization_denominator = 1.0f;
    if (bias == B2_BIAS_1_OVER_M && !(out.weight_sum <= 0.0f))
    {
        /* SpatiotemporalNormalizationWeight.h:19-77: the confidence of every neighbour that passes the heuristics, whatever it could have produced */
        normalization_denominator = 0.0f;
        for_all_spatial_neighbors(prev_gbuffer, [&](int, int M_j, DiSurface&) { normalization_denominator += (float)M_j; });
        normalization_denominator += (float)temporal_neighbor_reservoir.M;
    }
    if ((bias == B2_BIAS_1_OVER_Z || bias == B2_BIAS_MIS_LIKE) && !(out.weight_sum <= 0.0f))
    {
        const DiSample final_reservoir_sample = out.sample;
        normalization_denominator = 0.0f;
        normalization_numerator = bias == B2_BIAS_1_OVER_Z ? 1.0f : 0.0f;
        for_all_spatial_neighbors(bias == B2_BIAS_1_OVER_Z ? prev_gbuffer : false, [&](int j, int M_j, DiSurface& neighbor_surface) {
            float target_function_at_neighbor = di_target_function_bias<V>(ctx, final_reservoir_sample, neighbor_surface);
            if (target_function_at_neighbor > 0.0f)
            {
                if (bias == B2_BIAS_1_OVER_Z)
                    normalization_denominator += (float)M_j;
                else
                {
                    int M = use_confidence_weights ? M_j : 1;
                    if (j + 1 == selected_neighbor)
                        normalization_numerator += target_function_at_neighbor;
                    normalization_denominator += target_function_at_neighbor * M;
                }
            }
        });
        float target_function_at_temporal_neighbor = di_target_function_bias<V>(ctx, final_reservoir_sample, temporal_neighbor_surface);
        if (bias == B2_BIAS_1_OVER_Z)
        {
            if (target_function_at_temporal_neighbor > 0.0f)
                normalization_denominator += (float)temporal_neighbor_reservoir.M;
        }
        else
        {
            if (selected_neighbor == 0)
                normalization_numerator += target_function_at_temporal_neighbor;
            int temporal_M = use_confidence_weights ? temporal_neighbor_reservoir.M : 1;
            normalization_denominator += target_function_at_temporal_neighbor * temporal_M;
        }
    }
    out.end_with_normalization(normalization_numerator, normalization_denominator);

    if (bias != B2_BIAS_1_OVER_M && bias != B2_BIAS_MIS_LIKE && bias != B2_BIAS_MIS_GBH && di_end_of_pass_visibility_reuse(ctx) &&
        (di.common_temporal_pass.do_temporal_reuse_pass || di.common_spatial_pass.number_of_passes - 1 != di.common_spatial_pass.spatial_pass_index))
        di_visibility_test_kill_reservoir(ctx, out, center_pixel_surface.shading_point, center_pixel_surface.primitive_index);

    if (di.m_cap > 0)
        out.M = b2_min(out.M, di.m_cap);
    di_store(rd, di.spatial_pass.output_reservoirs, (int)center_pixel_index, out);
}

template <class V>
B2_HD void restir_di_spatiotemporal_pixel(const Ctx& ctx, uint32_t x, uint32_t y)
{
    if (ctx.bias_correction != B2_BIAS_PAIRWISE_MIS || (ctx.restir_options & B2_DI_OPT_PHASE_CUT_MASK) != B2_DI_OPT_DEFAULTS || !restir_di_spatiotemporal_supported<V>(*ctx.rd))
    {
        restir_di_spatiotemporal_inplace_pixel<V>(ctx, x, y);
        return;
    }
    DiSpatiotemporalState st;
    RayRequest rays[B2_ST_RAYS];
    B2Hit results[B2_ST_RAYS];
    restir_di_spatiotemporal_phase_a<V>(ctx, x, y, st, rays);
    for (int k = 0; k < B2_ST_RAYS; k++)
        results[k] = di_trace_request(ctx, rays[k]);
    RayRequest visibility_ray;
    restir_di_spatiotemporal_phase_b<V>(ctx, x, y, st, results, visibility_ray);
    B2Hit visibility_hit = di_trace_request(ctx, visibility_ray);
    restir_di_spatiotemporal_phase_c<V>(ctx, x, y, st, visibility_hit);
}

template <class V>
B2_HD void restir_di_pass_pixel(int pass, const Ctx& ctx, uint32_t x, uint32_t y)
{
    switch (pass)
    {
    case B2_PASS_RESTIR_DI_INITIAL_CANDIDATES: restir_di_initial_candidates_pixel<V>(ctx, x, y); break;
    case B2_PASS_RESTIR_DI_TEMPORAL: restir_di_temporal_pixel<V>(ctx, x, y); break;
    case B2_PASS_RESTIR_DI_SPATIAL: restir_di_spatial_pixel<V>(ctx, x, y); break;
    case B2_PASS_RESTIR_DI_SPATIOTEMPORAL: restir_di_spatiotemporal_pixel<V>(ctx, x, y); break;
    default: break;
    }
}

#endif
