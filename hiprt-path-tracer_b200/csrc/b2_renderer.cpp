/*
 * b2_renderer.cpp -- headless host mirror of the reference's renderer front end (include/b2pt.h, group 2).
 *
 * Mirrors GPURenderer's public surface (set_scene / set_camera / set_envmap / render / reset / get_render_data,
 * /root/reference/src/Renderer/GPURenderer.h:61-346; set_scene uploads at GPURenderer.cpp:769-939) on top of the
 * device layer of this library -- it only calls b2pt_* group-1 entry points, the same ones a maintainer would bind
 * from the reference's own GPURenderer. Pass order and host RNG draws: B2FrameLoop (b2_frame_loop.h).
 */
#include "b2pt.h"

#include "b2_comm.h"
#include "b2_frame_loop.h"
#include "b2_host_scene.h"
#include "b2_material.h"
#include "b2pt_alias_table.h"

#include <cstddef>
#include <cstdio>
#include <cstring>
#include <new>
#include <string>
#include <vector>

void b2_set_error(const char* fmt, ...);

#define B2_TRY(call)          \
    do                        \
    {                         \
        int rc_ = (call);     \
        if (rc_ != B2PT_OK)   \
            return rc_;       \
    } while (0)

/* inside b2pt_renderer_create: a failure releases everything allocated so far */
#define B2_TRY_R(call)                  \
    do                                  \
    {                                   \
        int rc_ = (call);               \
        if (rc_ != B2PT_OK)             \
        {                               \
            b2pt_renderer_destroy(r);   \
            return rc_;                 \
        }                               \
    } while (0)

/* C ABI boundary: no C++ exception may cross it */
#define B2_GUARDED(...)                                                    \
    try                                                                    \
    {                                                                      \
        __VA_ARGS__                                                            \
    }                                                                      \
    catch (const std::bad_alloc&)                                          \
    {                                                                      \
        b2_set_error("out of host memory");                                \
        return B2PT_ERR_IO;                                                \
    }                                                                      \
    catch (const std::exception& e)                                        \
    {                                                                      \
        b2_set_error("unexpected exception: %s", e.what());                \
        return B2PT_ERR_INVALID_ARGUMENT;                                  \
    }

struct b2pt_renderer
{
    b2pt_ctx* ctx = nullptr;
    int width = 0, height = 0;
    std::string variant;
    int precise = 1;
    int counters = 0;
    bool kernels_stale = true;
    std::vector<std::pair<std::string, int>> macros; /* b2pt_renderer_set_macro: kernel options on top of the variant's, in the order they were set */
    bool alpha_test = false; /* the current scene has materials that are not fully opaque: kernels are asked for B2PT_ALPHA_TEST */
    bool profile_passes = false;

    B2FrameLoop loop;
    B2HostScene scene;
    bool has_scene = false;

    b2pt_stream stream = nullptr;
    b2pt_event ev_start = nullptr, ev_stop = nullptr;
    bool timed = false;
    std::vector<void*> frame_allocs, scene_allocs, lut_handles, envmap_allocs;
    struct Upload { void* device; const void* host; size_t bytes; bool pinned; };
    bool scene_pin_attempted = false; /* the host copies are page-locked on the first b2pt_renderer_reupload_scene */
    std::vector<Upload> scene_uploads; /* (device, host, bytes) of every scene array, for b2pt_renderer_reupload_scene */
    b2pt_bvh* bvh = nullptr;
    void* textures = nullptr; /* b2pt_textures_create: the descriptor array RenderBuffers::material_textures points to */
    b2pt_kernel* kernels[B2_PASS_COUNT] = {};
    double pass_ms[B2_PASS_COUNT] = {};
    uint64_t pass_launches[B2_PASS_COUNT] = {};
    /* multi-GPU: this renderer owns the interleaved B2PT_BAND_ROWS-row bands `rank`, rank + world, ... of the frame */
    int rank = 0, world = 1;
    b2pt_comm* comm = nullptr;      /* not owned */
    int band_rows = B2PT_BAND_ROWS; /* 32 for the ReSTIR DI variants: a band must be at least as tall as the halo */
    bool dynamic_camera = false;
    int halo_rows = 0;              /* halo the camera-row list below was built for */
    int* d_camera_rows = nullptr;   /* own bands + halo rows (ReSTIR DI partitions): what CameraRays covers */
    int camera_row_count = 0;
    int owned_row_count = 0;
    uint8_t* d_status = nullptr; /* [0] still_one_ray_active, [4..7] pixel_count_converged_so_far */
    void* d_reservoirs[3] = {};

    size_t pixels() const { return (size_t)width * height; }
};

namespace
{
int alloc_zero(b2pt_renderer* r, std::vector<void*>& owner, size_t bytes, void** out)
{
    B2_TRY(b2pt_malloc(r->ctx, bytes, out));
    owner.push_back(*out);
    return b2pt_memset(r->ctx, *out, 0, bytes, nullptr);
}

/* Every reservoir of a buffer = a default-constructed ReSTIRDIReservoir (Reservoir.h:22-50): all words zero except the
 * emissive triangle index, which is -1. The buffer is 9 SoA planes of one 32-bit word per pixel; the index is plane 6
 * (DI_PLANE_TRIANGLE, b2_restir_di_shading.h). */
int clear_reservoirs(b2pt_renderer* r, void* buffer, b2pt_stream stream)
{
    const size_t plane_bytes = r->pixels() * 4;
    B2_TRY(b2pt_memset(r->ctx, buffer, 0, plane_bytes * 9, stream));
    return b2pt_memset(r->ctx, static_cast<char*>(buffer) + 6 * plane_bytes, 0xFF, plane_bytes, stream);
}

template <class T>
int upload(b2pt_renderer* r, std::vector<void*>& owner, const std::vector<T>& host, T** out)
{
    *out = nullptr;
    if (host.empty())
        return B2PT_OK;
    void* d = nullptr;
    B2_TRY(b2pt_malloc(r->ctx, host.size() * sizeof(T), &d));
    owner.push_back(d);
    *out = reinterpret_cast<T*>(d);
    if (&owner == &r->scene_allocs)
        r->scene_uploads.push_back({ d, host.data(), host.size() * sizeof(T), false });
    return b2pt_memcpy_h2d(r->ctx, d, host.data(), host.size() * sizeof(T), nullptr);
}

/* Releases the page locks on the host copies of the scene arrays; must run before those vectors are resized or freed. */
void unpin_scene(b2pt_renderer* r)
{
    b2pt_stream_synchronize(r->ctx, r->stream);
    for (b2pt_renderer::Upload& u : r->scene_uploads)
        if (u.pinned)
        {
            b2pt_host_unregister(r->ctx, u.host);
            u.pinned = false;
        }
    r->scene_pin_attempted = false;
}

void free_all(b2pt_renderer* r, std::vector<void*>& owner)
{
    for (void* p : owner)
        b2pt_free(r->ctx, p);
    owner.clear();
}

struct VariantKey { const char* name; int bsdf, direct, base, later; bool restir; bool presampling; int bias_correction; int restir_options = 4 | 8; int presampling_strategy = 1; };
const VariantKey kVariants[] = {
    { "c1", 1, 1, 0, 3, false, false, 4 },
    { "c2", 0, 3, 1, 3, false, false, 4 },
    { "c3", 0, 4, 1, 3, false, false, 4 },
    { "c4", 0, 5, 1, 3, true, false, 4 },
    { "c4p", 0, 5, 1, 3, true, true, 4 }, /* c4 with ReSTIR_DI_DoLightPresampling */
    { "c4pu", 0, 5, 1, 3, true, true, 4, 4 | 8, 0 }, /* ... and ReSTIR_DI_LightPresamplingStrategy = LSS_BASE_UNIFORM */
    /* c4 with another ReSTIR_DI_BiasCorrectionWeights (ReSTIRDIOptions.h:9-16): 1/Z, MIS-like, generalised balance heuristic, defensive
     * pairwise MIS, symmetric and asymmetric ratio */
    { "c4b0", 0, 5, 1, 3, true, false, 0 }, /* 1/M: fused pass, or temporal reuse alone (restir_di_do_spatial_reuse_pass = 0) */
    { "c4b1", 0, 5, 1, 3, true, false, 1 },
    { "c4b2", 0, 5, 1, 3, true, false, 2 },
    { "c4b3", 0, 5, 1, 3, true, false, 3 },
    { "c4b5", 0, 5, 1, 3, true, false, 5 },
    { "c4b6", 0, 5, 1, 3, true, false, 6 }, /* symmetric ratio */
    { "c4b7", 0, 5, 1, 3, true, false, 7 }, /* asymmetric ratio */
    /* c4 with one of the four ReSTIR DI visibility macros (ReSTIRDIOptions.h:50-84) away from its default; bits: 1 InitialTargetFunctionVisibility,
     * 2 SpatialTargetFunctionVisibility, 4 DoVisibilityReuse, 8 BiasCorrectionUseVisibility (defaults 4 | 8) */
    /* c4 with ReSTIR_DI_LaterBouncesSamplingStrategy (ReSTIRDIOptions.h:18-21,140) = uniform one light / BSDF only / light + BSDF MIS */
    { "c4l0", 0, 5, 1, 0, true, false, 4 },
    { "c4l1", 0, 5, 1, 1, true, false, 4 },
    { "c4l2", 0, 5, 1, 2, true, false, 4 },
    /* several at once: balance heuristic + no bias-correction visibility + spatial visibility + MIS later; symmetric ratio + no visibility reuse + initial visibility + BSDF later */
    { "c4x1", 0, 5, 1, 2, true, false, 3, 2 | 4 },
    { "c4x2", 0, 5, 1, 1, true, false, 6, 1 | 8 },
    { "c4o", 0, 5, 1, 3, true, false, 4, 4 | 8 | 16 }, /* ReSTIR_DI_DoOptimalVisibilitySampling (bit 16; read by the spatial pass: restir_di_do_fused_spatiotemporal = 0) */
    { "c4v1", 0, 5, 1, 3, true, false, 4, 1 | 4 | 8 }, /* InitialTargetFunctionVisibility on */
    { "c4v2", 0, 5, 1, 3, true, false, 4, 2 | 4 | 8 }, /* SpatialTargetFunctionVisibility on */
    { "c4v3", 0, 5, 1, 3, true, false, 4, 8 },         /* DoVisibilityReuse off */
    { "c4v4", 0, 5, 1, 3, true, false, 4, 4 },         /* BiasCorrectionUseVisibility off */
    { "c4v5", 0, 5, 1, 3, true, false, 4, 1 | 2 | 8 }, /* both target-function visibilities on, visibility reuse off */
};

const VariantKey* find_variant(const std::string& name)
{
    for (const VariantKey& v : kVariants)
        if (name == v.name)
            return &v;
    return nullptr;
}

const char* entry_name_of_pass(int pass)
{
    switch (pass)
    {
    case B2_PASS_CAMERA_RAYS: return "CameraRays";
    case B2_PASS_MEGAKERNEL: return "MegaKernel";
    case B2_PASS_RESTIR_DI_PRESAMPLING: return "ReSTIR_DI_LightsPresampling";
    case B2_PASS_RESTIR_DIRECTIONAL_REUSE: return "ReSTIR_Directional_Reuse_Compute";
    case B2_PASS_RESTIR_DI_INITIAL_CANDIDATES: return "ReSTIR_DI_InitialCandidates";
    case B2_PASS_RESTIR_DI_TEMPORAL: return "ReSTIR_DI_TemporalReuse";
    case B2_PASS_RESTIR_DI_SPATIAL: return "ReSTIR_DI_SpatialReuse";
    case B2_PASS_RESTIR_DI_SPATIOTEMPORAL: return "ReSTIR_DI_SpatiotemporalReuse";
    default: return nullptr;
    }
}

/* the kernel options a variant stands for (what resolve_kernels passes, before the macros set one by one) */
std::vector<b2pt_option> variant_options(const VariantKey* v, int precise, int counters, int alpha_test)
{
    return {
        { "BSDFOverride", v->bsdf }, { "DirectLightSamplingStrategy", v->direct }, { "DirectLightSamplingBaseStrategy", v->base },
        { "ReSTIR_DI_LaterBouncesSamplingStrategy", v->later }, { "ReSTIR_DI_DoLightPresampling", v->presampling ? 1 : 0 },
        { "B2PT_PRECISE", precise }, { "B2PT_COUNTERS", counters }, { "B2PT_ALPHA_TEST", alpha_test },
        { "ReSTIR_DI_BiasCorrectionWeights", v->bias_correction }, { "ReSTIR_DI_LightPresamplingStrategy", v->presampling_strategy },
        { "ReSTIR_DI_InitialTargetFunctionVisibility", (v->restir_options & 1) ? 1 : 0 }, { "ReSTIR_DI_SpatialTargetFunctionVisibility", (v->restir_options & 2) ? 1 : 0 },
        { "ReSTIR_DI_DoVisibilityReuse", (v->restir_options & 4) ? 1 : 0 }, { "ReSTIR_DI_BiasCorrectionUseVisibility", (v->restir_options & 8) ? 1 : 0 },
        { "ReSTIR_DI_DoOptimalVisibilitySampling", (v->restir_options & 16) ? 1 : 0 },
    };
}

bool variant_has_pass(const VariantKey* v, int pass)
{
    if (!entry_name_of_pass(pass))
        return false;
    bool restir_pass = pass != B2_PASS_CAMERA_RAYS && pass != B2_PASS_MEGAKERNEL;
    if (restir_pass && !v->restir)
        return false;
    if (pass == B2_PASS_RESTIR_DI_SPATIAL && v->bias_correction == 0)
        return false; /* 1/M: no spatial reuse kernel (the reference's does not compile with these weights); fused or temporal-only reuse */
    return pass != B2_PASS_RESTIR_DI_PRESAMPLING || v->presampling;
}

int resolve_kernels(b2pt_renderer* r)
{
    if (!r->kernels_stale)
        return B2PT_OK;
    const VariantKey* v = find_variant(r->variant);
    for (int pass = 0; pass < B2_PASS_COUNT; pass++)
    {
        if (r->kernels[pass])
        {
            b2pt_kernel_destroy(r->kernels[pass]);
            r->kernels[pass] = nullptr;
        }
        if (!variant_has_pass(v, pass))
            continue;
        const char* entry = entry_name_of_pass(pass);
        /* the variant's options first, then the ones set one by one (GPUKernelCompilerOptions::set_macro_value): a later entry wins */
        std::vector<b2pt_option> all = variant_options(v, r->precise, r->counters, r->alpha_test ? 1 : 0);
        for (const auto& macro : r->macros)
            all.push_back(b2pt_option{ macro.first.c_str(), macro.second });
        B2_TRY(b2pt_kernel_get(r->ctx, entry, all.data(), (int)all.size(), &r->kernels[pass]));
    }
    r->kernels_stale = false;
    return B2PT_OK;
}

/* Forgets the current scene: after this no device pointer of the render data refers to scene memory, so a failed
 * upload leaves the renderer in the "no scene set" state (render / reupload return an error) instead of pointing at
 * freed allocations. */
void invalidate_scene(b2pt_renderer* r, bool keep_bvh = false)
{
    unpin_scene(r);
    r->has_scene = false;
    r->scene_uploads.clear();
    b2pt_render_data& rd = r->loop.rd;
    b2pt_rgb* framebuffer = rd.buffers.accumulated_ray_colors;
    std::memset(&rd.buffers, 0, sizeof(rd.buffers));
    rd.buffers.accumulated_ray_colors = framebuffer;
    rd.GPU_BVH = nullptr;
    rd.light_GPU_BVH = nullptr;
    free_all(r, r->scene_allocs);
    if (r->bvh && !keep_bvh)
    {
        b2pt_bvh_destroy(r->ctx, r->bvh);
        r->bvh = nullptr;
    }
}

/* (Re)builds the list of rows CameraRays covers under a ReSTIR DI partition: this rank's bands plus `halo` rows around them,
 * halo = how far a reuse pass looks (fused pass: spatial neighbours around the temporal neighbour). */
int update_partition_rows(b2pt_renderer* r)
{
    const b2pt_restir_di_settings& di = r->loop.rd.render_settings.restir_di_settings;
    int halo = di.common_spatial_pass.reuse_radius + di.common_temporal_pass.neighbor_search_radius + 2;
    if (r->dynamic_camera)
        halo = r->height; /* a reprojected temporal neighbour can be anywhere */
    if (halo > r->band_rows && halo < r->height)
    {
        b2_set_error("ReSTIR DI partition: reuse radius %d needs a halo of %d rows, more than the %d-row bands", di.common_spatial_pass.reuse_radius, halo, r->band_rows);
        return B2PT_ERR_INVALID_ARGUMENT;
    }
    if (r->d_camera_rows && halo == r->halo_rows)
        return B2PT_OK;
    const std::vector<int> rows = b2_partition_rows(r->height, r->rank, r->world, r->band_rows, halo);
    if (r->d_camera_rows)
    {
        B2_TRY(b2pt_stream_synchronize(r->ctx, r->stream));
        b2pt_free(r->ctx, r->d_camera_rows);
        r->d_camera_rows = nullptr;
    }
    void* d = nullptr;
    B2_TRY(b2pt_malloc(r->ctx, rows.size() * sizeof(int), &d));
    B2_TRY(b2pt_memcpy_h2d(r->ctx, d, rows.data(), rows.size() * sizeof(int), nullptr));
    r->d_camera_rows = (int*)d;
    r->camera_row_count = (int)rows.size();
    r->halo_rows = halo;
    return B2PT_OK;
}

/* keep_bvh: the geometry did not change (update_all_materials): every scene array is re-derived and re-uploaded, the wide BVH stays */
int upload_scene(b2pt_renderer* r, bool keep_bvh = false)
{
    keep_bvh = keep_bvh && r->bvh != nullptr;
    invalidate_scene(r, keep_bvh);
    B2HostScene& s = r->scene;
    std::string error;
    if (!b2_host_scene_prepare(s, error, /* build_bvh = */ false))
    {
        b2_set_error("scene rejected: %s", error.c_str());
        return B2PT_ERR_UNSUPPORTED_SCENE;
    }
    if (r->alpha_test != s.has_alpha_tested_materials)
    {
        r->alpha_test = s.has_alpha_tested_materials;
        r->kernels_stale = true;
    }

    b2pt_render_data& rd = r->loop.rd;
    b2pt_render_buffers& b = rd.buffers;
    float* d_pos = nullptr;
    float* d_nrm = nullptr;
    float* d_uv = nullptr;
    float* d_emission = nullptr;
    B2_TRY(upload(r, r->scene_allocs, s.blob.indices, &b.triangles_indices));
    B2_TRY(upload(r, r->scene_allocs, s.blob.positions, &d_pos));
    B2_TRY(upload(r, r->scene_allocs, s.blob.normals, &d_nrm));
    B2_TRY(upload(r, r->scene_allocs, s.blob.texcoords, &d_uv));
    B2_TRY(upload(r, r->scene_allocs, s.blob.has_normals, &b.has_vertex_normals));
    B2_TRY(upload(r, r->scene_allocs, s.triangle_areas, &b.triangles_areas));
    B2_TRY(upload(r, r->scene_allocs, s.blob.material_indices, &b.material_indices));
    B2_TRY(upload(r, r->scene_allocs, s.material_opaque, &b.material_opaque));
    B2_TRY(upload(r, r->scene_allocs, s.emissive_triangles, &b.emissive_triangles_primitive_indices));
    b.vertices_positions = reinterpret_cast<b2pt_float3*>(d_pos);
    b.vertex_normals = reinterpret_cast<b2pt_float3*>(d_nrm);
    b.texcoords = reinterpret_cast<b2pt_float2*>(d_uv);
    b.emissive_triangles_count = (int)s.emissive_triangles.size();
    B2_TRY(upload(r, r->scene_allocs, s.emissive_triangles_and_textures, &b.emissive_triangles_primitive_indices_and_emissive_textures));
    B2_TRY(upload(r, r->scene_allocs, s.power_alias_alias, &b.emissives_power_alias_table.alias_table_alias));
    B2_TRY(upload(r, r->scene_allocs, s.power_alias_probas, &b.emissives_power_alias_table.alias_table_probas));
    b.emissives_power_alias_table.sum_elements = s.power_sum;
    b.emissives_power_alias_table.size = (unsigned)s.emissive_triangles.size();
    b.material_textures = nullptr;
    if (r->textures)
    {
        b2pt_textures_destroy(r->ctx, r->textures);
        r->textures = nullptr;
    }
    if (!s.blob.textures.empty())
    {
        std::vector<const uint8_t*> texels;
        std::vector<int> widths, heights, channels;
        for (const B2SceneTexture& t : s.blob.textures)
        {
            texels.push_back(t.texels.data());
            widths.push_back((int)t.width);
            heights.push_back((int)t.height);
            channels.push_back((int)t.channels);
        }
        B2_TRY(b2pt_textures_create(r->ctx, texels.data(), widths.data(), heights.data(), channels.data(), (int)texels.size(), &r->textures));
        b.material_textures = r->textures;
    }

    B2MaterialSoAHost& m = s.materials;
    b2pt_materials_soa& d = b.materials_buffer;
    B2_TRY(upload(r, r->scene_allocs, m.flags, &d.flags));
    B2_TRY(upload(r, r->scene_allocs, m.emission, &d_emission));
    d.emission = reinterpret_cast<b2pt_rgb*>(d_emission);
    B2_TRY(upload(r, r->scene_allocs, m.base_color_roughness, &d.base_color_roughness));
    B2_TRY(upload(r, r->scene_allocs, m.oren_nayar_sigma, &d.oren_nayar_sigma));
    B2_TRY(upload(r, r->scene_allocs, m.metallic_F90_and_metallic, &d.metallic_F90_and_metallic));
    B2_TRY(upload(r, r->scene_allocs, m.metallic_F82_packed_and_diffuse_transmission, &d.metallic_F82_packed_and_diffuse_transmission));
    B2_TRY(upload(r, r->scene_allocs, m.metallic_F90_falloff_exponent, &d.metallic_F90_falloff_exponent));
    B2_TRY(upload(r, r->scene_allocs, m.anisotropy_and_rotation_and_second_roughness, &d.anisotropy_and_rotation_and_second_roughness));
    B2_TRY(upload(r, r->scene_allocs, m.specular_color_and_tint_factor, &d.specular_color_and_tint_factor));
    B2_TRY(upload(r, r->scene_allocs, m.specular_and_darkening_and_coat_roughness, &d.specular_and_darkening_and_coat_roughness));
    B2_TRY(upload(r, r->scene_allocs, m.coat_medium_thickness, &d.coat_medium_thickness));
    B2_TRY(upload(r, r->scene_allocs, m.coat_and_medium_absorption, &d.coat_and_medium_absorption));
    B2_TRY(upload(r, r->scene_allocs, m.coat_roughening_darkening_anisotropy_and_rotation, &d.coat_roughening_darkening_anisotropy_and_rotation));
    B2_TRY(upload(r, r->scene_allocs, m.coat_ior, &d.coat_ior));
    B2_TRY(upload(r, r->scene_allocs, m.sheen_and_color, &d.sheen_and_color));
    B2_TRY(upload(r, r->scene_allocs, m.ior, &d.ior));
    B2_TRY(upload(r, r->scene_allocs, m.absorption_color_packed, &d.absorption_color_packed));
    B2_TRY(upload(r, r->scene_allocs, m.absorption_at_distance, &d.absorption_at_distance));
    B2_TRY(upload(r, r->scene_allocs, m.sheen_roughness_transmission_dispersion_thin_film, &d.sheen_roughness_transmission_dispersion_thin_film));
    B2_TRY(upload(r, r->scene_allocs, m.dispersion_abbe_number, &d.dispersion_abbe_number));
    B2_TRY(upload(r, r->scene_allocs, m.thin_film_ior, &d.thin_film_ior));
    B2_TRY(upload(r, r->scene_allocs, m.thin_film_thickness, &d.thin_film_thickness));
    B2_TRY(upload(r, r->scene_allocs, m.thin_film_kappa_3, &d.thin_film_kappa_3));
    B2_TRY(upload(r, r->scene_allocs, m.thin_film_base_ior_override, &d.thin_film_base_ior_override));
    B2_TRY(upload(r, r->scene_allocs, m.alpha_thin_film_hue_dielectric_priority, &d.alpha_thin_film_hue_dielectric_priority));
    B2_TRY(upload(r, r->scene_allocs, m.normal_map_emission_index, &d.normal_map_emission_index));
    B2_TRY(upload(r, r->scene_allocs, m.base_color_roughness_metallic_index, &d.base_color_roughness_metallic_index));
    B2_TRY(upload(r, r->scene_allocs, m.roughness_and_metallic_index, &d.roughness_and_metallic_index));
    B2_TRY(upload(r, r->scene_allocs, m.anisotropic_specular_index, &d.anisotropic_specular_index));
    B2_TRY(upload(r, r->scene_allocs, m.coat_sheen_index, &d.coat_sheen_index));
    B2_TRY(upload(r, r->scene_allocs, m.specular_transmission_index, &d.specular_transmission_index));

    if (!keep_bvh)
        B2_TRY(b2pt_bvh_build(r->ctx, s.blob.indices.data(), (int)s.blob.n_triangles(), s.blob.positions.data(), (int)s.blob.n_vertices(), 0, nullptr, &r->bvh));
    rd.GPU_BVH = b2pt_bvh_device_handle(r->bvh);
    rd.light_GPU_BVH = nullptr;

    /* camera (static): what Camera::to_hiprt hands over every sample */
    b2pt_camera& cam = r->loop.camera;
    std::memcpy(cam.inverse_view, s.blob.camera.inverse_view, 64);
    std::memcpy(cam.inverse_projection, s.blob.camera.inverse_projection, 64);
    std::memcpy(cam.view_projection, s.blob.camera.view_projection, 64);
    cam.position = b2pt_float3{ s.blob.camera.position[0], s.blob.camera.position[1], s.blob.camera.position[2] };
    cam.vertical_fov = s.blob.camera.vertical_fov;
    cam.sensor_width = r->width;
    cam.sensor_height = r->height;
    cam.do_jittering = 1;
    rd.current_camera = cam;
    r->has_scene = true;
    return B2PT_OK;
}
} // namespace

extern "C" {

int b2pt_renderer_create(b2pt_ctx* ctx, int width, int height, const char* variant, int precise, b2pt_renderer** out_renderer)
{
    if (!ctx || !variant || !out_renderer || width <= 0 || height <= 0)
    {
        b2_set_error("b2pt_renderer_create: invalid argument");
        return B2PT_ERR_INVALID_ARGUMENT;
    }
    if (!find_variant(variant))
    {
        b2_set_error("b2pt_renderer_create: unknown variant '%s' (c1, c2, c3, c4)", variant);
        return B2PT_ERR_NO_VARIANT;
    }
    B2_TRY(b2pt_ctx_make_current(ctx));
    b2pt_renderer* r = new (std::nothrow) b2pt_renderer();
    if (!r) { b2_set_error("out of host memory"); return B2PT_ERR_INVALID_ARGUMENT; }
    r->ctx = ctx;
    r->width = width;
    r->height = height;
    r->variant = variant;
    r->precise = precise ? 1 : 0;
    r->loop.restir_di = find_variant(variant)->restir;
    r->loop.restir_di_presampling = find_variant(variant)->presampling;
    B2_TRY_R(b2pt_stream_create(ctx, &r->stream));
    B2_TRY_R(b2pt_event_create(ctx, &r->ev_start));
    B2_TRY_R(b2pt_event_create(ctx, &r->ev_stop));

    const size_t n = r->pixels();
    b2pt_render_data& rd = r->loop.rd;
    rd.render_settings.render_resolution = b2pt_int2{ width, height };
    void* p = nullptr;
    B2_TRY_R(alloc_zero(r, r->frame_allocs, n * sizeof(b2pt_rgb), &p)); rd.buffers.accumulated_ray_colors = (b2pt_rgb*)p;
    B2_TRY_R(alloc_zero(r, r->frame_allocs, n, &p)); rd.aux_buffers.pixel_active = (uint8_t*)p;
    B2_TRY_R(alloc_zero(r, r->frame_allocs, n * sizeof(b2pt_float3), &p)); rd.aux_buffers.denoiser_normals = (b2pt_float3*)p;
    B2_TRY_R(alloc_zero(r, r->frame_allocs, n * sizeof(b2pt_rgb), &p)); rd.aux_buffers.denoiser_albedo = (b2pt_rgb*)p;
    B2_TRY_R(alloc_zero(r, r->frame_allocs, n * sizeof(int), &p)); rd.aux_buffers.pixel_sample_count = (int*)p;
    B2_TRY_R(alloc_zero(r, r->frame_allocs, n * sizeof(float), &p)); rd.aux_buffers.pixel_squared_luminance = (float*)p;
    B2_TRY_R(alloc_zero(r, r->frame_allocs, n * sizeof(int), &p)); rd.aux_buffers.pixel_converged_sample_count = (int*)p;
    B2_TRY_R(alloc_zero(r, r->frame_allocs, 16, &p)); r->d_status = (uint8_t*)p;
    rd.aux_buffers.still_one_ray_active = r->d_status;
    rd.aux_buffers.pixel_count_converged_so_far = (unsigned int*)(r->d_status + 4);
    b2pt_gbuffer* gb[2] = { &rd.g_buffer, &rd.g_buffer_prev_frame };
    /* the previous-frame G-buffer and the reservoir buffers exist only for the ReSTIR DI variants (use_prev_frame_g_buffer) */
    for (int i = 0; i < (r->loop.restir_di ? 2 : 1); i++)
    {
        B2_TRY_R(alloc_zero(r, r->frame_allocs, n * 108, &p)); gb[i]->materials = p;
        B2_TRY_R(alloc_zero(r, r->frame_allocs, n * sizeof(int), &p)); gb[i]->first_hit_prim_index = (int*)p;
        B2_TRY_R(alloc_zero(r, r->frame_allocs, n * sizeof(b2pt_float3), &p)); gb[i]->primary_hit_position = (b2pt_float3*)p;
        B2_TRY_R(alloc_zero(r, r->frame_allocs, n * 4, &p)); gb[i]->shading_normals = (uint32_t*)p;
        B2_TRY_R(alloc_zero(r, r->frame_allocs, n * 4, &p)); gb[i]->geometric_normals = (uint32_t*)p;
    }
    for (int i = 0; i < (r->loop.restir_di ? 3 : 0); i++)
    {
        B2_TRY_R(alloc_zero(r, r->frame_allocs, n * 36, &r->d_reservoirs[i]));
        B2_TRY_R(clear_reservoirs(r, r->d_reservoirs[i], nullptr));
    }
    if (r->loop.restir_di)
    {
        /* per-pixel radius + 64-bit direction mask of the adaptive-directional spatial reuse (ReSTIRDIRenderPass.h) */
        B2_TRY_R(alloc_zero(r, r->frame_allocs, n * sizeof(unsigned long long), &p));
        rd.render_settings.restir_di_settings.common_spatial_pass.per_pixel_spatial_reuse_directions_mask_ull = (unsigned long long*)p;
        B2_TRY_R(alloc_zero(r, r->frame_allocs, n, &p));
        rd.render_settings.restir_di_settings.common_spatial_pass.per_pixel_spatial_reuse_radius = (uint8_t*)p;
    }
    if (r->loop.restir_di_presampling)
    {
        /* ReSTIRDIRenderPass allocates number_of_subsets * subset_size presampled lights of 48 bytes (defaults 128 x 1024) */
        const b2pt_restir_di_settings& di = rd.render_settings.restir_di_settings;
        B2_TRY_R(alloc_zero(r, r->frame_allocs, (size_t)di.light_presampling.number_of_subsets * di.light_presampling.subset_size * 48, &p));
        rd.render_settings.restir_di_settings.light_presampling.light_samples = p;
    }
    r->loop.initial_reservoirs = r->d_reservoirs[0];
    r->loop.spatial_reservoirs_1 = r->d_reservoirs[1];
    r->loop.spatial_reservoirs_2 = r->d_reservoirs[2];
    rd.render_settings.restir_di_settings.initial_candidates.output_reservoirs = r->d_reservoirs[0];
    rd.render_settings.restir_di_settings.restir_output_reservoirs = r->d_reservoirs[1];
    *out_renderer = r;
    return B2PT_OK;
}

int b2pt_renderer_destroy(b2pt_renderer* r)
{
    if (!r) return B2PT_OK;
    b2pt_ctx_make_current(r->ctx);
    if (r->stream)
        unpin_scene(r);
    for (int pass = 0; pass < B2_PASS_COUNT; pass++)
        if (r->kernels[pass])
            b2pt_kernel_destroy(r->kernels[pass]);
    if (r->d_camera_rows)
        b2pt_free(r->ctx, r->d_camera_rows);
    free_all(r, r->frame_allocs);
    free_all(r, r->scene_allocs);
    free_all(r, r->envmap_allocs);
    for (void* h : r->lut_handles)
        b2pt_lut_destroy(r->ctx, h);
    if (r->bvh) b2pt_bvh_destroy(r->ctx, r->bvh);
    if (r->textures) b2pt_textures_destroy(r->ctx, r->textures);
    b2pt_event_destroy(r->ctx, r->ev_start);
    b2pt_event_destroy(r->ctx, r->ev_stop);
    b2pt_stream_destroy(r->ctx, r->stream);
    delete r;
    return B2PT_OK;
}

int b2pt_renderer_set_scene_file(b2pt_renderer* r, const char* scene_blob_path)
{
    if (!r || !scene_blob_path) { b2_set_error("null argument"); return B2PT_ERR_INVALID_ARGUMENT; }
    B2_TRY(b2pt_ctx_make_current(r->ctx));
    B2_GUARDED(
        /* parsed into a temporary: a file that cannot be read leaves the current scene untouched and valid */
        B2Scene incoming;
        std::string err = b2scene_read(scene_blob_path, incoming);
        if (!err.empty())
        {
            b2_set_error("%s", err.c_str());
            return B2PT_ERR_IO;
        }
        invalidate_scene(r);
        r->scene.blob = std::move(incoming);
        return upload_scene(r);
    )
}

int b2pt_renderer_set_scene_arrays(b2pt_renderer* r, const float* positions, const float* normals, const float* texcoords,
                                   const uint8_t* has_normals, int vertex_count, const int* indices, const int* material_indices,
                                   int triangle_count, const float* materials, int material_count, const float* camera52)
{
    if (!r || !positions || !indices || !material_indices || !materials || !camera52 || vertex_count <= 0 || triangle_count <= 0 || material_count <= 0)
    {
        b2_set_error("b2pt_renderer_set_scene_arrays: invalid argument");
        return B2PT_ERR_INVALID_ARGUMENT;
    }
    B2_TRY(b2pt_ctx_make_current(r->ctx));
    B2_GUARDED(
    invalidate_scene(r);
    B2Scene& s = r->scene.blob;
    s.positions.assign(positions, positions + 3 * (size_t)vertex_count);
    if (normals) s.normals.assign(normals, normals + 3 * (size_t)vertex_count); else s.normals.assign(3 * (size_t)vertex_count, 0.0f);
    if (texcoords) s.texcoords.assign(texcoords, texcoords + 2 * (size_t)vertex_count); else s.texcoords.assign(2 * (size_t)vertex_count, 0.0f);
    if (has_normals) s.has_normals.assign(has_normals, has_normals + vertex_count); else s.has_normals.assign(vertex_count, normals ? 1 : 0);
    s.indices.assign(indices, indices + 3 * (size_t)triangle_count);
    s.material_indices.assign(material_indices, material_indices + triangle_count);
    s.materials.assign(materials, materials + (size_t)material_count * B2MAT_FLOAT_COUNT);
    std::memcpy(&s.camera, camera52, sizeof(B2SceneCamera));
    return upload_scene(r);
    )
}

/* GPURenderer::update_all_materials (GPURenderer.cpp:1022-1040): new values for every material of the current scene. Everything
 * derived from materials is rebuilt the way set_scene builds it (packed SoA, opacity flags, emissive triangle list, power alias
 * table) and re-uploaded; the geometry and its BVH stay. */
int b2pt_renderer_update_materials(b2pt_renderer* r, const float* materials, int material_count)
{
    if (!r || !materials || !r->has_scene) { b2_set_error("b2pt_renderer_update_materials: no scene / null argument"); return B2PT_ERR_INVALID_ARGUMENT; }
    if ((size_t)material_count != r->scene.blob.n_materials())
    {
        b2_set_error("b2pt_renderer_update_materials: the scene has %zu materials, %d given", r->scene.blob.n_materials(), material_count);
        return B2PT_ERR_INVALID_ARGUMENT;
    }
    B2_TRY(b2pt_ctx_make_current(r->ctx));
    B2_GUARDED(
    B2_TRY(b2pt_stream_synchronize(r->ctx, r->stream));
    r->scene.blob.materials.assign(materials, materials + (size_t)material_count * B2MAT_FLOAT_COUNT);
    return upload_scene(r, /* keep_bvh = */ true);
    )
}

/* RendererEnvmap::do_animation / update_renderer (RendererEnvmap.cpp:129-181): the two 3x3 rotations between envmap and world space,
 * row-major as in WorldSettings (the reference host computes them with glm::orientate3 and its inverse and hands them over). */
int b2pt_renderer_set_envmap_matrices(b2pt_renderer* r, const float* envmap_to_world9, const float* world_to_envmap9)
{
    if (!r || !envmap_to_world9 || !world_to_envmap9) { b2_set_error("null argument"); return B2PT_ERR_INVALID_ARGUMENT; }
    std::memcpy(r->loop.rd.world_settings.envmap_to_world_matrix, envmap_to_world9, 36);
    std::memcpy(r->loop.rd.world_settings.world_to_envmap_matrix, world_to_envmap9, 36);
    return B2PT_OK;
}

/* Re-sends every scene array (geometry, per-triangle data, material SoA, alias table, wide-BVH blob) from the host
 * copies to their existing device allocations, asynchronously on the renderer's stream: the per-step host->device
 * traffic of the end-to-end measurement (what GPURenderer::set_scene hands to the device layer, BVH build excluded). */
int b2pt_renderer_reupload_scene(b2pt_renderer* r, uint64_t* out_bytes)
{
    if (!r || !r->has_scene) { b2_set_error("no scene"); return B2PT_ERR_INVALID_ARGUMENT; }
    B2_TRY(b2pt_ctx_make_current(r->ctx));
    uint64_t total = 0;
    if (!r->scene_pin_attempted)
    {
        /* best effort, once per scene: a range that cannot be page-locked is still copied (through the driver's staging buffer) */
        r->scene_pin_attempted = true;
        for (b2pt_renderer::Upload& u : r->scene_uploads)
            u.pinned = u.bytes >= 65536 && b2pt_host_register(r->ctx, u.host, u.bytes) == B2PT_OK;
    }
    for (const b2pt_renderer::Upload& u : r->scene_uploads)
    {
        B2_TRY(b2pt_memcpy_h2d(r->ctx, u.device, u.host, u.bytes, r->stream));
        total += u.bytes;
    }
    uint64_t bvh_bytes = 0;
    B2_TRY(b2pt_bvh_reupload(r->ctx, r->bvh, r->stream, &bvh_bytes));
    total += bvh_bytes;
    if (out_bytes) *out_bytes = total;
    return B2PT_OK;
}

int b2pt_renderer_set_luts_file(b2pt_renderer* r, const char* lut_blob_path)
{
    if (!r || !lut_blob_path) { b2_set_error("null argument"); return B2PT_ERR_INVALID_ARGUMENT; }
    B2_TRY(b2pt_ctx_make_current(r->ctx));
    B2_GUARDED(
    FILE* f = std::fopen(lut_blob_path, "rb");
    long file_bytes = 0;
    if (f)
    {
        std::fseek(f, 0, SEEK_END);
        file_bytes = std::ftell(f);
        std::fseek(f, 0, SEEK_SET);
    }
    char magic[8];
    uint32_t count = 0;
    if (!f || std::fread(magic, 1, 8, f) != 8 || std::memcmp(magic, "B2LUTS01", 8) != 0 || std::fread(&count, 4, 1, f) != 1)
    {
        if (f) std::fclose(f);
        b2_set_error("cannot read LUT container %s", lut_blob_path);
        return B2PT_ERR_IO;
    }
    for (void* h : r->lut_handles)
        b2pt_lut_destroy(r->ctx, h);
    r->lut_handles.clear();
    b2pt_bsdfs_data& b = r->loop.rd.bsdfs_data;
    for (uint32_t i = 0; i < count; i++)
    {
        char name[32];
        int32_t dims[3];
        if (std::fread(name, 1, 32, f) != 32 || std::fread(dims, 4, 3, f) != 3)
        {
            std::fclose(f);
            b2_set_error("truncated LUT container");
            return B2PT_ERR_IO;
        }
        if (dims[0] <= 0 || dims[1] <= 0 || dims[2] <= 0 || (unsigned long long)dims[0] * dims[1] * dims[2] * 4ull > (unsigned long long)file_bytes)
        {
            std::fclose(f);
            b2_set_error("LUT container: dimensions of '%.32s' exceed the file size", name);
            return B2PT_ERR_IO;
        }
        std::vector<float> texels((size_t)dims[0] * dims[1] * dims[2]);
        if (std::fread(texels.data(), 4, texels.size(), f) != texels.size())
        {
            std::fclose(f);
            b2_set_error("truncated LUT container");
            return B2PT_ERR_IO;
        }
        void* handle = nullptr;
        const bool is_sheen = !std::strcmp(name, "sheen_ltc"); /* stored as 96x32x1 floats = 32x32 RGB texels */
        int rc = is_sheen ? b2pt_lut_create(r->ctx, texels.data(), dims[0] / 3, dims[1], dims[2], 3, &handle)
                          : b2pt_lut_create(r->ctx, texels.data(), dims[0], dims[1], dims[2], 1, &handle);
        if (rc != B2PT_OK)
        {
            std::fclose(f);
            return rc;
        }
        r->lut_handles.push_back(handle);
        if (!std::strcmp(name, "ggx_conductor")) b.GGX_conductor_directional_albedo = handle;
        else if (!std::strcmp(name, "glossy_dielectric")) b.glossy_dielectric_directional_albedo = handle;
        else if (!std::strcmp(name, "ggx_glass")) b.GGX_glass_directional_albedo = handle;
        else if (!std::strcmp(name, "ggx_glass_inv")) b.GGX_glass_directional_albedo_inverse = handle;
        else if (!std::strcmp(name, "ggx_thin_glass")) b.GGX_thin_glass_directional_albedo = handle;
        else if (is_sheen) b.sheen_ltc_parameters_texture = handle;
    }
    std::fclose(f);
    return B2PT_OK;
    )
}

/* RendererEnvmap / CPURenderer::set_envmap (CPURenderer.cpp:467-508): luminance alias table (Image.cpp:536-545,
 * 652-671) + RGBE9995 packing (EnvmapRGBE9995.h:21-46, Packing.h:443-468) */
int b2pt_renderer_set_envmap_file(b2pt_renderer* r, const char* envmap_blob_path)
{
    if (!r || !envmap_blob_path) { b2_set_error("null argument"); return B2PT_ERR_INVALID_ARGUMENT; }
    B2_GUARDED(
    std::vector<float> rgba;
    int width = 0, height = 0;
    if (!b2_envmap_blob_read(envmap_blob_path, rgba, width, height))
    {
        b2_set_error("cannot read envmap blob %s", envmap_blob_path);
        return B2PT_ERR_IO;
    }
    return b2pt_renderer_set_envmap(r, rgba.data(), width, height);
    )
}

int b2pt_renderer_set_envmap_hdr_file(b2pt_renderer* r, const char* hdr_path)
{
    if (!r || !hdr_path) { b2_set_error("null argument"); return B2PT_ERR_INVALID_ARGUMENT; }
    B2_GUARDED(
    std::vector<float> rgba;
    int width = 0, height = 0;
    std::string error;
    if (!b2_hdr_read(hdr_path, /* flip_y = */ true, rgba, width, height, error))
    {
        b2_set_error("%s", error.c_str());
        return B2PT_ERR_IO;
    }
    return b2pt_renderer_set_envmap(r, rgba.data(), width, height);
    )
}

/* host-only helper behind b2pt_renderer_set_envmap_hdr_file, exported so that the decoder can be checked without a device:
 * out_rgba == null only reports the size; otherwise it receives width * height * 4 floats */
int b2pt_image_read_hdr(const char* hdr_path, int flip_y, float* out_rgba, int* out_width, int* out_height)
{
    if (!hdr_path || !out_width || !out_height) { b2_set_error("null argument"); return B2PT_ERR_INVALID_ARGUMENT; }
    B2_GUARDED(
    std::vector<float> rgba;
    std::string error;
    if (!b2_hdr_read(hdr_path, flip_y != 0, rgba, *out_width, *out_height, error))
    {
        b2_set_error("%s", error.c_str());
        return B2PT_ERR_IO;
    }
    if (out_rgba)
        std::memcpy(out_rgba, rgba.data(), rgba.size() * sizeof(float));
    return B2PT_OK;
    )
}

int b2pt_renderer_set_envmap(b2pt_renderer* r, const float* rgba, int width, int height)
{
    if (!r) { b2_set_error("null argument"); return B2PT_ERR_INVALID_ARGUMENT; }
    B2_TRY(b2pt_ctx_make_current(r->ctx));
    B2_GUARDED(
    free_all(r, r->envmap_allocs);
    b2pt_world_settings& w = r->loop.rd.world_settings;
    if (!rgba || width <= 0 || height <= 0)
    {
        w.ambient_light_type = 1;
        w.envmap = nullptr;
        return B2PT_OK;
    }
    B2HostEnvmap env;
    b2_host_envmap_prepare(env, rgba, width, height);
    B2_TRY(upload(r, r->envmap_allocs, env.packed, &w.envmap));
    B2_TRY(upload(r, r->envmap_allocs, env.alias_probas, &w.envmap_alias_table.alias_table_probas));
    B2_TRY(upload(r, r->envmap_allocs, env.alias_alias, &w.envmap_alias_table.alias_table_alias));
    w.envmap_alias_table.sum_elements = env.total_luminance;
    w.envmap_alias_table.size = (unsigned)((size_t)width * height);
    w.envmap_total_sum = env.total_luminance;
    w.envmap_width = (unsigned)width;
    w.envmap_height = (unsigned)height;
    w.ambient_light_type = 2;
    return B2PT_OK;
    )
}

b2pt_render_data* b2pt_renderer_render_data(b2pt_renderer* r) { return r ? &r->loop.rd : nullptr; }

int b2pt_renderer_set_host_seed(b2pt_renderer* r, unsigned int seed)
{
    if (!r) { b2_set_error("null argument"); return B2PT_ERR_INVALID_ARGUMENT; }
    r->loop.host_rng = make_rng(seed);
    return B2PT_OK;
}

int b2pt_renderer_set_counters(b2pt_renderer* r, int enabled)
{
    if (!r) { b2_set_error("null argument"); return B2PT_ERR_INVALID_ARGUMENT; }
    r->counters = enabled ? 1 : 0;
    r->kernels_stale = true;
    return B2PT_OK;
}

/* GPURenderer::get_global_compiler_options()->set_macro_value(name, value) + recompile_kernels() (GPURenderer.cpp:482-501): the kernels
 * are looked up again before the next render. The macros that decide which passes and buffers a renderer has are its variant's and are
 * refused here; whether the value is served is answered at once (b2pt_kernel_query on the MegaKernel and, for ReSTIR DI, its passes). */
int b2pt_renderer_set_macro(b2pt_renderer* r, const char* name, int value)
{
    if (!r || !name) { b2_set_error("null argument"); return B2PT_ERR_INVALID_ARGUMENT; }
    for (const char* fixed : { "BSDFOverride", "DirectLightSamplingStrategy", "DirectLightSamplingBaseStrategy", "ReSTIR_DI_DoLightPresampling", "B2PT_PRECISE", "B2PT_COUNTERS", "B2PT_ALPHA_TEST" })
        if (!std::strcmp(name, fixed))
        {
            b2_set_error("b2pt_renderer_set_macro: %s belongs to the renderer's variant / flavour (b2pt_renderer_create, b2pt_renderer_set_counters)", name);
            return B2PT_ERR_INVALID_ARGUMENT;
        }
    const VariantKey* v = find_variant(r->variant);
    std::vector<std::pair<std::string, int>> macros = r->macros;
    bool replaced = false;
    for (auto& macro : macros)
        if (macro.first == name)
        {
            macro.second = value;
            replaced = true;
        }
    if (!replaced)
        macros.emplace_back(name, value);
    std::vector<b2pt_option> opts = { { "BSDFOverride", v->bsdf }, { "DirectLightSamplingStrategy", v->direct }, { "DirectLightSamplingBaseStrategy", v->base },
                                      { "ReSTIR_DI_LaterBouncesSamplingStrategy", v->later } };
    for (const auto& macro : macros)
        opts.push_back(b2pt_option{ macro.first.c_str(), macro.second });
    for (int pass = 0; pass < B2_PASS_COUNT; pass++)
    {
        const char* entry = entry_name_of_pass(pass);
        bool restir_pass = pass != B2_PASS_CAMERA_RAYS && pass != B2_PASS_MEGAKERNEL;
        if (!entry || (restir_pass && !v->restir) || (pass == B2_PASS_RESTIR_DI_PRESAMPLING && !v->presampling))
            continue;
        B2_TRY(b2pt_kernel_query(entry, opts.data(), (int)opts.size()));
    }
    r->macros = macros;
    r->kernels_stale = true;
    return B2PT_OK;
}

/* Host-only views of the variant table: tests walk it without a device (every pass of every named variant must resolve in both flavours) */
int b2pt_renderer_variant_count(void) { return (int)(sizeof(kVariants) / sizeof(kVariants[0])); }

const char* b2pt_renderer_variant_name(int index)
{
    return index >= 0 && index < b2pt_renderer_variant_count() ? kVariants[index].name : nullptr;
}

int b2pt_renderer_variant_query(const char* variant, int precise)
{
    const VariantKey* v = variant ? find_variant(variant) : nullptr;
    if (!v) { b2_set_error("unknown variant '%s'", variant ? variant : "(null)"); return B2PT_ERR_NO_VARIANT; }
    for (int counters = 0; counters < 2; counters++)
        for (int pass = 0; pass < B2_PASS_COUNT; pass++)
        {
            if (!variant_has_pass(v, pass))
                continue;
            std::vector<b2pt_option> opts = variant_options(v, precise, counters, 0);
            B2_TRY(b2pt_kernel_query(entry_name_of_pass(pass), opts.data(), (int)opts.size()));
        }
    return B2PT_OK;
}

int b2pt_renderer_set_profile_passes(b2pt_renderer* r, int enabled)
{
    if (!r) { b2_set_error("null argument"); return B2PT_ERR_INVALID_ARGUMENT; }
    r->profile_passes = enabled != 0;
    return B2PT_OK;
}

/* GPURenderer::reset (sample_number = 0, need_to_reset = true) + every render pass's reset(): ReSTIRDIRenderPass::reset
 * empties the three reservoir buffers and restarts the ping-pong (src/Renderer/RenderPasses/ReSTIRDIRenderPass.cpp:319-341),
 * so the first frame after a reset has no temporal / spatial history, exactly like the first frame ever. */
int b2pt_renderer_reset(b2pt_renderer* r)
{
    if (!r) { b2_set_error("null argument"); return B2PT_ERR_INVALID_ARGUMENT; }
    r->loop.rd.render_settings.need_to_reset = 1;
    r->loop.rd.render_settings.sample_number = 0;
    if (r->loop.restir_di)
    {
        B2_TRY(b2pt_ctx_make_current(r->ctx));
        for (int i = 0; i < 3; i++)
            if (r->d_reservoirs[i])
                B2_TRY(clear_reservoirs(r, r->d_reservoirs[i], r->stream));
        r->loop.odd_frame = false;
        r->loop.rd.render_settings.restir_di_settings.initial_candidates.output_reservoirs = r->d_reservoirs[0];
        r->loop.rd.render_settings.restir_di_settings.restir_output_reservoirs = r->d_reservoirs[1];
    }
    return B2PT_OK;
}

int b2pt_renderer_render(b2pt_renderer* r, int samples)
{
    if (!r || samples < 0) { b2_set_error("invalid argument"); return B2PT_ERR_INVALID_ARGUMENT; }
    if (!r->has_scene) { b2_set_error("b2pt_renderer_render: no scene set"); return B2PT_ERR_INVALID_ARGUMENT; }
    B2_TRY(b2pt_ctx_make_current(r->ctx));
    B2_TRY(resolve_kernels(r));
    if (r->world > 1 && r->loop.restir_di && r->loop.rd.render_settings.enable_adaptive_sampling)
    {
        /* CameraRays renders the halo rows around a rank's bands too, and with adaptive sampling it reads the sample counts and the
         * accumulated colours of those pixels (AdaptiveSampling.h:30-102), which only their owner keeps up to date: the G-buffers of
         * the ranks would drift apart. Refused rather than rendered differently from one GPU. */
        b2_set_error("b2pt_renderer_render: adaptive sampling is not served for ReSTIR DI on a partitioned frame");
        return B2PT_ERR_UNSUPPORTED_SCENE;
    }
    r->loop.camera.sensor_width = r->width;
    r->loop.camera.sensor_height = r->height;
    if (r->loop.restir_di)
    {
        /* With the fused pass on and the spatial pass off the reference's own pass order (CPURenderer.cpp:1018-1032) shades from the temporal
         * pass's output buffer, which the fused path leaves null: its CPU renderer crashes there (tools/settings_fuzz.py), a GPU would
         * fault. Refused instead. */
        const b2pt_restir_di_settings& di = r->loop.rd.render_settings.restir_di_settings;
        if (di.do_fused_spatiotemporal && !di.common_spatial_pass.do_spatial_reuse_pass)
        {
            b2_set_error("b2pt_renderer_render: restir_di_do_fused_spatiotemporal needs restir_di_do_spatial_reuse_pass (the reference shades from a null buffer otherwise)");
            return B2PT_ERR_INVALID_ARGUMENT;
        }
    }

    int rc = B2PT_OK;
    auto launch = [&](B2Pass pass) {
        if (rc != B2PT_OK)
            return;
        if (pass == B2_PASS_CAMERA_RAYS)
        {
            /* pre_render_update: clear still_one_ray_active and the converged-pixel counter (CPURenderer.cpp:574-584) */
            rc = b2pt_memset(r->ctx, r->d_status, 0, 16, r->stream);
            if (rc != B2PT_OK)
                return;
        }
        b2pt_kernel* k = r->kernels[pass];
        if (!k)
        {
            b2_set_error("pass %d has no kernel in variant %s", (int)pass, r->variant.c_str());
            rc = B2PT_ERR_NO_VARIANT;
            return;
        }
        const int* launch_rows = nullptr;
        int launch_row_count = 0;
        if (r->world > 1 && r->loop.restir_di)
        {
            const b2pt_restir_di_settings& di = r->loop.rd.render_settings.restir_di_settings;
            /* The reuse passes read reservoirs AND G-buffer entries of neighbouring pixels, which may belong to another GPU.
             * G-buffer: CameraRays also renders the halo rows around this rank's bands (same seeds => same bytes as on their
             * owner; cheaper than exchanging 132 B/px). Reservoirs: the halo rows of the buffer a pass reads around a pixel are
             * fetched from their owners (36 B/px, NCCL send / recv on this stream) right before that pass. */
            rc = update_partition_rows(r);
            if (rc != B2PT_OK)
                return;
            if (pass == B2_PASS_CAMERA_RAYS)
            {
                launch_rows = r->d_camera_rows;
                launch_row_count = r->camera_row_count;
            }
            void* needs = nullptr;
            if (pass == B2_PASS_RESTIR_DI_SPATIOTEMPORAL || pass == B2_PASS_RESTIR_DI_TEMPORAL)
                needs = di.temporal_pass.input_reservoirs; /* reprojection + its neighbour search, and the fused pass's spatial neighbours */
            else if (pass == B2_PASS_RESTIR_DI_SPATIAL)
                needs = di.spatial_pass.input_reservoirs;
            if (needs)
            {
                if (!r->comm)
                {
                    b2_set_error("ReSTIR DI on %d GPUs needs a communicator (b2pt_renderer_set_partition)", r->world);
                    rc = B2PT_ERR_INVALID_ARGUMENT;
                    return;
                }
                rc = b2pt_comm_exchange_halo(r->comm, needs, 9, r->pixels() * 4, (size_t)r->width * 4, r->height, r->band_rows, r->halo_rows, r->stream);
                if (rc != B2PT_OK)
                    return;
            }
        }
        /* LightPresamplingParameters as the reference's pass fills it (ReSTIRDIRenderPass.cpp:417-444) */
        struct { int number_of_subsets, subset_size; void* out_light_samples; float envmap_sampling_probability; } presampling;
        const b2pt_restir_di_settings& di_settings = r->loop.rd.render_settings.restir_di_settings;
        presampling.number_of_subsets = di_settings.light_presampling.number_of_subsets;
        presampling.subset_size = di_settings.light_presampling.subset_size;
        presampling.out_light_samples = di_settings.light_presampling.light_samples;
        presampling.envmap_sampling_probability = di_settings.initial_candidates.envmap_candidate_probability;
        void* args[4] = { &r->loop.rd, nullptr, nullptr, nullptr };
        unsigned int* masks_u = nullptr;
        unsigned long long* masks_ull = r->loop.rd.render_settings.restir_di_settings.common_spatial_pass.per_pixel_spatial_reuse_directions_mask_ull;
        uint8_t* radius = r->loop.rd.render_settings.restir_di_settings.common_spatial_pass.per_pixel_spatial_reuse_radius;
        if (pass == B2_PASS_RESTIR_DI_PRESAMPLING)
        {
            args[0] = &presampling;
            args[1] = &r->loop.rd;
        }
        else if (pass == B2_PASS_RESTIR_DIRECTIONAL_REUSE)
        {
            args[1] = &masks_u;
            args[2] = &masks_ull;
            args[3] = &radius;
        }
        rc = b2pt_kernel_launch_partition(k, r->width, r->height, r->rank, r->world, r->band_rows, launch_rows, launch_row_count, args, r->stream);
        r->pass_launches[pass]++;
        if (rc == B2PT_OK && r->profile_passes)
        {
            float ms = 0.0f;
            rc = b2pt_kernel_last_ms(k, &ms);
            r->pass_ms[pass] += ms;
        }
    };

    B2_TRY(b2pt_event_record(r->ctx, r->ev_start, r->stream));
    for (int s = 0; s < samples && rc == B2PT_OK; s++)
        r->loop.render_one_sample(launch);
    B2_TRY(rc);
    B2_TRY(b2pt_event_record(r->ctx, r->ev_stop, r->stream));
    r->timed = true;
    return B2PT_OK;
}

/* Multi-GPU partition of the frame: interleaved row bands, scene replicated. */
/* Band height of a partition: interleaved bands of `preferred` rows leave the ranks with unequal shares whenever the number of bands
 * is not a multiple of the number of ranks (2160 rows in 32-row bands over 8 GPUs: 9 bands = 288 rows on four ranks, 8 = 256 on the
 * others, and a step takes as long as its slowest rank). The height between `minimum` and `preferred` whose busiest rank owns the
 * fewest rows is taken (2160 rows, 8 GPUs: 30-row bands, 270 rows everywhere); ties go to the taller band (fewer halo rows). */
static int balanced_band_rows(int height, int world, int preferred, int minimum)
{
    if (world <= 1)
        return preferred;
    int best = preferred, best_rows = height + 1;
    for (int band = preferred; band >= minimum; band--)
    {
        int worst = 0;
        for (int rank = 0; rank < world; rank++)
        {
            int rows = 0;
            for (int b = rank; b * band < height; b += world)
                rows += std::min(band, height - b * band);
            worst = std::max(worst, rows);
        }
        if (worst < best_rows)
        {
            best_rows = worst;
            best = band;
        }
    }
    return best;
}

int b2pt_renderer_set_partition(b2pt_renderer* r, int rank, int world, b2pt_comm* comm, int dynamic_camera)
{
    if (!r || world < 1 || rank < 0 || rank >= world) { b2_set_error("b2pt_renderer_set_partition: invalid argument"); return B2PT_ERR_INVALID_ARGUMENT; }
    if (comm && (b2_comm_rank(comm) != rank || b2_comm_world(comm) != world))
    {
        b2_set_error("b2pt_renderer_set_partition: the communicator is rank %d of %d, not %d of %d", b2_comm_rank(comm), b2_comm_world(comm), rank, world);
        return B2PT_ERR_INVALID_ARGUMENT;
    }
    B2_TRY(b2pt_ctx_make_current(r->ctx));
    B2_TRY(b2pt_stream_synchronize(r->ctx, r->stream));
    r->rank = rank;
    r->world = world;
    r->comm = comm;
    r->dynamic_camera = dynamic_camera != 0;
    r->band_rows = balanced_band_rows(r->height, world, r->loop.restir_di ? 32 : B2PT_BAND_ROWS, r->loop.restir_di ? 24 : 8);
    if (r->d_camera_rows)
    {
        b2pt_free(r->ctx, r->d_camera_rows);
        r->d_camera_rows = nullptr;
    }
    r->camera_row_count = 0;
    r->halo_rows = 0;
    r->owned_row_count = (int)b2_partition_rows(r->height, rank, world, r->band_rows, 0).size();
    return B2PT_OK;
}

int b2pt_renderer_gather_framebuffer(b2pt_renderer* r)
{
    if (!r) { b2_set_error("null argument"); return B2PT_ERR_INVALID_ARGUMENT; }
    if (r->world == 1)
        return B2PT_OK;
    if (!r->comm) { b2_set_error("b2pt_renderer_gather_framebuffer: no communicator (b2pt_renderer_set_partition)"); return B2PT_ERR_INVALID_ARGUMENT; }
    B2_TRY(b2pt_ctx_make_current(r->ctx));
    B2_TRY(b2pt_comm_allgather_rows(r->comm, r->loop.rd.buffers.accumulated_ray_colors, 1, r->pixels() * 12, (size_t)r->width * 12, r->height, r->band_rows, r->stream));
    /* the step's device time = render + gather: the stop event moves behind the gather */
    return b2pt_event_record(r->ctx, r->ev_stop, r->stream);
}

int b2pt_renderer_partition_info(b2pt_renderer* r, uint64_t* out_camera_pixels, uint64_t* out_owned_pixels, int* out_band_rows, int* out_halo_rows)
{
    if (!r) { b2_set_error("null argument"); return B2PT_ERR_INVALID_ARGUMENT; }
    const bool halo_partition = r->world > 1 && r->loop.restir_di;
    if (halo_partition)
        B2_TRY(update_partition_rows(r));
    const int owned = r->world > 1 ? r->owned_row_count : r->height;
    if (out_owned_pixels) *out_owned_pixels = (uint64_t)owned * r->width;
    if (out_camera_pixels) *out_camera_pixels = (uint64_t)(halo_partition ? r->camera_row_count : owned) * r->width;
    if (out_band_rows) *out_band_rows = r->band_rows;
    if (out_halo_rows) *out_halo_rows = halo_partition ? r->halo_rows : 0;
    return B2PT_OK;
}

int b2pt_renderer_set_camera(b2pt_renderer* r, const float* camera52)
{
    if (!r || !camera52) { b2_set_error("null argument"); return B2PT_ERR_INVALID_ARGUMENT; }
    B2SceneCamera c;
    std::memcpy(&c, camera52, sizeof(c));
    b2pt_camera& cam = r->loop.camera;
    std::memcpy(cam.inverse_view, c.inverse_view, 64);
    std::memcpy(cam.inverse_projection, c.inverse_projection, 64);
    std::memcpy(cam.view_projection, c.view_projection, 64);
    cam.position = b2pt_float3{ c.position[0], c.position[1], c.position[2] };
    cam.vertical_fov = c.vertical_fov;
    return B2PT_OK;
}

int b2pt_render_data_framebuffer_offset(void)
{
    return (int)offsetof(b2pt_render_data, buffers.accumulated_ray_colors);
}

int b2pt_renderer_synchronize(b2pt_renderer* r)
{
    if (!r) { b2_set_error("null argument"); return B2PT_ERR_INVALID_ARGUMENT; }
    B2_TRY(b2pt_ctx_make_current(r->ctx));
    return b2pt_stream_synchronize(r->ctx, r->stream);
}

int b2pt_renderer_last_render_ms(b2pt_renderer* r, float* out_ms)
{
    if (!r || !out_ms) { b2_set_error("null argument"); return B2PT_ERR_INVALID_ARGUMENT; }
    *out_ms = 0.0f;
    if (!r->timed)
        return B2PT_OK;
    B2_TRY(b2pt_event_synchronize(r->ctx, r->ev_stop));
    return b2pt_event_elapsed_ms(r->ctx, r->ev_start, r->ev_stop, out_ms);
}

int b2pt_renderer_pass_ms(b2pt_renderer* r, int pass, float* out_ms, uint64_t* out_launches)
{
    if (!r || pass < 0 || pass >= B2_PASS_COUNT) { b2_set_error("invalid argument"); return B2PT_ERR_INVALID_ARGUMENT; }
    if (out_ms) *out_ms = (float)r->pass_ms[pass];
    if (out_launches) *out_launches = r->pass_launches[pass];
    return B2PT_OK;
}

int b2pt_renderer_download_framebuffer(b2pt_renderer* r, float* out_rgb)
{
    if (!r || !out_rgb) { b2_set_error("null argument"); return B2PT_ERR_INVALID_ARGUMENT; }
    B2_TRY(b2pt_ctx_make_current(r->ctx));
    B2_TRY(b2pt_memcpy_d2h(r->ctx, out_rgb, r->loop.rd.buffers.accumulated_ray_colors, r->pixels() * 12, r->stream));
    return b2pt_stream_synchronize(r->ctx, r->stream);
}

int b2pt_renderer_download_gbuffer(b2pt_renderer* r, int* out_prim, float* out_position, uint32_t* out_geo, uint32_t* out_shade, uint32_t* out_material27)
{
    if (!r) { b2_set_error("null argument"); return B2PT_ERR_INVALID_ARGUMENT; }
    B2_TRY(b2pt_ctx_make_current(r->ctx));
    const b2pt_gbuffer& g = r->loop.rd.g_buffer;
    const size_t n = r->pixels();
    B2_TRY(b2pt_stream_synchronize(r->ctx, r->stream));
    if (out_prim) B2_TRY(b2pt_memcpy_d2h(r->ctx, out_prim, g.first_hit_prim_index, n * 4, nullptr));
    if (out_position) B2_TRY(b2pt_memcpy_d2h(r->ctx, out_position, g.primary_hit_position, n * 12, nullptr));
    if (out_geo) B2_TRY(b2pt_memcpy_d2h(r->ctx, out_geo, g.geometric_normals, n * 4, nullptr));
    if (out_shade) B2_TRY(b2pt_memcpy_d2h(r->ctx, out_shade, g.shading_normals, n * 4, nullptr));
    if (out_material27) B2_TRY(b2pt_memcpy_d2h(r->ctx, out_material27, g.materials, n * 108, nullptr));
    return B2PT_OK;
}

int b2pt_renderer_download_reservoirs(b2pt_renderer* r, int which, uint32_t* out_words9)
{
    if (!r || !out_words9 || which < 0 || which > 3) { b2_set_error("invalid argument"); return B2PT_ERR_INVALID_ARGUMENT; }
    if (!r->loop.restir_di) { b2_set_error("variant %s has no ReSTIR DI reservoirs", r->variant.c_str()); return B2PT_ERR_INVALID_ARGUMENT; }
    B2_TRY(b2pt_ctx_make_current(r->ctx));
    B2_TRY(b2pt_stream_synchronize(r->ctx, r->stream));
    const void* src = which < 3 ? r->d_reservoirs[which] : r->loop.rd.render_settings.restir_di_settings.restir_output_reservoirs;
    return b2pt_memcpy_d2h(r->ctx, out_words9, src, r->pixels() * 36, nullptr);
}

int b2pt_renderer_scene_info(b2pt_renderer* r, b2pt_bvh_stats_t* out_bvh, int* out_triangles, int* out_emissive_triangles)
{
    if (!r || !r->has_scene) { b2_set_error("no scene"); return B2PT_ERR_INVALID_ARGUMENT; }
    if (out_bvh) B2_TRY(b2pt_bvh_stats(r->bvh, out_bvh));
    if (out_triangles) *out_triangles = (int)r->scene.blob.n_triangles();
    if (out_emissive_triangles) *out_emissive_triangles = (int)r->scene.emissive_triangles.size();
    return B2PT_OK;
}

int b2pt_renderer_download_derived(b2pt_renderer* r, uint32_t* out_packed33, float* out_areas, int* out_emissive, float* out_probas, int* out_alias)
{
    if (!r || !r->has_scene) { b2_set_error("no scene"); return B2PT_ERR_INVALID_ARGUMENT; }
    const B2HostScene& s = r->scene;
    if (out_packed33) std::memcpy(out_packed33, s.packed_materials.data(), s.packed_materials.size() * sizeof(B2PackedTexturedMaterial));
    if (out_areas) std::memcpy(out_areas, s.triangle_areas.data(), s.triangle_areas.size() * 4);
    if (out_emissive) std::memcpy(out_emissive, s.emissive_triangles.data(), s.emissive_triangles.size() * 4);
    if (out_probas) std::memcpy(out_probas, s.power_alias_probas.data(), s.power_alias_probas.size() * 4);
    if (out_alias) std::memcpy(out_alias, s.power_alias_alias.data(), s.power_alias_alias.size() * 4);
    return B2PT_OK;
}

} /* extern "C" */
