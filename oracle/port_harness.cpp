/*
 * oracle/port_harness.cpp -- TEST / DEBUG INFRASTRUCTURE ONLY ("port" oracle). Never part of the product path.
 *
 * Compiles the product's OWN per-pixel headers (hiprt-path-tracer_b200/csrc/b2_*.h -- the code the CUDA kernels
 * instantiate) for the host and runs them pixel by pixel under OpenMP, with the same command line and the same
 * dump files as oracle/ref_harness.cpp (the real reference). Purpose: (1) in a container without a GPU, diff the
 * B200 shading / traversal code against the reference per pixel and per pass; (2) on the GPU box, tell apart
 * "CUDA codegen / libdevice rounding" differences from logic differences (port vs CUDA must agree to the last
 * transcendental ulp, port vs reference must agree within the parity tolerances).
 * The authoritative oracle is oracle/_ref (reference sources); this one is only as good as its agreement with it,
 * which tests/test_oracle.py checks (test_port_oracle_matches_golden and following).
 * --staged runs the MegaKernel through the stage functions the GPU kernels use (see staged_megakernel_pixel below).
 */
#include "b2_frame_loop.h"
#include "b2_host_scene.h"
#include "b2_kernels_body.h"

#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <omp.h>
#include <string>

struct Args
{
    std::string scene, luts, envmap, dump, variant = "c1", cameras;
    int debug_x = -1, debug_y = -1, debug_frame = -1; /* --debug-visibility X Y FRAME */
    bool trace_ray = false;                           /* --trace-ray ox oy oz dx dy dz t_max skip_prim: one ray through the product's traversal */
    float ray[7] = { 0, 0, 0, 0, 0, 1, 3.0e38f };
    int ray_skip_prim = -1;
    int width = 512, height = 512, spp = 1, bounces = 3;
    int win_x0 = 0, win_y0 = 0, win_x1 = -1, win_y1 = -1;
    int threads = 0;
    bool fused = true, dump_per_sample = false, no_jitter = false, directional_reuse = false, kat = false, confidence_weights = true, staged = false, spatial_pass = true;
    int spatial_passes = 1, neighbors = -1;
    std::vector<std::string> pokes; /* --poke OFFSET:KIND:VALUE, as in ref_harness.cpp */
    unsigned int seed = 42;
};

static Args parse_args(int argc, char** argv)
{
    Args a;
    for (int i = 1; i < argc; i++)
    {
        std::string k = argv[i];
        auto next = [&]() -> const char* { if (i + 1 >= argc) { std::fprintf(stderr, "missing value for %s\n", k.c_str()); std::exit(2); } return argv[++i]; };
        if (k == "--scene") a.scene = next();
        else if (k == "--luts") a.luts = next();
        else if (k == "--envmap") a.envmap = next();
        else if (k == "--dump") a.dump = next();
        else if (k == "--variant") a.variant = next();
        else if (k == "--width") a.width = std::atoi(next());
        else if (k == "--height") a.height = std::atoi(next());
        else if (k == "--spp") a.spp = std::atoi(next());
        else if (k == "--bounces") a.bounces = std::atoi(next());
        else if (k == "--threads") a.threads = std::atoi(next());
        else if (k == "--seed") a.seed = (unsigned)std::strtoul(next(), nullptr, 10);
        else if (k == "--window") { a.win_x0 = std::atoi(next()); a.win_y0 = std::atoi(next()); a.win_x1 = std::atoi(next()); a.win_y1 = std::atoi(next()); }
        else if (k == "--no-fused") a.fused = false;
        else if (k == "--spatial-passes") a.spatial_passes = std::atoi(next());
        else if (k == "--dump-per-sample") a.dump_per_sample = true;
        else if (k == "--no-jitter") a.no_jitter = true;
        else if (k == "--kat") a.kat = true; /* function-level known-answer dump instead of a render (see run_bsdf_kat) */
        else if (k == "--poke") a.pokes.push_back(next());
        else if (k == "--neighbors") a.neighbors = std::atoi(next());
        else if (k == "--trace-ray") { a.trace_ray = true; for (int j = 0; j < 7; j++) a.ray[j] = std::strtof(next(), nullptr); a.ray_skip_prim = std::atoi(next()); }
        else if (k == "--debug-visibility") { a.debug_x = std::atoi(next()); a.debug_y = std::atoi(next()); a.debug_frame = std::atoi(next()); }
        else if (k == "--cameras") a.cameras = next(); /* a moving camera: one B2SceneCamera (52 floats) per frame, as ref_extras_* takes it */
        else if (k == "--staged") a.staged = true; /* the MegaKernel through megakernel_stage1 / stage2, the way the GPU kernels run it */
        else if (k == "--no-confidence-weights") a.confidence_weights = false;
        else if (k == "--no-spatial") a.spatial_pass = false;
        else if (k == "--directional-reuse") a.directional_reuse = true; /* use_adaptive_directional_spatial_reuse + the prepass */
        else { std::fprintf(stderr, "unknown argument %s\n", k.c_str()); std::exit(2); }
    }
    if (a.win_x1 < 0) { a.win_x1 = a.width; a.win_y1 = a.height; }
    return a;
}

template <typename T>
static void dump_array(const std::string& prefix, const char* name, const T* data, size_t count)
{
    if (prefix.empty()) return;
    std::string path = prefix + name + ".bin";
    FILE* f = std::fopen(path.c_str(), "wb");
    if (!f) { std::fprintf(stderr, "cannot write %s\n", path.c_str()); std::exit(1); }
    std::fwrite(data, sizeof(T), count, f);
    std::fclose(f);
}

struct LutStore
{
    std::map<std::string, std::vector<float>> data;
    std::map<std::string, b2pt_lut> luts;
    bool load(const char* path)
    {
        FILE* f = std::fopen(path, "rb");
        if (!f) return false;
        char magic[8];
        uint32_t count = 0;
        if (std::fread(magic, 1, 8, f) != 8 || std::memcmp(magic, "B2LUTS01", 8) != 0 || std::fread(&count, 4, 1, f) != 1) { std::fclose(f); return false; }
        for (uint32_t i = 0; i < count; i++)
        {
            char name[32];
            int32_t dims[3];
            if (std::fread(name, 1, 32, f) != 32 || std::fread(dims, 4, 3, f) != 3) { std::fclose(f); return false; }
            std::vector<float>& v = data[name];
            v.resize((size_t)dims[0] * dims[1] * dims[2]);
            if (std::fread(v.data(), 4, v.size(), f) != v.size()) { std::fclose(f); return false; }
            if (!std::strcmp(name, "sheen_ltc"))
                luts[name] = b2pt_lut{ v.data(), dims[0] / 3, dims[1], dims[2], 3 };
            else
                luts[name] = b2pt_lut{ v.data(), dims[0], dims[1], dims[2], 1 };
        }
        std::fclose(f);
        return true;
    }
};

static unsigned int g_port_faults[B2_FAULT_COUNT] = { 0, 0, 0 };
static unsigned long long g_port_counters[3] = { 0, 0, 0 }; /* rays, nodes, triangles: traversal steps per ray of the CW8 BVH */

/* "c4b<N>" = c4 with ReSTIR_DI_BiasCorrectionWeights = N (the renderer's variant table, csrc/b2_renderer.cpp) */
/* c4x<N>: combinations (the table of b2_renderer.cpp / oracle/Makefile): c4x1 = weights 3, options 2 | 4, later MIS; c4x2 = weights 6, options 1 | 8, later BSDF */
static int bias_correction_of(const std::string& variant)
{
    if (variant == "c4x1") return 3;
    if (variant == "c4x2") return 6;
    return variant.rfind("c4b", 0) == 0 ? std::atoi(variant.c_str() + 3) : 4;
}
/* c4v<N>: one of the four ReSTIR DI visibility macros away from its default (bits B2_DI_OPT_*, the table of b2_renderer.cpp) */
static int restir_options_of(const std::string& variant)
{
    if (variant == "c4v1") return 1 | 4 | 8;
    if (variant == "c4v2") return 2 | 4 | 8;
    if (variant == "c4v3") return 8;
    if (variant == "c4v4") return 4;
    if (variant == "c4v5") return 1 | 2 | 8;
    if (variant == "c4o") return 4 | 8 | 16; /* ReSTIR_DI_DoOptimalVisibilitySampling */
    if (variant == "c4x1") return 2 | 4;
    if (variant == "c4x2") return 1 | 8;
    return 4 | 8;
}
static bool is_c4(const std::string& variant) { return variant == "c4" || variant == "c4p" || variant == "c4pu" || variant == "c4o" || variant.rfind("c4b", 0) == 0 || variant.rfind("c4v", 0) == 0 || variant.rfind("c4l", 0) == 0 || variant.rfind("c4x", 0) == 0; }
static bool g_scene_has_alpha_tested_materials = false; /* = what the renderer passes as B2PT_ALPHA_TEST */

/* The MegaKernel the way the phased and wavefront GPU kernels run it (b2_kernels.cu: megakernel_phased, b2_wavefront.h): one path
 * carried through begin -> { trace A, stage 1, trace B, stage 2 } with the rays the stages request traced between them, instead of
 * megakernel_bounce's straight-line code. Same protocol as the device loop: unrequested rays answer "miss", deferred Y only where the
 * variant has it, the sequential ray only with RIS. */
static int g_debug_rays_x = -1, g_debug_rays_y = -1; /* --staged --debug-visibility X Y FRAME: every ray of that pixel's path, with a brute-force check */

template <class V>
static void staged_megakernel_pixel(const Ctx& ctx, uint32_t x, uint32_t y)
{
    const b2pt_render_data& rd = *ctx.rd;
    if (x >= (uint32_t)rd.render_settings.render_resolution.x || y >= (uint32_t)rd.render_settings.render_resolution.y)
        return;
    const uint32_t pixel_index = x + y * (uint32_t)rd.render_settings.render_resolution.x;
    if (!rd.aux_buffers.pixel_active[pixel_index])
        return;
    if (rd.g_buffer.first_hit_prim_index[pixel_index] == -1)
    {
        megakernel_primary_miss<V>(ctx, x, y, pixel_index);
        return;
    }
    auto trace = [&](const RayRequest& ray) {
        B2Hit none;
        none.t = -1.0f;
        none.u = none.v = 0.0f;
        none.prim = -1;
        none.ng = v3(0.0f);
        if (ray.kind == B2_RAY_NONE)
            return none;
        B2Hit result = b2_trace(ctx, ray.kind == B2_RAY_ANY_HIT, ray.origin, ray.direction, ray.t_max, ray.skip_prim, ray.main_path != 0);
        if ((int)x == g_debug_rays_x && (int)y == g_debug_rays_y)
        {
            int brute_prim = -1;
            float brute_t = 3.0e38f;
            for (uint32_t i = 0; i < ctx.bvh.tri_count; i++)
            {
                const B2Tri& tri = ctx.bvh.tris[i];
                float t, u, v;
                if (tri.prim != ray.skip_prim && b2_intersect_triangle(ray.origin, ray.direction, v3(tri.ax, tri.ay, tri.az), v3(tri.e1x, tri.e1y, tri.e1z), v3(tri.e2x, tri.e2y, tri.e2z), t, u, v) && t < brute_t)
                {
                    brute_t = t;
                    brute_prim = tri.prim;
                }
            }
            std::fprintf(stderr, "ray %s o %.9g %.9g %.9g d %.9g %.9g %.9g tmax %.9g skip %d -> prim %d t %.9g | brute force nearest prim %d t %.9g\n", ray.kind == B2_RAY_ANY_HIT ? "shadow " : "closest",
                         ray.origin.x, ray.origin.y, ray.origin.z, ray.direction.x, ray.direction.y, ray.direction.z, ray.t_max, ray.skip_prim, result.prim, result.t, brute_prim, brute_t);
        }
        return result;
    };
    PathState s;
    BouncePhase<V> ph;
    bool alive = megakernel_begin<V>(ctx, x, y, s);
    ph.reset();
    RayRequest rays[5];
    for (int k = 0; k < 5; k++)
        request_none(rays[k]);
    while (alive)
    {
        B2Hit results[5];
        for (int k = 0; k < 5; k++)
            results[k] = trace(rays[k]);
        RayRequest sequential;
        request_none(sequential);
        if (!megakernel_stage1<V>(ctx, s, ph, s.payload.bounce > 0 ? &results[0] : nullptr, results[1], results[2], results[3], results[4], sequential))
        {
            megakernel_end(ctx, s);
            break;
        }
        RayRequest none;
        request_none(none);
        B2Hit sequential_result = trace(Nee<V>::HAS_SEQUENTIAL_RAY ? sequential : none);
        for (int k = 0; k < 5; k++)
            request_none(rays[k]);
        if (!megakernel_stage2<V>(ctx, s, ph, sequential_result, rays[0], rays[1], rays[2], rays[3], rays[4]))
        {
            megakernel_end(ctx, s);
            break;
        }
        if (!Nee<V>::HAS_DEFERRED_Y)
            request_none(rays[2]);
    }
}

static int g_current_frame = 0;
static bool frame_of_debug(const Args& a) { return a.debug_x >= 0 && g_current_frame == a.debug_frame; }

template <class V>
static void run_pass(B2Pass pass, const b2pt_render_data& rd, const B2BvhDevice& bvh, const Args& a)
{
#pragma omp parallel
    {
        uint64_t stack[B2_STACK_ENTRIES];
        Ctx ctx;
        ctx.rd = &rd;
        ctx.bvh = bvh;
        ctx.stack = stack;
        ctx.stack_stride = 1;
        B2TraversalCounters local;
        local.rays = local.nodes = local.tris = 0;
        ctx.counters = &local;
        ctx.faults = g_port_faults;
        ctx.light_presampling = a.variant == "c4p" ? 1 : a.variant == "c4pu" ? 2 : 0; /* c4pu: ReSTIR_DI_LightPresamplingStrategy = LSS_BASE_UNIFORM */
        ctx.bias_correction = bias_correction_of(a.variant);
        ctx.restir_options = restir_options_of(a.variant);
        ctx.alpha_test = g_scene_has_alpha_tested_materials;
        B2_SET_RAY_BOUNCE(ctx, 0);
#pragma omp for schedule(dynamic)
        for (int y = a.win_y0; y < a.win_y1; y++)
            for (int x = a.win_x0; x < a.win_x1; x++)
            {
                if (a.staged && frame_of_debug(a))
                {
                    g_debug_rays_x = a.debug_x;
                    g_debug_rays_y = a.debug_y;
                }
                if (pass == B2_PASS_MEGAKERNEL && a.staged)
                    staged_megakernel_pixel<V>(ctx, (uint32_t)x, (uint32_t)y);
                else
                    b2_run_pass_pixel<V>(pass, ctx, (uint32_t)x, (uint32_t)y);
            }
#pragma omp critical
        {
            g_port_counters[0] += local.rays;
            g_port_counters[1] += local.nodes;
            g_port_counters[2] += local.tris;
        }
    }
}

/* Function-level known-answer vectors: same inputs and record layout as run_bsdf_kat in ref_harness.cpp, computed with the
 * product's bsdf_eval / bsdf_sample (variant c2: principled BSDF). */
static void kat_directions(unsigned int index, V3& normal, V3& view, V3& light, unsigned int& seed)
{
    Rng g = make_rng(wang_hash(index + 1u));
    auto unit = [&]() {
        float z = 2.0f * g.next_float() - 1.0f, phi = 6.2831853f * g.next_float(), r = sqrtf(fmaxf(0.0f, 1.0f - z * z));
        return v3(r * cosf(phi), r * sinf(phi), z);
    };
    normal = unit();
    view = unit();
    if (dot(view, normal) < 0.0f) view = view * -1.0f;
    light = unit();
    seed = g.next_u32() | 1u;
}

#ifdef B2_PORT_HAVE_C2
static void run_bsdf_kat(const b2pt_render_data& rd, const Args& a, int n_materials)
{
    const int per_material = 64;
    std::vector<float> eval_out((size_t)n_materials * per_material * 4), sample_out((size_t)n_materials * per_material * 7);
    for (int m = 0; m < n_materials; m++)
        for (int k = 0; k < per_material; k++)
        {
            V3 normal, view, light;
            unsigned int seed;
            kat_directions((unsigned int)(m * per_material + k), normal, view, light, seed);
            Mat material = load_scene_material(rd.buffers.materials_buffer, m, rd.bsdfs_data.white_furnace_mode != 0);
            VolumeState volume_state;
            volume_state.init(); /* = the reference's default-constructed RayVolumeState (the harness used to leave it unwritten) */
            Rng rng = make_rng(seed);
            {
                int info = B2_LOBE_NO_INFO;
                BsdfCtx ctx = make_bsdf_ctx(view, normal, normal, light, &info, &volume_state, false, &material, 1, 0.0f);
                float pdf = 0.0f;
                Rgb c = bsdf_eval<B2VariantC2>(rd, ctx, pdf);
                float* out = &eval_out[((size_t)m * per_material + k) * 4];
                out[0] = c.r; out[1] = c.g; out[2] = c.b; out[3] = pdf;
            }
            {
                int info = B2_LOBE_NO_INFO;
                BsdfCtx ctx = make_bsdf_ctx(view, normal, normal, v3(0.0f), &info, &volume_state, false, &material, 1, 0.0f);
                float pdf = 0.0f;
                V3 dir = v3(0.0f);
                Rgb c = bsdf_sample<B2VariantC2>(rd, ctx, dir, pdf, rng);
                float* out = &sample_out[((size_t)m * per_material + k) * 7];
                out[0] = c.r; out[1] = c.g; out[2] = c.b; out[3] = pdf; out[4] = dir.x; out[5] = dir.y; out[6] = dir.z;
            }
        }
    dump_array(a.dump, "kat_bsdf_eval", eval_out.data(), eval_out.size());
    dump_array(a.dump, "kat_bsdf_sample", sample_out.data(), sample_out.size());
    std::printf("{\"oracle\": \"port\", \"kat\": \"bsdf\", \"materials\": %d, \"per_material\": %d}\n", n_materials, per_material);
}

/* same records as run_light_kat in ref_harness.cpp, from the product's samplers (c1 = uniform base strategy, c2 = power) */
static void run_light_kat(const b2pt_render_data& rd, const Args& a)
{
    const int n = 4096;
    std::vector<float> uniform_out((size_t)n * 12), power_out((size_t)n * 12), env_sample_out, env_eval_out;
    auto put = [](float* out, const LightSample& s) {
        std::memcpy(&out[0], &s.emissive_triangle_index, 4);
        out[1] = s.point_on_light.x; out[2] = s.point_on_light.y; out[3] = s.point_on_light.z;
        out[4] = s.light_source_normal.x; out[5] = s.light_source_normal.y; out[6] = s.light_source_normal.z;
        out[7] = s.light_area; out[8] = s.emission.r; out[9] = s.emission.g; out[10] = s.emission.b; out[11] = s.area_measure_pdf;
    };
    for (int i = 0; i < n; i++)
    {
        Rng ra = make_rng(wang_hash(7919u * (unsigned)i + 1u) | 1u), rb = make_rng(wang_hash(7919u * (unsigned)i + 1u) | 1u);
        put(&uniform_out[(size_t)i * 12], sample_one_emissive_triangle<B2VariantC1>(rd, ra));
        put(&power_out[(size_t)i * 12], sample_one_emissive_triangle<B2VariantC2>(rd, rb));
    }
    dump_array(a.dump, "kat_light_uniform", uniform_out.data(), uniform_out.size());
    dump_array(a.dump, "kat_light_power", power_out.data(), power_out.size());
    if (!a.envmap.empty())
    {
        env_sample_out.resize((size_t)n * 7);
        env_eval_out.resize((size_t)n * 4);
        for (int i = 0; i < n; i++)
        {
            Rng rng = make_rng(wang_hash(104729u * (unsigned)i + 3u) | 1u);
            V3 dir = v3(0.0f, 1.0f, 0.0f);
            float pdf = 0.0f;
            Rgb c = envmap_sample<B2VariantC2>(rd.world_settings, dir, pdf, rng);
            float* out = &env_sample_out[(size_t)i * 7];
            out[0] = c.r; out[1] = c.g; out[2] = c.b; out[3] = pdf; out[4] = dir.x; out[5] = dir.y; out[6] = dir.z;

            V3 normal, view, light;
            unsigned int seed;
            kat_directions(100000u + (unsigned)i, normal, view, light, seed);
            float eval_pdf = 0.0f;
            Rgb e = envmap_eval<B2VariantC2>(rd.world_settings, light, eval_pdf);
            float* eo = &env_eval_out[(size_t)i * 4];
            eo[0] = e.r; eo[1] = e.g; eo[2] = e.b; eo[3] = eval_pdf;
        }
        dump_array(a.dump, "kat_envmap_sample", env_sample_out.data(), env_sample_out.size());
        dump_array(a.dump, "kat_envmap_eval", env_eval_out.data(), env_eval_out.size());
    }
    std::printf("{\"oracle\": \"port\", \"kat\": \"lights\", \"records\": %d, \"envmap\": %d}\n", n, a.envmap.empty() ? 0 : 1);
}
#endif

int main(int argc, char** argv)
{
    Args a = parse_args(argc, argv);
    if (a.threads > 0)
        omp_set_num_threads(a.threads);
    if (a.scene.empty() || a.luts.empty())
    {
        std::fprintf(stderr, "usage: %s --variant c1|c2|c3|c4 --scene X.b2scene --luts L.b2luts --width W --height H --spp N --bounces B [--dump prefix]\n", argv[0]);
        return 2;
    }

    B2HostScene scene;
    std::string err = b2scene_read(a.scene.c_str(), scene.blob);
    if (err.empty() && !b2_host_scene_prepare(scene, err) && err.empty())
        err = "scene preparation failed";
    if (!err.empty()) { std::fprintf(stderr, "%s\n", err.c_str()); return 1; }
    g_scene_has_alpha_tested_materials = scene.has_alpha_tested_materials;

    LutStore luts;
    if (!luts.load(a.luts.c_str())) { std::fprintf(stderr, "cannot read LUT blob %s\n", a.luts.c_str()); return 1; }

    const size_t n = (size_t)a.width * a.height;
    B2FrameLoop loop;
    b2pt_render_data& rd = loop.rd;
    loop.host_rng = make_rng(a.seed);
    rd.render_settings.render_resolution = b2pt_int2{ a.width, a.height };
    rd.render_settings.nb_bounces = a.bounces;
    rd.render_settings.samples_per_frame = a.spp;
    rd.render_settings.output_debug_sample_N = a.spp - 1;
    rd.render_settings.restir_di_settings.do_fused_spatiotemporal = a.fused;
    rd.render_settings.restir_di_settings.common_spatial_pass.use_adaptive_directional_spatial_reuse = a.directional_reuse ? 1 : 0;
    rd.render_settings.restir_di_settings.common_spatial_pass.number_of_passes = a.spatial_passes;
    rd.render_settings.restir_di_settings.use_confidence_weights = a.confidence_weights ? 1 : 0;
    if (!a.spatial_pass)
        rd.render_settings.restir_di_settings.common_spatial_pass.do_spatial_reuse_pass = 0;
    if (a.neighbors >= 0)
    {
        rd.render_settings.restir_di_settings.common_spatial_pass.reuse_neighbor_count = a.neighbors;
        rd.render_settings.restir_di_settings.common_spatial_pass.disocclusion_reuse_count = a.neighbors;
    }

    /* camera */
    std::memcpy(loop.camera.inverse_view, scene.blob.camera.inverse_view, 64);
    std::memcpy(loop.camera.inverse_projection, scene.blob.camera.inverse_projection, 64);
    std::memcpy(loop.camera.view_projection, scene.blob.camera.view_projection, 64);
    loop.camera.position = b2pt_float3{ scene.blob.camera.position[0], scene.blob.camera.position[1], scene.blob.camera.position[2] };
    loop.camera.vertical_fov = scene.blob.camera.vertical_fov;
    loop.camera.sensor_width = a.width;
    loop.camera.sensor_height = a.height;
    loop.camera.do_jittering = !a.no_jitter;
    rd.current_camera = loop.camera;

    /* scene buffers */
    std::vector<b2pt_rgb> framebuffer(n, b2pt_rgb{ 0, 0, 0 });
    rd.buffers.accumulated_ray_colors = framebuffer.data();
    rd.buffers.triangles_indices = scene.blob.indices.data();
    rd.buffers.vertices_positions = reinterpret_cast<b2pt_float3*>(scene.blob.positions.data());
    rd.buffers.has_vertex_normals = scene.blob.has_normals.data();
    rd.buffers.vertex_normals = reinterpret_cast<b2pt_float3*>(scene.blob.normals.data());
    rd.buffers.texcoords = reinterpret_cast<b2pt_float2*>(scene.blob.texcoords.data());
    rd.buffers.triangles_areas = scene.triangle_areas.data();
    rd.buffers.material_indices = scene.blob.material_indices.data();
    rd.buffers.material_opaque = scene.material_opaque.data();
    rd.buffers.emissive_triangles_count = (int)scene.emissive_triangles.size();
    rd.buffers.emissive_triangles_primitive_indices = scene.emissive_triangles.data();
    rd.buffers.emissive_triangles_primitive_indices_and_emissive_textures = scene.emissive_triangles_and_textures.data();
    /* material textures: host-side descriptors of the product's device format (b2_material.h), sRGB table from the same powf */
    std::vector<b2pt_texture> texture_descriptors;
    std::vector<float> srgb_to_linear(256);
    for (int v = 0; v < 256; v++)
        srgb_to_linear[v] = powf((float)v / 255.0f, 2.2f);
    for (const B2SceneTexture& t : scene.blob.textures)
        texture_descriptors.push_back(b2pt_texture{ t.texels.data(), (int)t.width, (int)t.height, (int)t.channels, srgb_to_linear.data() });
    rd.buffers.material_textures = texture_descriptors.empty() ? nullptr : texture_descriptors.data();
    rd.buffers.emissives_power_alias_table.alias_table_alias = scene.power_alias_alias.data();
    rd.buffers.emissives_power_alias_table.alias_table_probas = scene.power_alias_probas.data();
    rd.buffers.emissives_power_alias_table.sum_elements = scene.power_sum;
    rd.buffers.emissives_power_alias_table.size = (unsigned)scene.emissive_triangles.size();
    {
        B2MaterialSoAHost& m = scene.materials;
        b2pt_materials_soa& d = rd.buffers.materials_buffer;
        d.flags = m.flags.data();
        d.emission = reinterpret_cast<b2pt_rgb*>(m.emission.data());
        d.base_color_roughness = m.base_color_roughness.data();
        d.oren_nayar_sigma = m.oren_nayar_sigma.data();
        d.metallic_F90_and_metallic = m.metallic_F90_and_metallic.data();
        d.metallic_F82_packed_and_diffuse_transmission = m.metallic_F82_packed_and_diffuse_transmission.data();
        d.metallic_F90_falloff_exponent = m.metallic_F90_falloff_exponent.data();
        d.anisotropy_and_rotation_and_second_roughness = m.anisotropy_and_rotation_and_second_roughness.data();
        d.specular_color_and_tint_factor = m.specular_color_and_tint_factor.data();
        d.specular_and_darkening_and_coat_roughness = m.specular_and_darkening_and_coat_roughness.data();
        d.coat_medium_thickness = m.coat_medium_thickness.data();
        d.coat_and_medium_absorption = m.coat_and_medium_absorption.data();
        d.coat_roughening_darkening_anisotropy_and_rotation = m.coat_roughening_darkening_anisotropy_and_rotation.data();
        d.coat_ior = m.coat_ior.data();
        d.sheen_and_color = m.sheen_and_color.data();
        d.ior = m.ior.data();
        d.absorption_color_packed = m.absorption_color_packed.data();
        d.absorption_at_distance = m.absorption_at_distance.data();
        d.sheen_roughness_transmission_dispersion_thin_film = m.sheen_roughness_transmission_dispersion_thin_film.data();
        d.dispersion_abbe_number = m.dispersion_abbe_number.data();
        d.thin_film_ior = m.thin_film_ior.data();
        d.thin_film_thickness = m.thin_film_thickness.data();
        d.thin_film_kappa_3 = m.thin_film_kappa_3.data();
        d.thin_film_base_ior_override = m.thin_film_base_ior_override.data();
        d.alpha_thin_film_hue_dielectric_priority = m.alpha_thin_film_hue_dielectric_priority.data();
        d.normal_map_emission_index = m.normal_map_emission_index.data();
        d.base_color_roughness_metallic_index = m.base_color_roughness_metallic_index.data();
        d.roughness_and_metallic_index = m.roughness_and_metallic_index.data();
        d.anisotropic_specular_index = m.anisotropic_specular_index.data();
        d.coat_sheen_index = m.coat_sheen_index.data();
        d.specular_transmission_index = m.specular_transmission_index.data();
    }

    rd.bsdfs_data.GGX_conductor_directional_albedo = &luts.luts["ggx_conductor"];
    rd.bsdfs_data.glossy_dielectric_directional_albedo = &luts.luts["glossy_dielectric"];
    rd.bsdfs_data.GGX_glass_directional_albedo = &luts.luts["ggx_glass"];
    rd.bsdfs_data.GGX_glass_directional_albedo_inverse = &luts.luts["ggx_glass_inv"];
    rd.bsdfs_data.GGX_thin_glass_directional_albedo = &luts.luts["ggx_thin_glass"];
    rd.bsdfs_data.sheen_ltc_parameters_texture = &luts.luts["sheen_ltc"];

    /* environment map (same host preparation the product's renderer uploads: b2_host_envmap_prepare) */
    B2HostEnvmap env;
    if (!a.envmap.empty())
    {
        std::vector<float> rgba;
        int ew = 0, eh = 0;
        if (!b2_envmap_blob_read(a.envmap.c_str(), rgba, ew, eh)) { std::fprintf(stderr, "cannot read envmap blob %s\n", a.envmap.c_str()); return 1; }
        b2_host_envmap_prepare(env, rgba.data(), ew, eh);
        b2pt_world_settings& w = rd.world_settings;
        w.envmap = env.packed.data();
        w.envmap_alias_table.alias_table_probas = env.alias_probas.data();
        w.envmap_alias_table.alias_table_alias = env.alias_alias.data();
        w.envmap_alias_table.sum_elements = env.total_luminance;
        w.envmap_alias_table.size = (unsigned)((size_t)ew * eh);
        w.envmap_total_sum = env.total_luminance;
        w.envmap_width = (unsigned)ew;
        w.envmap_height = (unsigned)eh;
        w.ambient_light_type = 2;
    }

    /* aux + G-buffers */
    std::vector<uint8_t> pixel_active(n, 0);
    std::vector<b2pt_rgb> denoiser_albedo(n, b2pt_rgb{ 0, 0, 0 });
    std::vector<b2pt_float3> denoiser_normals(n, b2pt_float3{ 0, 0, 0 });
    std::vector<int> pixel_sample_count(n, 0), pixel_converged(n, 0);
    std::vector<float> pixel_sq_lum(n, 0.0f);
    uint8_t still_one_ray_active = 1;
    unsigned int converged_count = 0;
    rd.aux_buffers.pixel_active = pixel_active.data();
    rd.aux_buffers.denoiser_albedo = denoiser_albedo.data();
    rd.aux_buffers.denoiser_normals = denoiser_normals.data();
    rd.aux_buffers.pixel_sample_count = pixel_sample_count.data();
    rd.aux_buffers.pixel_converged_sample_count = pixel_converged.data();
    rd.aux_buffers.pixel_squared_luminance = pixel_sq_lum.data();
    rd.aux_buffers.still_one_ray_active = &still_one_ray_active;
    rd.aux_buffers.pixel_count_converged_so_far = &converged_count;

    struct GBufferHost { std::vector<uint32_t> materials, geo, shade; std::vector<b2pt_float3> pos; std::vector<int> prim; } g[2];
    b2pt_gbuffer* gb[2] = { &rd.g_buffer, &rd.g_buffer_prev_frame };
    for (int i = 0; i < 2; i++)
    {
        g[i].materials.assign(n * PM_WORDS, 0);
        g[i].geo.assign(n, 0);
        g[i].shade.assign(n, 0);
        g[i].pos.assign(n, b2pt_float3{ 0, 0, 0 });
        g[i].prim.assign(n, 0);
        gb[i]->materials = g[i].materials.data();
        gb[i]->geometric_normals = g[i].geo.data();
        gb[i]->shading_normals = g[i].shade.data();
        gb[i]->primary_hit_position = g[i].pos.data();
        gb[i]->first_hit_prim_index = g[i].prim.data();
    }

    std::vector<uint32_t> reservoirs[3];
    for (auto& r : reservoirs)
    {
        r.assign(n * 9, 0);
        std::fill(r.begin() + (size_t)DI_PLANE_TRIANGLE * n, r.begin() + (size_t)(DI_PLANE_TRIANGLE + 1) * n, 0xFFFFFFFFu); /* a default ReSTIRDISample: triangle -1 */
    }
    loop.initial_reservoirs = reservoirs[0].data();
    loop.spatial_reservoirs_1 = reservoirs[1].data();
    loop.spatial_reservoirs_2 = reservoirs[2].data();
    rd.render_settings.restir_di_settings.restir_output_reservoirs = loop.spatial_reservoirs_1;

    B2BvhDevice bvh;
    bvh.nodes = scene.bvh.nodes.data();
    bvh.tris = scene.bvh.tris.data();
    bvh.node_count = (uint32_t)scene.bvh.nodes.size();
    bvh.tri_count = (uint32_t)scene.bvh.tris.size();
    rd.GPU_BVH = &bvh;

    if (a.trace_ray)
    {
        /* the counterpart of oracle/ray_check.cpp: the product's closest-hit and any-hit traversals on one ray */
        uint64_t stack[B2_STACK_ENTRIES];
        Ctx ctx;
        ctx.rd = &rd;
        ctx.bvh = bvh;
        ctx.stack = stack;
        ctx.stack_stride = 1;
        ctx.counters = nullptr;
        ctx.faults = g_port_faults;
        ctx.light_presampling = 0;
        ctx.bias_correction = 4;
        ctx.restir_options = 4 | 8;
        ctx.alpha_test = false;
        B2_SET_RAY_BOUNCE(ctx, 0);
        V3 origin = v3(a.ray[0], a.ray[1], a.ray[2]), direction = v3(a.ray[3], a.ray[4], a.ray[5]);
        B2Hit closest = b2_trace(ctx, false, origin, direction, 3.0e38f, a.ray_skip_prim);
        B2Hit any = b2_trace(ctx, true, origin, direction, a.ray[6], a.ray_skip_prim);
        std::printf("{\"trace_ray\": true, \"closest_prim\": %d, \"closest_t\": %.9g, \"occluded\": %s}\n", closest.prim, closest.t, any.prim >= 0 ? "true" : "false");
        return 0;
    }
#ifdef B2_PORT_HAVE_C2
    if (a.kat)
    {
        run_bsdf_kat(rd, a, (int)scene.blob.n_materials());
        run_light_kat(rd, a);
        return 0;
    }
#endif
    loop.restir_di = is_c4(a.variant);
    loop.restir_di_presampling = a.variant == "c4p" || a.variant == "c4pu";
    std::vector<unsigned long long> directional_masks((size_t)a.width * a.height, 0ull);
    std::vector<uint8_t> directional_radius((size_t)a.width * a.height, 0);
    rd.render_settings.restir_di_settings.common_spatial_pass.per_pixel_spatial_reuse_directions_mask_ull = directional_masks.data();
    rd.render_settings.restir_di_settings.common_spatial_pass.per_pixel_spatial_reuse_radius = directional_radius.data();
    std::vector<B2PresampledLight> presampled_lights;
    if (loop.restir_di_presampling)
    {
        const b2pt_restir_di_settings& di = rd.render_settings.restir_di_settings;
        presampled_lights.resize((size_t)di.light_presampling.number_of_subsets * di.light_presampling.subset_size);
        std::memset(presampled_lights.data(), 0, presampled_lights.size() * sizeof(B2PresampledLight));
        rd.render_settings.restir_di_settings.light_presampling.light_samples = presampled_lights.data();
    }

    int frame = 0;
    auto launch = [&](B2Pass pass) {
        if (pass == B2_PASS_CAMERA_RAYS)
        {
            still_one_ray_active = 0;
            converged_count = 0;
        }
        if (a.variant == "c1") run_pass<B2VariantC1>(pass, rd, bvh, a);
#ifdef B2_PORT_HAVE_C2
        else if (a.variant == "c2") run_pass<B2VariantC2>(pass, rd, bvh, a);
#endif
#ifdef B2_PORT_HAVE_C3
        else if (a.variant == "c3") run_pass<B2VariantC3>(pass, rd, bvh, a);
#endif
#ifdef B2_PORT_HAVE_C4
        else if (pass == B2_PASS_RESTIR_DIRECTIONAL_REUSE)
        {
            b2pt_restir_spatial_common& sp = rd.render_settings.restir_di_settings.common_spatial_pass;
#pragma omp parallel for schedule(dynamic)
            for (int y = a.win_y0; y < a.win_y1; y++)
                for (int x = a.win_x0; x < a.win_x1; x++)
                    restir_directional_reuse_compute_pixel(rd, (uint32_t)x, (uint32_t)y, sp.per_pixel_spatial_reuse_directions_mask_ull, sp.per_pixel_spatial_reuse_radius);
        }
        else if (pass == B2_PASS_RESTIR_DI_PRESAMPLING)
        {
            /* ReSTIR_DI_LightsPresampling: one "thread" per presampled light, whole buffer (not windowed) */
            const b2pt_restir_di_settings& di = rd.render_settings.restir_di_settings;
            B2LightPresamplingParameters params;
            params.number_of_subsets = di.light_presampling.number_of_subsets;
            params.subset_size = di.light_presampling.subset_size;
            params.out_light_samples = presampled_lights.data();
            params.envmap_sampling_probability = di.initial_candidates.envmap_candidate_probability;
#pragma omp parallel for schedule(static)
            for (int index = 0; index < params.number_of_subsets * params.subset_size; index++)
                restir_di_presample_light<B2VariantC4>(rd, params, (uint32_t)index, a.variant == "c4pu");
        }
        /* c4l<N>: ReSTIR_DI_LaterBouncesSamplingStrategy = N (ReSTIRDIOptions.h:18-21,140): only the MegaKernel differs */
        else if (a.variant == "c4l0" && pass == B2_PASS_MEGAKERNEL) run_pass<B2Variant<B2_BSDF_NONE, B2_LSS_RESTIR_DI, B2_BASE_POWER, B2_LATER_ONE_LIGHT>>(pass, rd, bvh, a);
        else if ((a.variant == "c4l1" || a.variant == "c4x2") && pass == B2_PASS_MEGAKERNEL) run_pass<B2Variant<B2_BSDF_NONE, B2_LSS_RESTIR_DI, B2_BASE_POWER, B2_LATER_BSDF>>(pass, rd, bvh, a);
        else if ((a.variant == "c4l2" || a.variant == "c4x1") && pass == B2_PASS_MEGAKERNEL) run_pass<B2Variant<B2_BSDF_NONE, B2_LSS_RESTIR_DI, B2_BASE_POWER, B2_LATER_MIS>>(pass, rd, bvh, a);
        else if (is_c4(a.variant)) run_pass<B2VariantC4>(pass, rd, bvh, a);
#endif
        else { std::fprintf(stderr, "unknown variant %s\n", a.variant.c_str()); std::exit(2); }
        if (pass == B2_PASS_RESTIR_DI_INITIAL_CANDIDATES && frame == a.debug_frame && a.debug_x >= 0)
        {
            /* debugging aid: the visibility ray of one pixel's initial reservoir, traced through the BVH and by brute force over every triangle */
            const size_t pixel = (size_t)a.debug_x + (size_t)a.debug_y * a.width;
            const uint32_t* soa = reinterpret_cast<const uint32_t*>(rd.render_settings.restir_di_settings.initial_candidates.output_reservoirs);
            auto word = [&](int plane) { return soa[(size_t)plane * n + pixel]; };
            auto as_float = [](uint32_t w) { float f; std::memcpy(&f, &w, 4); return f; };
            V3 point = v3(as_float(word(3)), as_float(word(4)), as_float(word(5)));
            V3 origin = load_v3(rd.g_buffer.primary_hit_position, (uint32_t)pixel);
            int skip = rd.g_buffer.first_hit_prim_index[pixel];
            V3 d = point - origin;
            float distance = length(d);
            d = d / distance;
            float t_max = distance - 1.0e-4f;
            std::fprintf(stderr, "debug pixel (%d, %d) frame %d: M %u UCW %g triangle %d flags %u origin (%.9g %.9g %.9g) dir (%.9g %.9g %.9g) distance %.9g skip %d\n", a.debug_x, a.debug_y,
                         frame, word(0), as_float(word(2)), (int)word(6), word(8) & 0xFF, origin.x, origin.y, origin.z, d.x, d.y, d.z, distance, skip);
            for (uint32_t i = 0; i < bvh.tri_count; i++)
            {
                const B2Tri& tri = bvh.tris[i];
                float t, u, v;
                if (tri.prim != skip && b2_intersect_triangle(origin, d, v3(tri.ax, tri.ay, tri.az), v3(tri.e1x, tri.e1y, tri.e1z), v3(tri.e2x, tri.e2y, tri.e2z), t, u, v) && t < t_max + 1.0e-3f)
                    std::fprintf(stderr, "  brute force: prim %d t %.9g (t_max %.9g) u %.9g v %.9g\n", tri.prim, t, t_max, u, v);
            }
        }
        if (a.dump_per_sample && !a.dump.empty() &&
            (pass == B2_PASS_RESTIR_DI_INITIAL_CANDIDATES || pass == B2_PASS_RESTIR_DI_TEMPORAL || pass == B2_PASS_RESTIR_DI_SPATIAL || pass == B2_PASS_RESTIR_DI_SPATIOTEMPORAL))
        {
            /* same files as oracle/ref_harness.cpp::dump_pass_reservoirs, transposed from the SoA planes to 36-byte records */
            const b2pt_restir_di_settings& di = rd.render_settings.restir_di_settings;
            const char* name = nullptr;
            const void* buffer = nullptr;
            if (pass == B2_PASS_RESTIR_DI_INITIAL_CANDIDATES) { name = "initial"; buffer = di.initial_candidates.output_reservoirs; }
            else if (pass == B2_PASS_RESTIR_DI_TEMPORAL) { name = "temporal"; buffer = di.temporal_pass.output_reservoirs; }
            else { name = "output"; buffer = di.spatial_pass.output_reservoirs; }
            std::vector<uint32_t> aos(n * 9);
            const uint32_t* soa = reinterpret_cast<const uint32_t*>(buffer);
            for (size_t i = 0; i < n; i++)
                for (int p = 0; p < 9; p++)
                    aos[i * 9 + p] = soa[p * n + i];
            std::string file = std::string("restir_") + name + "_f" + std::to_string(frame);
            dump_array(a.dump, file.c_str(), aos.data(), aos.size());
        }
        if (pass == B2_PASS_CAMERA_RAYS && frame == 1)
        {
            dump_array(a.dump, "gbuffer_prim", g[0].prim.data(), n);
            dump_array(a.dump, "gbuffer_position", g[0].pos.data(), n);
            dump_array(a.dump, "gbuffer_geometric_normal", g[0].geo.data(), n);
            dump_array(a.dump, "gbuffer_shading_normal", g[0].shade.data(), n);
            dump_array(a.dump, "gbuffer_material", g[0].materials.data(), n * PM_WORDS);
        }
    };

    std::vector<B2SceneCamera> camera_frames;
    if (!a.cameras.empty())
    {
        FILE* f = std::fopen(a.cameras.c_str(), "rb");
        if (!f) { std::fprintf(stderr, "cannot read camera file %s\n", a.cameras.c_str()); return 1; }
        B2SceneCamera c;
        while (std::fread(&c, sizeof(c), 1, f) == 1)
            camera_frames.push_back(c);
        std::fclose(f);
    }

    for (const std::string& poke : a.pokes)
    {
        size_t offset = 0;
        char kind = 0;
        double value = 0.0;
        if (std::sscanf(poke.c_str(), "%zu:%c:%lf", &offset, &kind, &value) != 3 || offset + 4 > sizeof(b2pt_render_data)) { std::fprintf(stderr, "bad --poke %s\n", poke.c_str()); return 2; }
        char* base = reinterpret_cast<char*>(&rd) + offset;
        if (kind == 'i') { int v = (int)value; std::memcpy(base, &v, 4); }
        else if (kind == 'f') { float v = (float)value; std::memcpy(base, &v, 4); }
        else if (kind == 'b') { unsigned char v = (unsigned char)value; std::memcpy(base, &v, 1); }
        else { std::fprintf(stderr, "bad --poke kind %c\n", kind); return 2; }
    }

    auto start = std::chrono::high_resolution_clock::now();
    for (frame = 1; frame <= a.spp; frame++)
    {
        if (!camera_frames.empty())
        {
            const B2SceneCamera& c = camera_frames[std::min<size_t>((size_t)(frame - 1), camera_frames.size() - 1)];
            std::memcpy(loop.camera.inverse_view, c.inverse_view, 64);
            std::memcpy(loop.camera.inverse_projection, c.inverse_projection, 64);
            std::memcpy(loop.camera.view_projection, c.view_projection, 64);
            loop.camera.position = b2pt_float3{ c.position[0], c.position[1], c.position[2] };
            loop.camera.vertical_fov = c.vertical_fov;
        }
        g_current_frame = frame;
        loop.render_one_sample(launch);
        if (a.dump_per_sample)
        {
            std::string name = "framebuffer_s" + std::to_string(frame);
            dump_array(a.dump, name.c_str(), framebuffer.data(), n);
        }
    }
    double seconds = std::chrono::duration<double>(std::chrono::high_resolution_clock::now() - start).count();

    dump_array(a.dump, "framebuffer", framebuffer.data(), n);
    dump_array(a.dump, "packed_materials", scene.packed_materials.data(), scene.packed_materials.size());
    dump_array(a.dump, "triangle_areas", scene.triangle_areas.data(), scene.triangle_areas.size());
    dump_array(a.dump, "emissive_prims", scene.emissive_triangles.data(), scene.emissive_triangles.size());
    dump_array(a.dump, "power_alias", scene.power_alias_alias.data(), scene.power_alias_alias.size());
    dump_array(a.dump, "power_probas", scene.power_alias_probas.data(), scene.power_alias_probas.size());

    long long window_pixels = (long long)(a.win_x1 - a.win_x0) * (a.win_y1 - a.win_y0);
    std::printf("{\"oracle\": \"port\", \"variant\": \"%s\", \"width\": %d, \"height\": %d, \"window_pixels\": %lld, \"spp\": %d, \"bounces\": %d, \"threads\": %d, "
                "\"render_seconds\": %.6f, \"bvh_build_seconds\": %.3f, \"mpaths_per_s\": %.6f, \"bvh_nodes\": %zu, \"bvh_depth\": %d, \"triangles\": %zu, \"emissive_triangles\": %zu, \"rays\": %llu, \"nodes_per_ray\": %.3f, \"triangles_per_ray\": %.3f}\n",
                a.variant.c_str(), a.width, a.height, window_pixels, a.spp, a.bounces, omp_get_max_threads(), seconds, scene.bvh_build_seconds,
                (double)window_pixels * a.spp / seconds / 1.0e6, scene.bvh.nodes.size(), scene.bvh.stats.max_depth, scene.blob.n_triangles(), scene.emissive_triangles.size(), g_port_counters[0], (double)g_port_counters[1] / (double)(g_port_counters[0] ? g_port_counters[0] : 1),
                (double)g_port_counters[2] / (double)(g_port_counters[0] ? g_port_counters[0] : 1));
    return 0;
}
