/*
 * b2pt.h -- C ABI of libb2pt.so, the B200-native device layer of the path tracer.
 *
 * Two groups of entry points:
 *
 *  (1) The DEVICE LAYER: what a maintainer binds in place of src/HIPRT-Orochi (+ the JIT half of src/Compiler) of
 *      the reference. Each entry cites the reference interface it replaces. Plain pointers and sizes only; every
 *      call returns 0 on success or a B2PT_ERR_* code, with a thread-local message in b2pt_last_error_string()
 *      (the reference wraps every oro / hiprt call in OROCHI_CHECK_ERROR / HIPRT_CHECK_ERROR -> log + exit,
 *      src/HIPRT-Orochi/HIPRTOrochiUtils.h:15-17; INTEGRATION.md shows the equivalent B2PT_CHECK macro).
 *
 *  (2) The HEADLESS RENDERER: a host-side mirror of the reference's GPURenderer frame / sample loop
 *      (set_scene / set_camera / set_envmap / render / get_framebuffer, src/Renderer/GPURenderer.h:61-346) with the
 *      CPURenderer's pass order and host RNG draws (src/Renderer/CPURenderer.cpp:534-604), driving group (1).
 *      tests/ and bench.py go through this group with HOST buffers.
 *
 * There is no CPU fallback anywhere behind this header: without a CUDA device b2pt_ctx_create fails.
 */
#ifndef B2PT_H
#define B2PT_H

#include <stddef.h>
#include <stdint.h>

#include "b2pt_render_data.h"

#ifdef __cplusplus
extern "C" {
#endif

#define B2PT_OK 0
#define B2PT_ERR_INVALID_ARGUMENT 1
#define B2PT_ERR_CUDA 2
#define B2PT_ERR_NO_DEVICE 3
#define B2PT_ERR_NO_VARIANT 4 /* the requested macro combination is not among the precompiled variants */
#define B2PT_ERR_UNSUPPORTED_SCENE 5
#define B2PT_ERR_IO 6

typedef struct b2pt_ctx b2pt_ctx;
typedef struct b2pt_bvh b2pt_bvh;
typedef struct b2pt_kernel b2pt_kernel;
typedef struct b2pt_renderer b2pt_renderer;
typedef void* b2pt_stream; /* cudaStream_t */
typedef void* b2pt_event;  /* cudaEvent_t */

const char* b2pt_last_error_string(void);
const char* b2pt_version_string(void);

/* ------------------------------------------------------------------------------------------------------------
 * (1) device layer
 * ------------------------------------------------------------------------------------------------------------ */

/* Context. Replaces HIPRTOrochiCtx::init (oroInitialize / oroInit / oroDeviceGet / oroCtxCreate /
 * hiprtCreateContext, src/HIPRT-Orochi/HIPRTOrochiCtx.h:40-106) and oroCtxSetCurrent
 * (src/Renderer/GPURenderer.cpp:245,444,854). One context per GPU. */
typedef struct
{
    char name[256];
    int sm_count;
    int compute_major, compute_minor;
    size_t total_memory_bytes;
    size_t l2_bytes;
    size_t shared_memory_per_block_optin;
    int clock_khz;
} b2pt_device_props_t;

int b2pt_ctx_create(int device_index, b2pt_ctx** out_ctx);
int b2pt_ctx_destroy(b2pt_ctx* ctx);
int b2pt_ctx_make_current(b2pt_ctx* ctx);
int b2pt_ctx_device_props(b2pt_ctx* ctx, b2pt_device_props_t* out);

/* Memory. Replaces OrochiBuffer<T> (resize / free / upload_data / download_data / memset_whole_buffer /
 * memcpy_from / resize_host_pinned_mem over oroMalloc / oroFree / oroMemcpy* / oroHostMalloc,
 * src/HIPRT-Orochi/OrochiBuffer.h:20-122,175-360). Device pointers are plain CUDA device pointers and are stored
 * verbatim in HIPRTRenderData; pointer identity is preserved (the host compares them, ReSTIRDIRenderPass.cpp:597). */
int b2pt_malloc(b2pt_ctx* ctx, size_t bytes, void** out_device_ptr);
int b2pt_free(b2pt_ctx* ctx, void* device_ptr);
int b2pt_memcpy_h2d(b2pt_ctx* ctx, void* dst_device, const void* src_host, size_t bytes, b2pt_stream stream);
int b2pt_memcpy_d2h(b2pt_ctx* ctx, void* dst_host, const void* src_device, size_t bytes, b2pt_stream stream);
int b2pt_memcpy_d2d(b2pt_ctx* ctx, void* dst_device, const void* src_device, size_t bytes, b2pt_stream stream);
int b2pt_memset(b2pt_ctx* ctx, void* device_ptr, int byte_value, size_t bytes, b2pt_stream stream);
int b2pt_host_alloc_pinned(b2pt_ctx* ctx, size_t bytes, void** out_host_ptr);
int b2pt_host_free_pinned(b2pt_ctx* ctx, void* host_ptr);
/* Page-locks / releases a host range the caller already owns, so that asynchronous copies from it run at full PCIe / C2C rate
 * (the renderer does this once for the host copies of the scene arrays it re-sends every step). No reference counterpart:
 * OrochiBuffer::upload_data copies from pageable std::vector storage (OrochiBuffer.h:263-289). */
int b2pt_host_register(b2pt_ctx* ctx, const void* host_ptr, size_t bytes);
int b2pt_host_unregister(b2pt_ctx* ctx, const void* host_ptr);

/* Streams / events. Replaces oroStreamCreate / oroStreamSynchronize / oroEventCreate / Record / Synchronize /
 * ElapsedTime (src/Renderer/GPURenderer.cpp:66,426; src/Compiler/GPUKernel.cpp:190-201,228-240). */
int b2pt_stream_create(b2pt_ctx* ctx, b2pt_stream* out_stream);
int b2pt_stream_destroy(b2pt_ctx* ctx, b2pt_stream stream);
int b2pt_stream_synchronize(b2pt_ctx* ctx, b2pt_stream stream);
int b2pt_event_create(b2pt_ctx* ctx, b2pt_event* out_event);
int b2pt_event_destroy(b2pt_ctx* ctx, b2pt_event event);
int b2pt_event_record(b2pt_ctx* ctx, b2pt_event event, b2pt_stream stream);
int b2pt_event_synchronize(b2pt_ctx* ctx, b2pt_event event);
int b2pt_event_elapsed_ms(b2pt_ctx* ctx, b2pt_event start, b2pt_event stop, float* out_ms);
/* Runs fn(user) on a driver thread once everything queued on `stream` before it has finished. Replaces oroLaunchHostFunc
 * (src/Renderer/GPURendererThread.cpp:285,345: the "frame rendered" callbacks). fn must not call back into this library. */
typedef void (*b2pt_host_fn)(void* user);
int b2pt_stream_launch_host_func(b2pt_ctx* ctx, b2pt_stream stream, b2pt_host_fn fn, void* user);

/* LUT "textures". Replaces OrochiTexture / OrochiTexture3D + get_device_texture()
 * (src/HIPRT-Orochi/OrochiTexture.cpp, OrochiTexture3D.cpp; src/Renderer/GPURenderer.cpp:116-171). The returned
 * handle (a device pointer to a b2pt_lut descriptor) goes into the BRDFsData fields of HIPRTRenderData. */
int b2pt_lut_create(b2pt_ctx* ctx, const float* texels, int width, int height, int depth, int channels, void** out_handle);
int b2pt_lut_destroy(b2pt_ctx* ctx, void* handle);
/* Material textures. Replaces the scene's OrochiTexture objects and the array of texture objects RenderBuffers::material_textures
 * points to (src/HIPRT-Orochi/OrochiTexture.h, src/Renderer/GPURenderer.cpp:905-939). texels[i]: 8-bit, widths[i] x heights[i] x
 * channels[i] (1..4), row 0 first, as Image8Bit::read_image returns them. *out_array goes into RenderBuffers::material_textures;
 * material texture indices index it. Sampling is the reference CPU path's (nearest texel, repeat addressing, Image.cpp:161-195). */
int b2pt_textures_create(b2pt_ctx* ctx, const uint8_t* const* texels, const int* widths, const int* heights, const int* channels, int count, void** out_array);
int b2pt_textures_destroy(b2pt_ctx* ctx, void* array);

/* Geometry / BVH. Replaces HIPRTGeometry::upload_triangle_indices / upload_vertices_positions / build_bvh
 * (src/HIPRT-Orochi/HIPRTScene.h:37-72,87-150). The BVH is built on the HOST (binned SAH -> compressed 8-wide) and
 * uploaded; `b2pt_bvh_device_handle` is what goes into HIPRTRenderData::GPU_BVH. `flags` is reserved (0). */
typedef struct
{
    uint64_t node_count;
    uint64_t triangle_count;
    uint64_t bytes;
    int max_depth;
    double sah_cost;
    double build_seconds;
} b2pt_bvh_stats_t;

int b2pt_bvh_build(b2pt_ctx* ctx, const int* triangle_indices, int triangle_count, const float* vertex_positions, int vertex_count,
                   uint32_t flags, b2pt_stream stream, b2pt_bvh** out_bvh);
/* Re-sends the host-built node / triangle arrays to the existing device allocations (after a host-side refit, or to
 * measure the upload): replaces the re-upload half of HIPRTGeometry::build_bvh on a geometry change. */
int b2pt_bvh_reupload(b2pt_ctx* ctx, b2pt_bvh* bvh, b2pt_stream stream, uint64_t* out_bytes);
void* b2pt_bvh_device_handle(b2pt_bvh* bvh);
int b2pt_bvh_stats(b2pt_bvh* bvh, b2pt_bvh_stats_t* out);
int b2pt_bvh_destroy(b2pt_ctx* ctx, b2pt_bvh* bvh);

/* Kernels. Replaces GPUKernel::compile -> hiprtBuildTraceKernels (src/Compiler/GPUKernel.cpp:72-83,
 * src/HIPRT-Orochi/HIPRTOrochiUtils.cpp:118-164) by a lookup in a table of PRECOMPILED sm_100a variants, and
 * GPUKernel::launch_asynchronous -> oroModuleLaunchKernel (src/Compiler/GPUKernel.cpp:151-165).
 * Entry names served: "CameraRays", "MegaKernel", "ReSTIR_DI_LightsPresampling", "ReSTIR_Directional_Reuse_Compute"
 * (args = { &render_data, &masks_u, &masks_ull, &radius }, ReSTIRDIRenderPass.cpp:409-411), "ReSTIR_DI_InitialCandidates",
 * "ReSTIR_DI_TemporalReuse", "ReSTIR_DI_SpatialReuse", "ReSTIR_DI_SpatiotemporalReuse" (MegaKernelRenderPass.cpp:20,
 * FillGBufferRenderPass.cpp:23, ReSTIRDIRenderPass.cpp:21-29). "ReSTIR_DI_LightsPresampling" takes the reference's two
 * arguments, args = { &LightPresamplingParameters, &render_data } (ReSTIRDIRenderPass.cpp:441-444); kernels obtained with
 * {"ReSTIR_DI_DoLightPresampling", 1} draw their initial light candidates from render_data's presampled-light buffer.
 * Options are the reference's macro names with integer values, e.g. {"DirectLightSamplingStrategy", 3}; macros not
 * listed take the reference's defaults. {"ReSTIR_DI_LaterBouncesSamplingStrategy", 0..3} selects among four precompiled ReSTIR DI
 * MegaKernels (the other ReSTIR DI entries do not read it). Two extra options select the codegen flavour:
 *   {"B2PT_PRECISE", 1}  (default) IEEE fp32 without FMA contraction: the flavour parity is asserted on;
 *   {"B2PT_PRECISE", 0}  FMA contraction + fast intrinsics.
 *   {"B2PT_COUNTERS", 1} traversal-step counters (rays / nodes / triangles) accumulated into b2pt_counters_read.
 * and {"B2PT_ALPHA_TEST", 1} asks for the any-hit filter's alpha test (src/Device/functions/FilterFunction.h:47-74: a hit on a
 * material that is not fully opaque is accepted with probability alpha_opacity x base-colour-texture alpha, subject to
 * render_settings.do_alpha_testing / alpha_testing_indirect_bounce). A host sets it when its scene has such a material
 * (RenderBuffers::material_opaque holds a 0); scenes without one keep the shorter traversal. The random numbers of the test come
 * from a per-ray generator, not from the path's (see csrc/b2_bvh.h): images agree with the reference's statistically.
 * {"ReSTIR_DI_BiasCorrectionWeights", N} (ReSTIRDIOptions.h:9-16,120): "ReSTIR_DI_TemporalReuse", "ReSTIR_DI_SpatialReuse" and
 * "ReSTIR_DI_SpatiotemporalReuse" are served for N = 1 (1/Z), 2 (MIS-like), 3 (generalised balance heuristic), 4 (pairwise MIS, the
 * default), 5 (defensive pairwise MIS), 6 (symmetric ratio) and 7 (asymmetric ratio); 0 (1/M) for the temporal and the fused pass --
 * the reference's spatial pass does not compile with 1/M (SpatialReuse.h:237) and "ReSTIR_DI_SpatialReuse" answers B2PT_ERR_NO_VARIANT.
 * {"ReSTIR_DI_InitialTargetFunctionVisibility", 0|1}, {"ReSTIR_DI_SpatialTargetFunctionVisibility", 0|1}, {"ReSTIR_DI_DoVisibilityReuse", 0|1},
 * {"ReSTIR_DI_BiasCorrectionUseVisibility", 0|1} (ReSTIRDIOptions.h:50-84; defaults 0, 0, 1, 1): launch-time switches of the ReSTIR DI
 * passes; give the same values to every ReSTIR DI kernel of a renderer, as the reference's global compiler options do.
 * {"ReSTIR_DI_DoOptimalVisibilitySampling", 0|1} (:176): read by "ReSTIR_DI_SpatialReuse" (OptimalVisibilitySampling.h:17-46).
 * {"ReSTIR_DI_LightPresamplingStrategy", 0|1} (:162; LSS_BASE_UNIFORM / LSS_BASE_POWER, the default): how "ReSTIR_DI_LightsPresampling" picks
 * its emissive triangles, and the matching pdf in "ReSTIR_DI_InitialCandidates" -- launch-time as well. */
typedef struct
{
    const char* name;
    int value;
} b2pt_option;

int b2pt_kernel_get(b2pt_ctx* ctx, const char* entry_name, const b2pt_option* options, int option_count, b2pt_kernel** out_kernel);
/* Would b2pt_kernel_get find a variant? Same answer and same error text, without a context or a device: what a host asks before it
 * offers a combination of options (and what the CPU tests walk the option surface with). */
int b2pt_kernel_query(const char* entry_name, const b2pt_option* options, int option_count);
/* The macro names b2pt_kernel_get understands (static storage). Stands in for GPUKernel::parse_option_macros_used -> GPUKernelCompiler::
 * get_option_macros_used_by_kernel (src/Compiler/GPUKernel.cpp:176-180), which finds the macros a kernel reads by parsing its source: a
 * binding passes these and leaves the other entries of GPUKernelCompilerOptions::ALL_MACROS_NAMES (ReGIR, ReSTIR GI, GMoN ...) out. */
int b2pt_kernel_option_names(const char* const** out_names, int* out_count);
/* For a macro that is served at one value only (the reference's default): that value; B2PT_ERR_NO_VARIANT for the switched ones. Lets a
 * binding check at start-up that the defaults it was built against are the ones the kernels were built with (oracle/binding_check.cpp). */
int b2pt_kernel_option_fixed_value(const char* name, int* out_value);
/* args[0] must point to a host HIPRTRenderData / b2pt_render_data (2784 bytes, passed by value to the kernel).
 * block_x / block_y are accepted for interface compatibility; the B200 kernels choose their own thread->pixel map. */
int b2pt_kernel_launch(b2pt_kernel* kernel, int block_x, int block_y, int width, int height, void** args, b2pt_stream stream);
/* Multi-GPU form: launches the kernel over the pixel rows of `rank`'s bands only. The frame is partitioned into
 * interleaved bands of B2PT_BAND_ROWS rows; band b belongs to rank b % world (scene replicated on every GPU). The
 * reference has no multi-GPU path; this is the one addition to its launch surface (BASELINE.json north_star). */
#define B2PT_BAND_ROWS 16
int b2pt_kernel_launch_rows(b2pt_kernel* kernel, int width, int height, int rank, int world, void** args, b2pt_stream stream);
/* GPUKernel's destructor (src/Compiler/GPUKernel.h): releases the kernel object and its two timing events */
int b2pt_kernel_destroy(b2pt_kernel* kernel);
int b2pt_kernel_last_ms(b2pt_kernel* kernel, float* out_ms);
int b2pt_kernel_launch_count(b2pt_kernel* kernel, uint64_t* out_count);

typedef struct
{
    uint64_t rays;      /* closest-hit + any-hit traversals started */
    uint64_t nodes;     /* wide-BVH nodes fetched */
    uint64_t triangles; /* triangles tested */
} b2pt_counters_t;
/* Defensive fault words: traversals abandoned by the step guard, boundary-skip loops abandoned, traversal-stack
 * overflows. All three must stay 0 on a valid scene (asserted by tests/). */
typedef struct
{
    uint32_t traversal_guard_trips;
    uint32_t boundary_guard_trips;
    uint32_t stack_overflows;
} b2pt_faults_t;
int b2pt_faults_read(b2pt_ctx* ctx, b2pt_faults_t* out, int reset);
/* CUDA kernels launched so far by the passes of this library (a MegaKernel pass in its wavefront form is ~80 kernels) */
int b2pt_cuda_launch_count(b2pt_ctx* ctx, uint64_t* out_count);
/* Per-kernel CUDA-event timing of the wavefront launch sequences (MegaKernel and ReSTIR DI passes on scenes beyond L2), for
 * bench.py's roofline: kinds in the order of B2PT_WF_KIND_NAMES; out_launch_groups[k] = timed intervals of kind k (the three sort
 * kernels of a round count as one). Off by default (an event record per kernel when on). */
#define B2PT_WF_KIND_NAMES "trace,stage1,stage2,sort,begin,advance,pass_a,pass_b,pass_c,trace_pass"
#define B2PT_WF_KIND_COUNT 10
int b2pt_wavefront_profile(b2pt_ctx* ctx, int enabled);
int b2pt_wavefront_profile_read(b2pt_ctx* ctx, double* out_ms, uint64_t* out_launch_groups, int kinds);
int b2pt_counters_reset(b2pt_ctx* ctx);
int b2pt_counters_read(b2pt_ctx* ctx, b2pt_counters_t* out);

/* ------------------------------------------------------------------------------------------------------------
 * (1b) multi-GPU: frame partition + NCCL exchanges issued from C on the caller's stream (csrc/b2_comm.cu). No counterpart
 * in the reference, which drives one GPU; the launch surface stays GPUKernel::launch's plus a row selection.
 * ------------------------------------------------------------------------------------------------------------ */
typedef struct b2pt_comm b2pt_comm;
#define B2PT_COMM_ID_BYTES 128 /* sizeof(ncclUniqueId) */
/* Host logic: the frame rows (ascending) within `halo` rows of a band of `band_rows` rows that rank owns (band b -> rank
 * b % world); halo 0 = the owned rows. out_rows may be null (count only); it needs room for `height` ints. */
int b2pt_partition_rows(int height, int rank, int world, int band_rows, int halo, int* out_rows, int* out_count);
/* ncclGetUniqueId on one rank, handed to the others by the host's own bootstrap (torch.distributed in bench.py / tests),
 * then ncclCommInitRank on every rank. NCCL is loaded at run time (the libnccl.so.2 already in the process, else the system's). */
int b2pt_comm_unique_id(unsigned char out_id[B2PT_COMM_ID_BYTES]);
int b2pt_comm_create(b2pt_ctx* ctx, const unsigned char id[B2PT_COMM_ID_BYTES], int rank, int world, b2pt_comm** out_comm);
int b2pt_comm_destroy(b2pt_comm* comm);
int b2pt_comm_stats(b2pt_comm* comm, uint64_t* out_exchanges, uint64_t* out_gathers, uint64_t* out_bytes_sent);
/* `base` = `planes` planes, plane_stride_bytes apart, of `height` rows of row_bytes bytes. Halo exchange: every rank receives the
 * rows within `halo` of its bands from their owners and sends its own rows to the ranks that want them (pack kernel, one NCCL
 * group of send / recv, unpack kernel -- all on `stream`). All-gather: every rank's owned rows to every rank. */
int b2pt_comm_exchange_halo(b2pt_comm* comm, void* base, int planes, size_t plane_stride_bytes, size_t row_bytes, int height, int band_rows, int halo, b2pt_stream stream);
int b2pt_comm_allgather_rows(b2pt_comm* comm, void* base, int planes, size_t plane_stride_bytes, size_t row_bytes, int height, int band_rows, b2pt_stream stream);
/* b2pt_kernel_launch_rows with the band height chosen by the caller, or over an explicit ascending list of frame rows
 * (device_rows != null: a device array of n_rows rows, e.g. a rank's bands plus their halo rows). */
int b2pt_kernel_launch_partition(b2pt_kernel* kernel, int width, int height, int rank, int world, int band_rows, const int* device_rows, int n_rows, void** args,
                                 b2pt_stream stream);

/* ------------------------------------------------------------------------------------------------------------
 * (2) headless renderer (host mirror of GPURenderer, CPURenderer pass order)
 * ------------------------------------------------------------------------------------------------------------ */

/* variant: "c1" | "c2" | "c3" | "c4" | "c4p" (= c4 + ReSTIR DI light presampling) | "c4pu" (= c4p with uniform instead of power presampling) | "c4b0" "c4b1" "c4b2" "c4b3" "c4b5" "c4b6" "c4b7" (= c4 with
 * ReSTIR_DI_BiasCorrectionWeights 0 / 1 / 2 / 3 / 5 / 6 / 7; c4b0 has no spatial reuse pass) | "c4v1" .. "c4v5" (= c4 with the ReSTIR DI visibility macros away from their
 * defaults: InitialTargetFunctionVisibility on / SpatialTargetFunctionVisibility on / DoVisibilityReuse off / BiasCorrectionUseVisibility
 * off / the first three together) | "c4l0" "c4l1" "c4l2" (= c4 with ReSTIR_DI_LaterBouncesSamplingStrategy 0 / 1 / 2: uniform one light /
 * BSDF only / light + BSDF MIS at the bounces after the first) | "c4o" (= c4 with ReSTIR_DI_DoOptimalVisibilitySampling) | "c4x1" "c4x2" (= two combinations of the above, see b2_renderer.cpp)
 * (see b2_kernels_body.h); precise: 1 = parity flavour, 0 = fast flavour */
int b2pt_renderer_create(b2pt_ctx* ctx, int width, int height, const char* variant, int precise, b2pt_renderer** out_renderer);
int b2pt_renderer_destroy(b2pt_renderer* renderer);
/* Scene from a B2SCENE1 file (include/b2pt_scene_blob.h) or from host arrays laid out like the reference's Scene. */
int b2pt_renderer_set_scene_file(b2pt_renderer* renderer, const char* scene_blob_path);
int b2pt_renderer_set_scene_arrays(b2pt_renderer* renderer, const float* positions, const float* normals, const float* texcoords,
                                   const uint8_t* has_normals, int vertex_count, const int* indices, const int* material_indices,
                                   int triangle_count, const float* materials, int material_count, const float* camera52);
/* Host -> device re-send of every scene array + the BVH blob on the renderer's stream (GPURenderer::set_scene's upload
 * half, src/Renderer/GPURenderer.cpp:769-939, without the host BVH build). out_bytes: bytes copied. */
int b2pt_renderer_reupload_scene(b2pt_renderer* renderer, uint64_t* out_bytes);
/* GPURenderer::update_all_materials (src/Renderer/GPURenderer.cpp:1022-1040): new parameter records (B2MAT_FLOAT_COUNT floats each,
 * include/b2pt_scene_blob.h) for every material of the current scene; packed SoA, opacity, emissive list and power alias table are
 * rebuilt and re-uploaded, geometry and BVH stay. */
int b2pt_renderer_update_materials(b2pt_renderer* renderer, const float* materials, int material_count);
int b2pt_renderer_set_luts_file(b2pt_renderer* renderer, const char* lut_blob_path);
/* RGBA fp32 texels, row 0 first (what Image32Bit holds after read_image_hdr(path, 4, flipY=true)) */
int b2pt_renderer_set_envmap(b2pt_renderer* renderer, const float* rgba_texels, int width, int height);
/* Radiance .hdr (RGBE) file, decoded like Image32Bit::read_image_hdr(path, 4, flipY = true) (src/Image/Image.cpp, stb_image) */
int b2pt_renderer_set_envmap_hdr_file(b2pt_renderer* renderer, const char* hdr_path);
/* the decoder alone (host only): out_rgba == null reports the size, else it receives width * height * 4 floats */
int b2pt_image_read_hdr(const char* hdr_path, int flip_y, float* out_rgba, int* out_width, int* out_height);
/* Envmap rotation (RendererEnvmap::do_animation, src/Renderer/RendererEnvmap.cpp:129-181): the row-major 3x3 envmap->world and
 * world->envmap matrices of WorldSettings, as the reference host computes them (glm::orientate3 and its inverse). */
int b2pt_renderer_set_envmap_matrices(b2pt_renderer* renderer, const float* envmap_to_world9, const float* world_to_envmap9);
/* same, from a "B2ENVMP1" blob (int32 width, height, then width*height*4 floats): the container tests and bench use */
int b2pt_renderer_set_envmap_file(b2pt_renderer* renderer, const char* envmap_blob_path);
/* Host copy of the render data; edit settings (nb_bounces, restir_di_settings, ...) before rendering. */
b2pt_render_data* b2pt_renderer_render_data(b2pt_renderer* renderer);
int b2pt_renderer_set_host_seed(b2pt_renderer* renderer, unsigned int seed);
int b2pt_renderer_set_counters(b2pt_renderer* renderer, int enabled);
/* GPURenderer::get_global_compiler_options()->set_macro_value(name, value) followed by recompile_kernels() (GPURenderer.cpp:482-501): a kernel
 * option on top of the variant's -- any name of b2pt_kernel_option_names except the ones that define the variant (BSDFOverride,
 * DirectLightSamplingStrategy, DirectLightSamplingBaseStrategy, ReSTIR_DI_DoLightPresampling). B2PT_ERR_NO_VARIANT, at once, for a value that
 * is not served; the kernels are looked up again before the next render. */
int b2pt_renderer_set_macro(b2pt_renderer* renderer, const char* name, int value);
/* The variant names b2pt_renderer_create takes, and whether every kernel of one resolves (b2pt_kernel_query over its passes, with and
 * without traversal counters) in a flavour. No device needed. */
int b2pt_renderer_variant_count(void);
const char* b2pt_renderer_variant_name(int index);
int b2pt_renderer_variant_query(const char* variant, int precise);
/* When enabled, every pass launch is followed by an event wait and its time is added to b2pt_renderer_pass_ms
 * (serialises host and device: for profiling runs only, never for the timed bench region). */
int b2pt_renderer_set_profile_passes(b2pt_renderer* renderer, int enabled);
int b2pt_renderer_reset(b2pt_renderer* renderer);
/* Renders `samples` samples of every pixel (all passes), asynchronously on the renderer's stream. */
int b2pt_renderer_render(b2pt_renderer* renderer, int samples);
/* Multi-GPU (BASELINE.json north_star; the reference itself is single-GPU): this renderer renders only the pixel rows of its
 * interleaved bands (band b -> rank b % world; bands of at most 16 rows, at most 32 for the ReSTIR DI variants, the exact height chosen so that
 * every rank owns the same number of rows -- b2pt_renderer_partition_info reports it); scene, BVH and LUTs are
 * replicated by giving every rank the same scene. Without ReSTIR DI no pass reads another pixel. With ReSTIR DI the reuse passes
 * read reservoirs and G-buffer entries up to reuse_radius + neighbor_search_radius pixels away: CameraRays also covers those
 * halo rows (deterministic, so neighbours' G-buffer entries are recomputed instead of exchanged) and the reservoir rows of the
 * halo are exchanged with the ranks that own them (b2pt_comm_exchange_halo: ncclSend / ncclRecv on the renderer's stream) right
 * before every pass that reads neighbours. `dynamic_camera` != 0: the camera may move between frames, so a temporal neighbour can
 * be anywhere: the halo becomes the whole frame. comm may be null when world == 1. */
int b2pt_renderer_set_partition(b2pt_renderer* renderer, int rank, int world, b2pt_comm* comm, int dynamic_camera);
/* All-gathers the framebuffer rows of every rank (ncclAllGather on the renderer's stream) so that each rank holds the whole
 * frame; counted in b2pt_renderer_last_render_ms. */
int b2pt_renderer_gather_framebuffer(b2pt_renderer* renderer);
/* (pixels CameraRays covers on this rank, pixels this rank owns, rows per band, halo rows) of the current partition */
int b2pt_renderer_partition_info(b2pt_renderer* renderer, uint64_t* out_camera_pixels, uint64_t* out_owned_pixels, int* out_band_rows, int* out_halo_rows);
/* GPURenderer::set_camera (src/Renderer/GPURenderer.h): the camera of the next frames as 52 floats (B2SceneCamera: inverse view,
 * inverse projection, view-projection, position, vertical fov). The previous camera becomes prev_camera at the next sample, as
 * in update_render_data (CPURenderer.cpp:600-604). */
int b2pt_renderer_set_camera(b2pt_renderer* renderer, const float* camera52);
/* offsetof(HIPRTRenderData, buffers.accumulated_ray_colors): where the framebuffer's device pointer sits in the host
 * copy returned by b2pt_renderer_render_data (bindings that wrap the framebuffer without a copy) */
int b2pt_render_data_framebuffer_offset(void);
int b2pt_renderer_synchronize(b2pt_renderer* renderer);
/* Device time of the last b2pt_renderer_render call (CUDA events on the renderer's stream), after synchronize. */
int b2pt_renderer_last_render_ms(b2pt_renderer* renderer, float* out_ms);
int b2pt_renderer_pass_ms(b2pt_renderer* renderer, int pass, float* out_ms, uint64_t* out_launches);
/* Result buffers -> host. framebuffer: width*height*3 floats (sum of samples, like CPURenderer::get_framebuffer). */
int b2pt_renderer_download_framebuffer(b2pt_renderer* renderer, float* out_rgb);
int b2pt_renderer_download_gbuffer(b2pt_renderer* renderer, int* out_prim, float* out_position, uint32_t* out_geometric_normal,
                                   uint32_t* out_shading_normal, uint32_t* out_material27);
int b2pt_renderer_download_reservoirs(b2pt_renderer* renderer, int which, uint32_t* out_words9);
int b2pt_renderer_scene_info(b2pt_renderer* renderer, b2pt_bvh_stats_t* out_bvh, int* out_triangles, int* out_emissive_triangles);
/* Derived host arrays, for bit-exact comparison with the oracle (packed materials: 33 words each). */
int b2pt_renderer_download_derived(b2pt_renderer* renderer, uint32_t* out_packed_materials33, float* out_triangle_areas,
                                   int* out_emissive_triangles, float* out_alias_probas, int* out_alias_alias);

#ifdef __cplusplus
}
#endif

#endif
