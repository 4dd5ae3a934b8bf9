/*
 * b2_capi.cu -- device layer of the C ABI (include/b2pt.h, group 1): context, memory, streams/events, LUTs,
 * BVH build + upload, precompiled-kernel lookup and launch. Thin by design: it is what replaces
 * src/HIPRT-Orochi and the JIT half of src/Compiler in the reference.
 */
#include "b2pt.h"

#include "b2_bvh_build.h"
#include "b2_kernel_table.h"

#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <string>
#include <unordered_map>
#include <vector>

static thread_local std::string g_last_error;

void b2_set_error(const char* fmt, ...)
{
    char buf[1024];
    va_list ap;
    va_start(ap, fmt);
    std::vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    g_last_error = buf;
}

#define B2_CUDA(call)                                                                                  \
    do                                                                                                 \
    {                                                                                                  \
        cudaError_t e_ = (call);                                                                       \
        if (e_ != cudaSuccess)                                                                         \
        {                                                                                              \
            b2_set_error("%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__);   \
            return B2PT_ERR_CUDA;                                                                      \
        }                                                                                              \
    } while (0)

struct b2pt_ctx
{
    int device = 0;
    cudaDeviceProp props;
};

/* every entry that allocates, copies or launches runs on the context's own device, whichever device another context of the
 * process made current last */
#define B2_SELECT_DEVICE(ctx)                     \
    do                                            \
    {                                             \
        if (ctx)                                  \
            B2_CUDA(cudaSetDevice((ctx)->device)); \
    } while (0)

struct b2pt_bvh
{
    B2Node* nodes = nullptr;
    B2Tri* tris = nullptr;
    B2BvhDevice* descriptor = nullptr;
    b2pt_bvh_stats_t stats;
    B2BvhHost host; /* kept for b2pt_bvh_reupload */
    bool host_pinned = false; /* page-locked on the first re-upload */
};

/* device descriptor -> depth and size of the tree, so a launch can size the traversal stack and pick the MegaKernel form
 * (b2pt_kernel_launch only sees the render data) */
static std::mutex g_bvh_depth_mutex;
static std::unordered_map<const void*, int> g_bvh_depth;
static std::unordered_map<const void*, uint64_t> g_bvh_nodes;

/* kernel units (objects of b2_kernels.cu) register here from their static initialisers */
static std::vector<B2KernelUnit>& kernel_units()
{
    static std::vector<B2KernelUnit> units;
    return units;
}
/* ---- per-kernel timing of the wavefront sequences (b2_kernel_table.h) ---- */
namespace
{
struct WfProfiler
{
    std::mutex mutex;
    bool enabled = false;
    std::vector<cudaEvent_t> events; /* pool, grown on demand, reused after every read */
    std::vector<int> kinds;          /* kinds[i] = what was launched after events[i] (B2_WF_K_END closes a sequence) */
    double ms[B2_WF_K_COUNT] = {};
    unsigned long long marks[B2_WF_K_COUNT] = {};
};
WfProfiler& wf_profiler()
{
    static WfProfiler p;
    return p;
}
/* folds the recorded marks into the totals; the caller has synchronised the device */
void wf_profiler_fold(WfProfiler& p)
{
    for (size_t i = 0; i + 1 < p.kinds.size(); i++)
    {
        if (p.kinds[i] == B2_WF_K_END)
            continue;
        float elapsed = 0.0f;
        if (cudaEventElapsedTime(&elapsed, p.events[i], p.events[i + 1]) == cudaSuccess)
        {
            p.ms[p.kinds[i]] += elapsed;
            p.marks[p.kinds[i]] += 1;
        }
    }
    p.kinds.clear();
}
} // namespace

bool b2_wfprof_enabled() { return wf_profiler().enabled; }

void b2_wfprof_mark(cudaStream_t stream, int kind)
{
    WfProfiler& p = wf_profiler();
    std::lock_guard<std::mutex> lock(p.mutex);
    if (p.kinds.size() >= p.events.size())
    {
        cudaEvent_t e = nullptr;
        if (cudaEventCreate(&e) != cudaSuccess)
            return;
        p.events.push_back(e);
    }
    if (cudaEventRecord(p.events[p.kinds.size()], stream) == cudaSuccess)
        p.kinds.push_back(kind);
}

void b2_register_kernel_unit(const B2KernelUnit* unit) { kernel_units().push_back(*unit); }

struct b2pt_kernel
{
    b2pt_ctx* ctx = nullptr;
    const B2KernelEntry* entry = nullptr;
    cudaEvent_t start = nullptr, stop = nullptr;
    bool timed = false;
    uint64_t launches = 0;
    int light_presampling = 0; /* ReSTIR_DI_DoLightPresampling the kernel was asked for */
    int precise = 1;           /* codegen flavour of the unit the entry comes from */
    int alpha_test = 0;        /* B2PT_ALPHA_TEST: candidate hits on non-opaque materials go through the alpha test */
    int bias_correction = 4;   /* ReSTIR_DI_BiasCorrectionWeights */
    int restir_options = 4 | 8; /* B2_DI_OPT_* bits: the four ReSTIR DI visibility macros, at the reference's defaults */
};

int b2_ctx_device(b2pt_ctx* ctx) { return ctx ? ctx->device : 0; }

extern "C" {

const char* b2pt_last_error_string(void) { return g_last_error.c_str(); }
const char* b2pt_version_string(void) { return "b2pt 0.1 (sm_100a)"; }

int b2pt_ctx_create(int device_index, b2pt_ctx** out_ctx)
{
    if (!out_ctx)
    {
        b2_set_error("b2pt_ctx_create: null output");
        return B2PT_ERR_INVALID_ARGUMENT;
    }
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count == 0)
    {
        b2_set_error("no CUDA device available (%s); libb2pt has no CPU fallback", e == cudaSuccess ? "device count is 0" : cudaGetErrorString(e));
        return B2PT_ERR_NO_DEVICE;
    }
    if (device_index < 0 || device_index >= count)
    {
        b2_set_error("device index %d out of range (%d devices)", device_index, count);
        return B2PT_ERR_INVALID_ARGUMENT;
    }
    B2_CUDA(cudaSetDevice(device_index));
    b2pt_ctx* ctx = new b2pt_ctx();
    ctx->device = device_index;
    B2_CUDA(cudaGetDeviceProperties(&ctx->props, device_index));
    *out_ctx = ctx;
    return B2PT_OK;
}

int b2pt_ctx_destroy(b2pt_ctx* ctx)
{
    delete ctx;
    return B2PT_OK;
}

int b2pt_ctx_make_current(b2pt_ctx* ctx)
{
    if (!ctx) { b2_set_error("null context"); return B2PT_ERR_INVALID_ARGUMENT; }
    B2_CUDA(cudaSetDevice(ctx->device));
    return B2PT_OK;
}

int b2pt_ctx_device_props(b2pt_ctx* ctx, b2pt_device_props_t* out)
{
    if (!ctx || !out) { b2_set_error("null argument"); return B2PT_ERR_INVALID_ARGUMENT; }
    std::memset(out, 0, sizeof(*out));
    std::strncpy(out->name, ctx->props.name, sizeof(out->name) - 1);
    out->sm_count = ctx->props.multiProcessorCount;
    out->compute_major = ctx->props.major;
    out->compute_minor = ctx->props.minor;
    out->total_memory_bytes = ctx->props.totalGlobalMem;
    out->l2_bytes = (size_t)ctx->props.l2CacheSize;
    out->shared_memory_per_block_optin = ctx->props.sharedMemPerBlockOptin;
    int khz = 0;
    cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, ctx->device);
    out->clock_khz = khz;
    return B2PT_OK;
}

int b2pt_malloc(b2pt_ctx* ctx, size_t bytes, void** out_device_ptr)
{
    B2_SELECT_DEVICE(ctx);
    if (!ctx || !out_device_ptr) { b2_set_error("null argument"); return B2PT_ERR_INVALID_ARGUMENT; }
    *out_device_ptr = nullptr;
    if (bytes == 0)
        return B2PT_OK;
    B2_CUDA(cudaMalloc(out_device_ptr, bytes));
    return B2PT_OK;
}

int b2pt_free(b2pt_ctx* ctx, void* device_ptr)
{
    B2_SELECT_DEVICE(ctx);
    if (device_ptr)
        B2_CUDA(cudaFree(device_ptr));
    return B2PT_OK;
}

int b2pt_memcpy_h2d(b2pt_ctx* ctx, void* dst, const void* src, size_t bytes, b2pt_stream stream)
{
    B2_SELECT_DEVICE(ctx);
    if (bytes == 0) return B2PT_OK;
    if (stream) B2_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, (cudaStream_t)stream));
    else B2_CUDA(cudaMemcpy(dst, src, bytes, cudaMemcpyHostToDevice));
    return B2PT_OK;
}

int b2pt_memcpy_d2h(b2pt_ctx* ctx, void* dst, const void* src, size_t bytes, b2pt_stream stream)
{
    B2_SELECT_DEVICE(ctx);
    if (bytes == 0) return B2PT_OK;
    if (stream) B2_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, (cudaStream_t)stream));
    else B2_CUDA(cudaMemcpy(dst, src, bytes, cudaMemcpyDeviceToHost));
    return B2PT_OK;
}

int b2pt_memcpy_d2d(b2pt_ctx* ctx, void* dst, const void* src, size_t bytes, b2pt_stream stream)
{
    B2_SELECT_DEVICE(ctx);
    if (bytes == 0) return B2PT_OK;
    if (stream) B2_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToDevice, (cudaStream_t)stream));
    else B2_CUDA(cudaMemcpy(dst, src, bytes, cudaMemcpyDeviceToDevice));
    return B2PT_OK;
}

int b2pt_memset(b2pt_ctx* ctx, void* device_ptr, int byte_value, size_t bytes, b2pt_stream stream)
{
    B2_SELECT_DEVICE(ctx);
    if (bytes == 0) return B2PT_OK;
    if (stream) B2_CUDA(cudaMemsetAsync(device_ptr, byte_value, bytes, (cudaStream_t)stream));
    else B2_CUDA(cudaMemset(device_ptr, byte_value, bytes));
    return B2PT_OK;
}

int b2pt_host_alloc_pinned(b2pt_ctx* ctx, size_t bytes, void** out_host_ptr)
{
    B2_SELECT_DEVICE(ctx);
    if (!out_host_ptr) { b2_set_error("null argument"); return B2PT_ERR_INVALID_ARGUMENT; }
    B2_CUDA(cudaMallocHost(out_host_ptr, bytes));
    return B2PT_OK;
}

int b2pt_host_free_pinned(b2pt_ctx* ctx, void* host_ptr)
{
    B2_SELECT_DEVICE(ctx);
    if (host_ptr)
        B2_CUDA(cudaFreeHost(host_ptr));
    return B2PT_OK;
}

int b2pt_host_register(b2pt_ctx* ctx, const void* host_ptr, size_t bytes)
{
    B2_SELECT_DEVICE(ctx);
    if (!host_ptr || !bytes) { b2_set_error("null argument"); return B2PT_ERR_INVALID_ARGUMENT; }
    B2_CUDA(cudaHostRegister(const_cast<void*>(host_ptr), bytes, cudaHostRegisterDefault));
    return B2PT_OK;
}

int b2pt_host_unregister(b2pt_ctx* ctx, const void* host_ptr)
{
    B2_SELECT_DEVICE(ctx);
    if (host_ptr)
        B2_CUDA(cudaHostUnregister(const_cast<void*>(host_ptr)));
    return B2PT_OK;
}

int b2pt_stream_create(b2pt_ctx* ctx, b2pt_stream* out_stream)
{
    B2_SELECT_DEVICE(ctx);
    cudaStream_t s;
    /* a blocking stream, like oroStreamCreate (cuStreamCreate, flags 0): work on it is ordered after the synchronous
     * uploads / memsets the host issues on the legacy default stream (GPURenderer relies on that ordering) */
    B2_CUDA(cudaStreamCreateWithFlags(&s, cudaStreamDefault));
    *out_stream = s;
    return B2PT_OK;
}
int b2pt_stream_destroy(b2pt_ctx*, b2pt_stream stream)
{
    if (stream) B2_CUDA(cudaStreamDestroy((cudaStream_t)stream));
    return B2PT_OK;
}
int b2pt_stream_synchronize(b2pt_ctx*, b2pt_stream stream)
{
    B2_CUDA(cudaStreamSynchronize((cudaStream_t)stream));
    return B2PT_OK;
}
int b2pt_stream_launch_host_func(b2pt_ctx*, b2pt_stream stream, b2pt_host_fn fn, void* user)
{
    if (!fn) { b2_set_error("b2pt_stream_launch_host_func: null function"); return B2PT_ERR_INVALID_ARGUMENT; }
    B2_CUDA(cudaLaunchHostFunc((cudaStream_t)stream, fn, user));
    return B2PT_OK;
}
int b2pt_event_create(b2pt_ctx* ctx, b2pt_event* out_event)
{
    B2_SELECT_DEVICE(ctx);
    cudaEvent_t e;
    B2_CUDA(cudaEventCreate(&e));
    *out_event = e;
    return B2PT_OK;
}
int b2pt_event_destroy(b2pt_ctx*, b2pt_event event)
{
    if (event) B2_CUDA(cudaEventDestroy((cudaEvent_t)event));
    return B2PT_OK;
}
int b2pt_event_record(b2pt_ctx*, b2pt_event event, b2pt_stream stream)
{
    B2_CUDA(cudaEventRecord((cudaEvent_t)event, (cudaStream_t)stream));
    return B2PT_OK;
}
int b2pt_event_synchronize(b2pt_ctx*, b2pt_event event)
{
    B2_CUDA(cudaEventSynchronize((cudaEvent_t)event));
    return B2PT_OK;
}
int b2pt_event_elapsed_ms(b2pt_ctx*, b2pt_event start, b2pt_event stop, float* out_ms)
{
    B2_CUDA(cudaEventElapsedTime(out_ms, (cudaEvent_t)start, (cudaEvent_t)stop));
    return B2PT_OK;
}

int b2pt_lut_create(b2pt_ctx* ctx, const float* texels, int width, int height, int depth, int channels, void** out_handle)
{
    B2_SELECT_DEVICE(ctx);
    if (!ctx || !texels || !out_handle || width <= 0 || height <= 0 || depth <= 0 || channels <= 0)
    {
        b2_set_error("b2pt_lut_create: invalid argument");
        return B2PT_ERR_INVALID_ARGUMENT;
    }
    /* descriptor and texels share one allocation: [b2pt_lut (padded to 256 B)][texels] */
    size_t texel_bytes = (size_t)width * height * depth * channels * sizeof(float);
    char* block = nullptr;
    B2_CUDA(cudaMalloc((void**)&block, 256 + texel_bytes));
    b2pt_lut desc;
    desc.texels = reinterpret_cast<const float*>(block + 256);
    desc.width = width;
    desc.height = height;
    desc.depth = depth;
    desc.channels = channels;
    B2_CUDA(cudaMemcpy(block, &desc, sizeof(desc), cudaMemcpyHostToDevice));
    B2_CUDA(cudaMemcpy(block + 256, texels, texel_bytes, cudaMemcpyHostToDevice));
    *out_handle = block;
    return B2PT_OK;
}

int b2pt_lut_destroy(b2pt_ctx* ctx, void* handle)
{
    B2_SELECT_DEVICE(ctx);
    if (handle) B2_CUDA(cudaFree(handle));
    return B2PT_OK;
}

/* Material textures. Replaces the std::vector<OrochiTexture> of the reference's scene + the array of texture objects it uploads
 * (GPURenderer::set_scene, GPURenderer.cpp:905-939): one allocation = [descriptors][sRGB decode table][texels...]. */
int b2pt_textures_create(b2pt_ctx* ctx, const uint8_t* const* texels, const int* widths, const int* heights, const int* channels, int count, void** out_array)
{
    B2_SELECT_DEVICE(ctx);
    if (!ctx || !out_array || count < 0 || (count > 0 && (!texels || !widths || !heights || !channels)))
    {
        b2_set_error("b2pt_textures_create: invalid argument");
        return B2PT_ERR_INVALID_ARGUMENT;
    }
    *out_array = nullptr;
    if (count == 0)
        return B2PT_OK;
    struct Descriptor { const uint8_t* texels; int width, height, channels; const float* srgb_to_linear; };
    static_assert(sizeof(Descriptor) == 32, "b2pt_texture layout (b2_material.h)");
    size_t offset = ((size_t)count * sizeof(Descriptor) + 255) & ~(size_t)255;
    const size_t table_offset = offset;
    offset += 256 * sizeof(float);
    std::vector<size_t> texel_offsets((size_t)count);
    for (int i = 0; i < count; i++)
    {
        if (!texels[i] || widths[i] <= 0 || heights[i] <= 0 || channels[i] < 1 || channels[i] > 4)
        {
            b2_set_error("b2pt_textures_create: texture %d is %dx%d with %d channels", i, widths[i], heights[i], channels[i]);
            return B2PT_ERR_INVALID_ARGUMENT;
        }
        texel_offsets[i] = offset;
        offset += ((size_t)widths[i] * heights[i] * channels[i] + 255) & ~(size_t)255;
    }
    char* block = nullptr;
    B2_CUDA(cudaMalloc((void**)&block, offset));
    std::vector<Descriptor> descriptors((size_t)count);
    for (int i = 0; i < count; i++)
    {
        descriptors[i].texels = reinterpret_cast<const uint8_t*>(block + texel_offsets[i]);
        descriptors[i].width = widths[i];
        descriptors[i].height = heights[i];
        descriptors[i].channels = channels[i];
        descriptors[i].srgb_to_linear = reinterpret_cast<const float*>(block + table_offset);
        B2_CUDA(cudaMemcpy(block + texel_offsets[i], texels[i], (size_t)widths[i] * heights[i] * channels[i], cudaMemcpyHostToDevice));
    }
    /* sample_texture_rgba's sRGB decode, intrin_pow(texel / 255, 2.2) (Texture.h:62-65): the host C library's powf on the 256 texel values */
    float table[256];
    for (int v = 0; v < 256; v++)
        table[v] = powf((float)v / 255.0f, 2.2f);
    B2_CUDA(cudaMemcpy(block + table_offset, table, sizeof(table), cudaMemcpyHostToDevice));
    B2_CUDA(cudaMemcpy(block, descriptors.data(), descriptors.size() * sizeof(Descriptor), cudaMemcpyHostToDevice));
    *out_array = block;
    return B2PT_OK;
}

int b2pt_textures_destroy(b2pt_ctx* ctx, void* array)
{
    B2_SELECT_DEVICE(ctx);
    if (array) B2_CUDA(cudaFree(array));
    return B2PT_OK;
}

int b2pt_bvh_build(b2pt_ctx* ctx, const int* triangle_indices, int triangle_count, const float* vertex_positions, int vertex_count,
                   uint32_t, b2pt_stream stream, b2pt_bvh** out_bvh)
{
    B2_SELECT_DEVICE(ctx);
    if (!ctx || !out_bvh || triangle_count < 0 || vertex_count < 0 || (triangle_count > 0 && (!triangle_indices || !vertex_positions)))
    {
        b2_set_error("b2pt_bvh_build: invalid argument");
        return B2PT_ERR_INVALID_ARGUMENT;
    }
    auto t0 = std::chrono::high_resolution_clock::now();
    b2pt_bvh* bvh = new b2pt_bvh();
    B2BvhHost& host = bvh->host;
    std::string error;
    bool built = false;
    try
    {
        built = b2_build_cw8(triangle_indices, (size_t)triangle_count, vertex_positions, (size_t)vertex_count, host, error);
    }
    catch (const std::exception& e)
    {
        error = std::string("host exception: ") + e.what();
    }
    if (!built)
    {
        delete bvh;
        b2_set_error("b2pt_bvh_build: %s", error.c_str());
        return B2PT_ERR_UNSUPPORTED_SCENE;
    }
    cudaStream_t s = (cudaStream_t)stream;
    if (!host.nodes.empty())
    {
        B2_CUDA(cudaMalloc((void**)&bvh->nodes, host.nodes.size() * sizeof(B2Node)));
        B2_CUDA(cudaMemcpyAsync(bvh->nodes, host.nodes.data(), host.nodes.size() * sizeof(B2Node), cudaMemcpyHostToDevice, s));
    }
    if (!host.tris.empty())
    {
        B2_CUDA(cudaMalloc((void**)&bvh->tris, host.tris.size() * sizeof(B2Tri)));
        B2_CUDA(cudaMemcpyAsync(bvh->tris, host.tris.data(), host.tris.size() * sizeof(B2Tri), cudaMemcpyHostToDevice, s));
    }
    B2BvhDevice desc;
    desc.nodes = bvh->nodes;
    desc.tris = bvh->tris;
    desc.node_count = (uint32_t)host.nodes.size();
    desc.tri_count = (uint32_t)host.tris.size();
    B2_CUDA(cudaMalloc((void**)&bvh->descriptor, sizeof(B2BvhDevice)));
    B2_CUDA(cudaMemcpyAsync(bvh->descriptor, &desc, sizeof(desc), cudaMemcpyHostToDevice, s));
    B2_CUDA(cudaStreamSynchronize(s)); /* the host vectors die at return */
    bvh->stats.node_count = host.stats.node_count;
    bvh->stats.triangle_count = host.stats.triangle_count;
    bvh->stats.bytes = host.stats.bytes;
    bvh->stats.max_depth = host.stats.max_depth;
    bvh->stats.sah_cost = host.stats.sah_cost;
    bvh->stats.build_seconds = std::chrono::duration<double>(std::chrono::high_resolution_clock::now() - t0).count();
    {
        std::lock_guard<std::mutex> lock(g_bvh_depth_mutex);
        g_bvh_depth[bvh->descriptor] = bvh->stats.max_depth;
        g_bvh_nodes[bvh->descriptor] = bvh->stats.node_count;
    }
    *out_bvh = bvh;
    return B2PT_OK;
}

int b2pt_bvh_reupload(b2pt_ctx* ctx, b2pt_bvh* bvh, b2pt_stream stream, uint64_t* out_bytes)
{
    B2_SELECT_DEVICE(ctx);
    if (!bvh) { b2_set_error("null argument"); return B2PT_ERR_INVALID_ARGUMENT; }
    cudaStream_t s = (cudaStream_t)stream;
    if (!bvh->host_pinned)
    {
        /* best effort: a range that cannot be page-locked is still copied, just through the driver's staging buffer */
        bool ok = true;
        if (!bvh->host.nodes.empty())
            ok = cudaHostRegister(bvh->host.nodes.data(), bvh->host.nodes.size() * sizeof(B2Node), cudaHostRegisterDefault) == cudaSuccess;
        if (ok && !bvh->host.tris.empty() && cudaHostRegister(bvh->host.tris.data(), bvh->host.tris.size() * sizeof(B2Tri), cudaHostRegisterDefault) != cudaSuccess)
        {
            ok = false;
            if (!bvh->host.nodes.empty())
                cudaHostUnregister(bvh->host.nodes.data());
        }
        if (!ok)
            (void)cudaGetLastError();
        bvh->host_pinned = ok;
    }
    if (!bvh->host.nodes.empty())
        B2_CUDA(cudaMemcpyAsync(bvh->nodes, bvh->host.nodes.data(), bvh->host.nodes.size() * sizeof(B2Node), cudaMemcpyHostToDevice, s));
    if (!bvh->host.tris.empty())
        B2_CUDA(cudaMemcpyAsync(bvh->tris, bvh->host.tris.data(), bvh->host.tris.size() * sizeof(B2Tri), cudaMemcpyHostToDevice, s));
    if (out_bytes) *out_bytes = bvh->host.nodes.size() * sizeof(B2Node) + bvh->host.tris.size() * sizeof(B2Tri);
    return B2PT_OK;
}

void* b2pt_bvh_device_handle(b2pt_bvh* bvh) { return bvh ? bvh->descriptor : nullptr; }

int b2pt_bvh_stats(b2pt_bvh* bvh, b2pt_bvh_stats_t* out)
{
    if (!bvh || !out) { b2_set_error("null argument"); return B2PT_ERR_INVALID_ARGUMENT; }
    *out = bvh->stats;
    return B2PT_OK;
}

int b2pt_bvh_destroy(b2pt_ctx* ctx, b2pt_bvh* bvh)
{
    B2_SELECT_DEVICE(ctx);
    if (!bvh) return B2PT_OK;
    {
        std::lock_guard<std::mutex> lock(g_bvh_depth_mutex);
        g_bvh_depth.erase(bvh->descriptor);
        g_bvh_nodes.erase(bvh->descriptor);
    }
    if (bvh->host_pinned)
    {
        if (!bvh->host.nodes.empty()) cudaHostUnregister(bvh->host.nodes.data());
        if (!bvh->host.tris.empty()) cudaHostUnregister(bvh->host.tris.data());
    }
    cudaFree(bvh->nodes);
    cudaFree(bvh->tris);
    cudaFree(bvh->descriptor);
    delete bvh;
    return B2PT_OK;
}

/* ---- kernels ---- */

struct MacroDefault { const char* name; int value; };
/* macros a caller may pass with their reference default; any other value has no variant */
static const MacroDefault kFixedMacros[] = {
    { "DirectLightSamplingNEESampleCount", 1 }, { "DirectLightSamplingDeltaDistributionOptimization", 1 },
    { "DirectLightSamplingAllowBackfacingLights", 0 }, { "TrianglePointSamplingStrategy", 1 }, { "EnvmapSamplingDoBSDFMIS", 1 },
    { "EnvmapSamplingDoBilinearFiltering", 0 }, { "PathSamplingStrategy", 0 }, { "RISUseVisiblityTargetFunction", 0 },
    { "NestedDielectricsStackSize", 4 }, { "UseSharedStackBVHTraversal", 1 }, { "KernelWorkgroupThreadCount", 64 },
    { "PrincipledBSDFDiffuseLobe", 0 }, { "PrincipledBSDFAnisotropicGGXSampleFunction", 0 }, { "PrincipledBSDFDoEnergyCompensation", 1 },
    { "PrincipledBSDFDoGlassEnergyCompensation", 1 }, { "PrincipledBSDFDoClearcoatEnergyCompensation", 1 },
    { "PrincipledBSDFDoMetallicEnergyCompensation", 1 }, { "PrincipledBSDFDoMetallicFresnelEnergyCompensation", 1 },
    { "PrincipledBSDFDoSpecularEnergyCompensation", 1 }, { "PrincipledBSDFDeltaDistributionEvaluationOptimization", 1 },
    { "PrincipledBSDFSampleGlossyBasedOnFresnel", 0 }, { "PrincipledBSDFSampleCoatBasedOnFresnel", 0 },
    { "PrincipledBSDFDoMicrofacetRegularization", 1 }, { "PrincipledBSDFDoMicrofacetRegularizationConsistentParameterization", 1 },
    { "PrincipledBSDFMicrofacetRegularizationDiffusionHeuristic", 1 }, 
    { "DirectLightUseNEEPlusPlus", 0 }, { "SharedStackBVHTraversalSize", 16 },
};

static int pass_of_entry_name(const char* name)
{
    if (!std::strcmp(name, "CameraRays")) return B2_PASS_CAMERA_RAYS;
    if (!std::strcmp(name, "MegaKernel")) return B2_PASS_MEGAKERNEL;
    if (!std::strcmp(name, "ReSTIR_DI_LightsPresampling")) return B2_PASS_RESTIR_DI_PRESAMPLING;
    if (!std::strcmp(name, "ReSTIR_Directional_Reuse_Compute")) return B2_PASS_RESTIR_DIRECTIONAL_REUSE;
    if (!std::strcmp(name, "ReSTIR_DI_InitialCandidates")) return B2_PASS_RESTIR_DI_INITIAL_CANDIDATES;
    if (!std::strcmp(name, "ReSTIR_DI_TemporalReuse")) return B2_PASS_RESTIR_DI_TEMPORAL;
    if (!std::strcmp(name, "ReSTIR_DI_SpatialReuse")) return B2_PASS_RESTIR_DI_SPATIAL;
    if (!std::strcmp(name, "ReSTIR_DI_SpatiotemporalReuse")) return B2_PASS_RESTIR_DI_SPATIOTEMPORAL;
    return -1;
}

/* The macro names b2pt_kernel_get understands: the ones it switches on (below) and the ones it only accepts at the reference's default.
 * A binding of the reference's GPUKernel passes these and leaves the rest of GPUKernelCompilerOptions::ALL_MACROS_NAMES out, the way the
 * reference passes a kernel only the macros its source reads (GPUKernel::parse_option_macros_used). */
static const char* const kSwitchedMacros[] = {
    "BSDFOverride", "DirectLightSamplingStrategy", "DirectLightSamplingBaseStrategy", "ReSTIR_DI_LaterBouncesSamplingStrategy", "EnvmapSamplingStrategy",
    "ReSTIR_DI_DoLightPresampling", "ReSTIR_DI_LightPresamplingStrategy", "ReSTIR_DI_BiasCorrectionWeights", "ReSTIR_DI_InitialTargetFunctionVisibility",
    "ReSTIR_DI_SpatialTargetFunctionVisibility", "ReSTIR_DI_DoVisibilityReuse", "ReSTIR_DI_BiasCorrectionUseVisibility", "ReSTIR_DI_DoOptimalVisibilitySampling",
};

int b2pt_kernel_option_names(const char* const** out_names, int* out_count)
{
    if (!out_names || !out_count) { b2_set_error("null argument"); return B2PT_ERR_INVALID_ARGUMENT; }
    static std::vector<const char*> names;
    static std::once_flag once;
    std::call_once(once, [] {
        for (const char* n : kSwitchedMacros)
            names.push_back(n);
        for (const MacroDefault& m : kFixedMacros)
            names.push_back(m.name);
    });
    *out_names = names.data();
    *out_count = (int)names.size();
    return B2PT_OK;
}

int b2pt_kernel_option_fixed_value(const char* name, int* out_value)
{
    if (!name || !out_value) { b2_set_error("null argument"); return B2PT_ERR_INVALID_ARGUMENT; }
    for (const MacroDefault& m : kFixedMacros)
        if (!std::strcmp(name, m.name))
        {
            *out_value = m.value;
            return B2PT_OK;
        }
    b2_set_error("'%s' is not a macro served at one value only", name);
    return B2PT_ERR_NO_VARIANT;
}

/* Entry name + macro values -> the precompiled entry and the launch-time switches. Host logic only: no context, no CUDA call. */
struct B2ResolvedKernel
{
    const B2KernelEntry* entry = nullptr;
    int unit_precise = 1;
    int light_presampling = 0, alpha_test = 0, bias_correction = 4, restir_options = 4 | 8;
};

static int resolve_kernel(const char* entry_name, const b2pt_option* options, int option_count, B2ResolvedKernel& resolved)
{
    if (!entry_name || option_count < 0 || (option_count > 0 && !options))
    {
        b2_set_error("b2pt_kernel_get: invalid argument");
        return B2PT_ERR_INVALID_ARGUMENT;
    }
    int pass = pass_of_entry_name(entry_name);
    if (pass < 0)
    {
        b2_set_error("b2pt_kernel_get: unknown kernel entry '%s'", entry_name);
        return B2PT_ERR_NO_VARIANT;
    }
    /* reference defaults */
    int bsdf = 0, direct = 4, base = 1, later = 3, envmap = 2, precise = 1, counters = 0, light_presampling = 0, alpha_test = 0, bias_correction = 4, restir_options = 4 | 8;
    int presampling_strategy = 1; /* ReSTIR_DI_LightPresamplingStrategy: LSS_BASE_POWER */
    for (int i = 0; i < option_count; i++)
    {
        const char* n = options[i].name;
        int v = options[i].value;
        if (!n) { b2_set_error("b2pt_kernel_get: null option name"); return B2PT_ERR_INVALID_ARGUMENT; }
        if (!std::strcmp(n, "BSDFOverride")) bsdf = (v == 3) ? 0 : v; /* BSDF_PRINCIPLED == BSDF_NONE in the dispatcher */
        else if (!std::strcmp(n, "DirectLightSamplingStrategy")) direct = v;
        else if (!std::strcmp(n, "DirectLightSamplingBaseStrategy")) base = v;
        else if (!std::strcmp(n, "ReSTIR_DI_LaterBouncesSamplingStrategy")) later = v;
        else if (!std::strcmp(n, "EnvmapSamplingStrategy")) envmap = v;
        else if (!std::strcmp(n, "ReSTIR_DI_DoLightPresampling")) light_presampling = v != 0; /* launch-time switch of the same kernels */
        else if (!std::strcmp(n, "ReSTIR_DI_LightPresamplingStrategy")) presampling_strategy = v; /* 0 uniform, 1 power alias table: launch-time too */
        else if (!std::strcmp(n, "B2PT_PRECISE")) precise = v;
        else if (!std::strcmp(n, "B2PT_ALPHA_TEST")) alpha_test = v != 0; /* launch-time switch of the same kernels */
        else if (!std::strcmp(n, "ReSTIR_DI_BiasCorrectionWeights")) bias_correction = v; /* launch-time switch of the reuse passes */
        /* ReSTIRDIOptions.h:50-84, launch-time switches of the ReSTIR DI passes (bits B2_DI_OPT_*) */
        else if (!std::strcmp(n, "ReSTIR_DI_InitialTargetFunctionVisibility")) restir_options = v ? (restir_options | 1) : (restir_options & ~1);
        else if (!std::strcmp(n, "ReSTIR_DI_SpatialTargetFunctionVisibility")) restir_options = v ? (restir_options | 2) : (restir_options & ~2);
        else if (!std::strcmp(n, "ReSTIR_DI_DoVisibilityReuse")) restir_options = v ? (restir_options | 4) : (restir_options & ~4);
        else if (!std::strcmp(n, "ReSTIR_DI_BiasCorrectionUseVisibility")) restir_options = v ? (restir_options | 8) : (restir_options & ~8);
        else if (!std::strcmp(n, "ReSTIR_DI_DoOptimalVisibilitySampling")) restir_options = v ? (restir_options | 16) : (restir_options & ~16); /* spatial pass only */
        else if (!std::strcmp(n, "B2PT_COUNTERS")) counters = v;
        else
        {
            bool known = false;
            for (const MacroDefault& m : kFixedMacros)
                if (!std::strcmp(n, m.name))
                {
                    known = true;
                    if (m.value != v)
                    {
                        b2_set_error("no precompiled variant with %s=%d (only the reference default %d is built)", n, v, m.value);
                        return B2PT_ERR_NO_VARIANT;
                    }
                }
            if (!known)
            {
                b2_set_error("no precompiled variant knows the option '%s'", n);
                return B2PT_ERR_NO_VARIANT;
            }
        }
    }
    /* ReSTIR_DI_BiasCorrectionWeights (ReSTIRDIOptions.h:9-16): 1/Z, MIS-like, generalised balance heuristic, pairwise MIS (the default),
     * defensive pairwise MIS and the symmetric / asymmetric ratio heuristics are built for the temporal, spatial and fused spatiotemporal
     * reuse passes; 1/M does not compile in the reference itself (SpatialReuse.h:237 calls its normalisation with 6 of 7 arguments) */
    if (presampling_strategy != 0 && presampling_strategy != 1)
    {
        b2_set_error("no precompiled variant with ReSTIR_DI_LightPresamplingStrategy=%d (uniform = 0 and power = 1 are built)", presampling_strategy);
        return B2PT_ERR_NO_VARIANT;
    }
    if (light_presampling && presampling_strategy == 0)
        light_presampling = 2;
    {
        /* 1/M (0): the reference's spatial pass does not compile with it; its temporal and fused passes do, and are served */
        const bool built = (bias_correction >= 1 && bias_correction <= 7) || (bias_correction == 0 && pass != B2_PASS_RESTIR_DI_SPATIAL);
        const bool reuse_pass = pass == B2_PASS_RESTIR_DI_TEMPORAL || pass == B2_PASS_RESTIR_DI_SPATIAL || pass == B2_PASS_RESTIR_DI_SPATIOTEMPORAL;
        if (reuse_pass && !built)
        {
            b2_set_error("no precompiled variant of %s with ReSTIR_DI_BiasCorrectionWeights=%d", entry_name, bias_correction);
            return B2PT_ERR_NO_VARIANT;
        }
    }
    for (const B2KernelUnit& unit : kernel_units())
    {
        if ((unit.precise != 0) != (precise != 0))
            continue;
        for (int i = 0; i < unit.count; i++)
        {
            const B2KernelEntry& e = unit.entries[i];
            bool later_matters = e.direct == 5 && pass == B2_PASS_MEGAKERNEL; /* LSS_RESTIR_DI: ReSTIR_DI_LaterBouncesSamplingStrategy is read by the MegaKernel only */
            if (e.pass == pass && e.bsdf_override == bsdf && e.direct == direct && e.base == base && (!later_matters || e.later == later) && e.envmap == envmap && e.counters == (counters ? 1 : 0))
            {
                resolved.entry = &e;
                resolved.unit_precise = unit.precise;
                resolved.light_presampling = light_presampling;
                resolved.alpha_test = alpha_test;
                resolved.bias_correction = bias_correction;
                resolved.restir_options = restir_options;
                return B2PT_OK;
            }
        }
    }
    b2_set_error("no precompiled variant for %s with BSDFOverride=%d DirectLightSamplingStrategy=%d DirectLightSamplingBaseStrategy=%d "
                 "ReSTIR_DI_LaterBouncesSamplingStrategy=%d EnvmapSamplingStrategy=%d precise=%d counters=%d",
                 entry_name, bsdf, direct, base, later, envmap, precise, counters);
    return B2PT_ERR_NO_VARIANT;
}

int b2pt_kernel_query(const char* entry_name, const b2pt_option* options, int option_count)
{
    B2ResolvedKernel resolved;
    return resolve_kernel(entry_name, options, option_count, resolved);
}

int b2pt_kernel_get(b2pt_ctx* ctx, const char* entry_name, const b2pt_option* options, int option_count, b2pt_kernel** out_kernel)
{
    if (!ctx || !out_kernel)
    {
        b2_set_error("b2pt_kernel_get: invalid argument");
        return B2PT_ERR_INVALID_ARGUMENT;
    }
    B2ResolvedKernel resolved;
    int rc = resolve_kernel(entry_name, options, option_count, resolved);
    if (rc != B2PT_OK)
        return rc;
    B2_SELECT_DEVICE(ctx);
    b2pt_kernel* k = new b2pt_kernel();
    k->ctx = ctx;
    k->entry = resolved.entry;
    k->light_presampling = resolved.light_presampling;
    k->precise = resolved.unit_precise;
    k->alpha_test = resolved.alpha_test;
    k->bias_correction = resolved.bias_correction;
    k->restir_options = resolved.restir_options;
    B2_CUDA(cudaEventCreate(&k->start));
    B2_CUDA(cudaEventCreate(&k->stop));
    *out_kernel = k;
    return B2PT_OK;
}

/* Form of the MegaKernel (all four render bit-identical frames, tests/test_gpu_parity.py):
 *   3 = "wavefront": path records parked in HBM between stages and moved by the warp in full lines, all rays of a bounce in one
 *       queue drained at 32 warps / SM, shading kernels over paths sorted by material (b2_wavefront.h). The fastest form of the
 *       PRECISE flavour everywhere (10M triangles: 2.3x - 3.1x the phased form; Cornell 1080p: 7.5 ms against 12.2 ms per sample,
 *       profiles/r02_results.md) and of the fast flavour once the scene no longer fits the L2.
 *   2 = "phased": persistent 512-thread blocks, bounce cut at its rays, block-pooled traversal; the fastest form of the FAST flavour
 *       on L2-resident scenes (Cornell: 5.6 ms against 6.0 ms), where its 1.3 KB stack frames cost least.
 *   1 = "persistent" (persistent warps, rays traced in place), 0 = "tiles" (one thread per pixel): kept for A/B runs and tests.
 * Default: wavefront for the precise flavour and for any BVH of at least B2_WAVEFRONT_MIN_NODES nodes, phased otherwise.
 * B2PT_MEGAKERNEL=wavefront|phased|persistent|tiles forces one; it is read at every launch so a test can switch between them. */
#define B2_WAVEFRONT_MIN_NODES 32768u
static int megakernel_form(uint64_t bvh_nodes, bool precise)
{
    const char* e = std::getenv("B2PT_MEGAKERNEL");
    if (e && !std::strcmp(e, "tiles"))
        return 0;
    if (e && !std::strcmp(e, "persistent"))
        return 1;
    if (e && !std::strcmp(e, "phased"))
        return 2;
    if (e && !std::strcmp(e, "wavefront"))
        return 3;
    return (precise || bvh_nodes >= B2_WAVEFRONT_MIN_NODES) ? 3 : 2;
}

int b2pt_kernel_launch_rows(b2pt_kernel* kernel, int width, int height, int rank, int world, void** args, b2pt_stream stream)
{
    return b2pt_kernel_launch_partition(kernel, width, height, rank, world, B2PT_BAND_ROWS, nullptr, 0, args, stream);
}

int b2pt_kernel_launch_partition(b2pt_kernel* kernel, int width, int height, int rank, int world, int band_rows, const int* device_rows, int n_rows, void** args,
                                 b2pt_stream stream)
{
    if (!kernel || !args || !args[0] || width <= 0 || height <= 0 || world < 1 || rank < 0 || rank >= world || band_rows < 1 || n_rows < 0 || (device_rows && n_rows > height))
    {
        b2_set_error("b2pt_kernel_launch: invalid argument");
        return B2PT_ERR_INVALID_ARGUMENT;
    }
    /* ReSTIR_DI_LightsPresampling takes (LightPresamplingParameters, HIPRTRenderData) like the reference's kernel
     * (ReSTIRDIRenderPass.cpp:441-444); every other entry takes the render data alone */
    const bool presampling_pass = kernel->entry->pass == B2_PASS_RESTIR_DI_PRESAMPLING;
    if (presampling_pass && !args[1])
    {
        b2_set_error("ReSTIR_DI_LightsPresampling: args = { &LightPresamplingParameters, &render_data }");
        return B2PT_ERR_INVALID_ARGUMENT;
    }
    const b2pt_render_data* rd = reinterpret_cast<const b2pt_render_data*>(presampling_pass ? args[1] : args[0]);
    cudaStream_t s = (cudaStream_t)stream;
    B2_SELECT_DEVICE(kernel->ctx);
    B2LaunchGeom geom;
    geom.width = width;
    geom.height = height;
    geom.rank = rank;
    geom.world = world;
    geom.band_rows = band_rows;
    geom.row_map = device_rows;
    geom.row_count = device_rows ? n_rows : 0;
    geom.light_presampling = kernel->light_presampling;
    geom.alpha_test = kernel->alpha_test;
    geom.bias_correction = kernel->bias_correction;
    geom.restir_options = kernel->restir_options;
    geom.presampling_parameters = presampling_pass ? args[0] : nullptr;
    geom.directional_masks_ull = nullptr;
    geom.directional_radius = nullptr;
    if (kernel->entry->pass == B2_PASS_RESTIR_DIRECTIONAL_REUSE)
    {
        /* args = { &render_data, &masks_u, &masks_ull, &radius } (ReSTIRDIRenderPass.cpp:409-411) */
        if (!args[1] || !args[2] || !args[3])
        {
            b2_set_error("ReSTIR_Directional_Reuse_Compute: args = { &render_data, &masks_u, &masks_ull, &radius }");
            return B2PT_ERR_INVALID_ARGUMENT;
        }
        geom.directional_masks_ull = *reinterpret_cast<void**>(args[2]);
        geom.directional_radius = *reinterpret_cast<void**>(args[3]);
    }
    geom.pool_rays = (rd->world_settings.ambient_light_type == 2 && rd->world_settings.envmap_intensity > 0.0f) ? 5 : 3;
    geom.stack_entries = B2_STACK_ENTRIES;
    uint64_t bvh_nodes = 0;
    {
        /* a traversal pushes at most one entry per level: depth + 1 entries are enough for this tree */
        std::lock_guard<std::mutex> lock(g_bvh_depth_mutex);
        auto it = g_bvh_depth.find(rd->GPU_BVH);
        if (it != g_bvh_depth.end())
        {
            if (it->second + 1 > B2_STACK_ENTRIES)
            {
                /* a deeper tree would overflow the per-thread traversal stack and silently lose subtrees: refuse the launch */
                b2_set_error("the BVH is %d levels deep; the traversal stack holds %d entries (B2_STACK_ENTRIES)", it->second, B2_STACK_ENTRIES);
                return B2PT_ERR_UNSUPPORTED_SCENE;
            }
            geom.stack_entries = std::max(4, std::min(B2_STACK_ENTRIES, it->second + 1));
            bvh_nodes = g_bvh_nodes[rd->GPU_BVH];
        }
    }
    geom.persistent = megakernel_form(bvh_nodes, kernel->precise != 0);
    if (kernel->alpha_test && (geom.persistent == 1 || geom.persistent == 2))
        geom.persistent = 3; /* the block-pooled forms queue rays without the alpha-test bit: alpha-tested scenes take the wavefront form */
    B2_CUDA(cudaEventRecord(kernel->start, s));
    B2_CUDA(kernel->entry->launch(rd, &geom, s));
    B2_CUDA(cudaEventRecord(kernel->stop, s));
    kernel->timed = true;
    kernel->launches++;
    return B2PT_OK;
}

int b2pt_kernel_launch(b2pt_kernel* kernel, int, int, int width, int height, void** args, b2pt_stream stream)
{
    return b2pt_kernel_launch_rows(kernel, width, height, 0, 1, args, stream);
}

int b2pt_kernel_destroy(b2pt_kernel* kernel)
{
    if (!kernel) return B2PT_OK;
    if (kernel->ctx) cudaSetDevice(kernel->ctx->device);
    if (kernel->start) cudaEventDestroy(kernel->start);
    if (kernel->stop) cudaEventDestroy(kernel->stop);
    delete kernel;
    return B2PT_OK;
}

int b2pt_kernel_last_ms(b2pt_kernel* kernel, float* out_ms)
{
    if (!kernel || !out_ms) { b2_set_error("null argument"); return B2PT_ERR_INVALID_ARGUMENT; }
    *out_ms = 0.0f;
    if (!kernel->timed)
        return B2PT_OK;
    B2_CUDA(cudaEventSynchronize(kernel->stop));
    B2_CUDA(cudaEventElapsedTime(out_ms, kernel->start, kernel->stop));
    return B2PT_OK;
}

int b2pt_kernel_launch_count(b2pt_kernel* kernel, uint64_t* out_count)
{
    if (!kernel || !out_count) { b2_set_error("null argument"); return B2PT_ERR_INVALID_ARGUMENT; }
    *out_count = kernel->launches;
    return B2PT_OK;
}

int b2pt_counters_reset(b2pt_ctx* ctx)
{
    B2_SELECT_DEVICE(ctx);
    for (const B2KernelUnit& unit : kernel_units())
        B2_CUDA(unit.counters_reset());
    return B2PT_OK;
}

int b2pt_counters_read(b2pt_ctx* ctx, b2pt_counters_t* out)
{
    B2_SELECT_DEVICE(ctx);
    if (!out) { b2_set_error("null argument"); return B2PT_ERR_INVALID_ARGUMENT; }
    B2_CUDA(cudaDeviceSynchronize());
    out->rays = out->nodes = out->triangles = 0;
    for (const B2KernelUnit& unit : kernel_units())
    {
        unsigned long long a[3];
        B2_CUDA(unit.counters_read(a));
        out->rays += a[0];
        out->nodes += a[1];
        out->triangles += a[2];
    }
    return B2PT_OK;
}

int b2pt_wavefront_profile(b2pt_ctx* ctx, int enabled)
{
    B2_SELECT_DEVICE(ctx);
    WfProfiler& p = wf_profiler();
    B2_CUDA(cudaDeviceSynchronize());
    std::lock_guard<std::mutex> lock(p.mutex);
    p.enabled = enabled != 0;
    p.kinds.clear();
    for (int k = 0; k < B2_WF_K_COUNT; k++)
    {
        p.ms[k] = 0.0;
        p.marks[k] = 0;
    }
    return B2PT_OK;
}

int b2pt_wavefront_profile_read(b2pt_ctx* ctx, double* out_ms, uint64_t* out_launch_groups, int kinds)
{
    B2_SELECT_DEVICE(ctx);
    if (!out_ms || !out_launch_groups || kinds < 0) { b2_set_error("null argument"); return B2PT_ERR_INVALID_ARGUMENT; }
    WfProfiler& p = wf_profiler();
    B2_CUDA(cudaDeviceSynchronize());
    std::lock_guard<std::mutex> lock(p.mutex);
    wf_profiler_fold(p);
    for (int k = 0; k < kinds; k++)
    {
        out_ms[k] = k < B2_WF_K_COUNT ? p.ms[k] : 0.0;
        out_launch_groups[k] = k < B2_WF_K_COUNT ? p.marks[k] : 0;
    }
    return B2PT_OK;
}

/* CUDA kernels launched by this library's path-tracing passes so far (all contexts of the process) */
int b2pt_cuda_launch_count(b2pt_ctx*, uint64_t* out_count)
{
    if (!out_count) { b2_set_error("null argument"); return B2PT_ERR_INVALID_ARGUMENT; }
    uint64_t n = 0;
    for (const B2KernelUnit& unit : kernel_units())
        n += unit.launches_read();
    *out_count = n;
    return B2PT_OK;
}

int b2pt_faults_read(b2pt_ctx* ctx, b2pt_faults_t* out, int reset)
{
    B2_SELECT_DEVICE(ctx);
    if (!out) { b2_set_error("null argument"); return B2PT_ERR_INVALID_ARGUMENT; }
    B2_CUDA(cudaDeviceSynchronize());
    out->traversal_guard_trips = out->boundary_guard_trips = out->stack_overflows = 0;
    for (const B2KernelUnit& unit : kernel_units())
    {
        unsigned int a[B2_FAULT_COUNT];
        B2_CUDA(unit.faults_read(a, reset));
        out->traversal_guard_trips += a[B2_FAULT_TRAVERSAL_GUARD];
        out->boundary_guard_trips += a[B2_FAULT_BOUNDARY_GUARD];
        out->stack_overflows += a[B2_FAULT_STACK_OVERFLOW];
    }
    return B2PT_OK;
}

} /* extern "C" */
