/*
 * b2_kernels.cu -- the sm_100a kernels of the path-tracing hot path.
 *
 * One __global__ per (variant, pass); each takes the 2784-byte render data BY VALUE exactly like the reference's
 * kernels (src/Device/kernels/CameraRays.h:82, Megakernel.h:23, kernels/ReSTIR/DI/*.h) -- as a __grid_constant__
 * so the per-pixel code reads it straight from the constant bank without a local copy.
 *
 * Thread -> pixel map: blocks of 64 threads (2 warps) own 8x8 pixel tiles; a warp covers an 8x4 sub-tile so the
 * 32 lanes start with coherent primary rays and coalesce their G-buffer / reservoir / framebuffer words.
 * Traversal stack: B2_STACK_ENTRIES 8-byte entries per thread in SHARED memory, entry-major ([entry][thread]) so a
 * warp's push/pop touches 32 consecutive 8-byte words (bank-conflict free). BVH nodes and triangles are fetched
 * with 128-bit read-only loads (b2_bvh.h).
 * This file is compiled twice (see build.py): B2_FLAVOUR_PRECISE with -fmad=false (parity flavour: IEEE fp32, no
 * contraction -- what the reference's host build computes) and B2_FLAVOUR_FAST with contraction + fast intrinsics.
 */
#include "b2_kernel_table.h"
#include "b2_kernels_body.h"

#if defined(B2_FLAVOUR_PRECISE)
#define B2_NS b2_precise
#elif defined(B2_FLAVOUR_FAST)
#define B2_NS b2_fast
#else
#error "define B2_FLAVOUR_PRECISE or B2_FLAVOUR_FAST"
#endif

#define B2_TILE_W 8
#define B2_TILE_H 8
#define B2_BLOCK_THREADS (B2_TILE_W * B2_TILE_H)

namespace B2_NS
{
__device__ unsigned long long g_counters[3];
__device__ unsigned int g_faults[B2_FAULT_COUNT];

template <class V, int PASS>
__global__ void __launch_bounds__(B2_BLOCK_THREADS) pass_kernel(const __grid_constant__ b2pt_render_data rd, int tiles_x)
{
    __shared__ uint64_t stacks[B2_STACK_ENTRIES * B2_BLOCK_THREADS];

    const int tile = blockIdx.x;
    const int tile_x = tile % tiles_x, tile_y = tile / tiles_x;
    /* lane l of warp w: x = l & 7, y = 4 w + (l >> 3) */
    const uint32_t x = tile_x * B2_TILE_W + (threadIdx.x & 7);
    const uint32_t y = tile_y * B2_TILE_H + (threadIdx.x >> 3);

    B2TraversalCounters local_counters;
    local_counters.rays = local_counters.nodes = local_counters.tris = 0;

    Ctx ctx;
    ctx.rd = &rd;
    if (rd.GPU_BVH != nullptr)
        ctx.bvh = *reinterpret_cast<const B2BvhDevice*>(rd.GPU_BVH);
    else
    {
        ctx.bvh.nodes = nullptr;
        ctx.bvh.tris = nullptr;
        ctx.bvh.node_count = ctx.bvh.tri_count = 0;
    }
    ctx.stack = stacks + threadIdx.x;
    ctx.stack_stride = B2_BLOCK_THREADS;
    ctx.counters = V::COUNTERS ? &local_counters : nullptr;
    ctx.faults = g_faults;

    b2_run_pass_pixel<V>(PASS, ctx, x, y);

    if (V::COUNTERS)
    {
        unsigned long long r = local_counters.rays, n = local_counters.nodes, t = local_counters.tris;
        for (int off = 16; off > 0; off >>= 1)
        {
            r += __shfl_down_sync(0xFFFFFFFFu, r, off);
            n += __shfl_down_sync(0xFFFFFFFFu, n, off);
            t += __shfl_down_sync(0xFFFFFFFFu, t, off);
        }
        if ((threadIdx.x & 31) == 0)
        {
            atomicAdd(&g_counters[0], r);
            atomicAdd(&g_counters[1], n);
            atomicAdd(&g_counters[2], t);
        }
    }
}

template <class V, int PASS>
cudaError_t launch_pass(const b2pt_render_data* rd, int width, int height, cudaStream_t stream)
{
    const int tiles_x = (width + B2_TILE_W - 1) / B2_TILE_W;
    const int tiles_y = (height + B2_TILE_H - 1) / B2_TILE_H;
    pass_kernel<V, PASS><<<tiles_x * tiles_y, B2_BLOCK_THREADS, 0, stream>>>(*rd, tiles_x);
    return cudaGetLastError();
}

template <int BSDF, int DIRECT, int BASE, int LATER, int ENVMAP, bool COUNTERS>
struct Reg
{
    typedef B2Variant<BSDF, DIRECT, BASE, LATER, ENVMAP, COUNTERS> V;
    template <int PASS>
    static B2KernelEntry entry()
    {
        B2KernelEntry e;
        e.pass = PASS;
        e.bsdf_override = BSDF;
        e.direct = DIRECT;
        e.base = BASE;
        e.later = LATER;
        e.envmap = ENVMAP;
        e.counters = COUNTERS ? 1 : 0;
        e.launch = &launch_pass<V, PASS>;
        return e;
    }
};

#define B2_REG_BASIC(BSDF, DIRECT, BASE, COUNTERS)                                                        \
    Reg<BSDF, DIRECT, BASE, B2_LATER_RIS, B2_ESS_ALIAS_TABLE, COUNTERS>::entry<B2_PASS_CAMERA_RAYS>(),    \
    Reg<BSDF, DIRECT, BASE, B2_LATER_RIS, B2_ESS_ALIAS_TABLE, COUNTERS>::entry<B2_PASS_MEGAKERNEL>()

#define B2_REG_RESTIR(BSDF, BASE, COUNTERS)                                                                                       \
    Reg<BSDF, B2_LSS_RESTIR_DI, BASE, B2_LATER_RIS, B2_ESS_ALIAS_TABLE, COUNTERS>::entry<B2_PASS_RESTIR_DI_INITIAL_CANDIDATES>(), \
    Reg<BSDF, B2_LSS_RESTIR_DI, BASE, B2_LATER_RIS, B2_ESS_ALIAS_TABLE, COUNTERS>::entry<B2_PASS_RESTIR_DI_TEMPORAL>(),           \
    Reg<BSDF, B2_LSS_RESTIR_DI, BASE, B2_LATER_RIS, B2_ESS_ALIAS_TABLE, COUNTERS>::entry<B2_PASS_RESTIR_DI_SPATIAL>(),            \
    Reg<BSDF, B2_LSS_RESTIR_DI, BASE, B2_LATER_RIS, B2_ESS_ALIAS_TABLE, COUNTERS>::entry<B2_PASS_RESTIR_DI_SPATIOTEMPORAL>()

static const B2KernelEntry* table(int* count)
{
    static const B2KernelEntry entries[] = {
        B2_REG_BASIC(B2_BSDF_LAMBERTIAN, B2_LSS_ONE_LIGHT, B2_BASE_UNIFORM, false),
        B2_REG_BASIC(B2_BSDF_LAMBERTIAN, B2_LSS_ONE_LIGHT, B2_BASE_UNIFORM, true),
#ifdef B2_HAVE_PRINCIPLED
        B2_REG_BASIC(B2_BSDF_NONE, B2_LSS_MIS, B2_BASE_POWER, false),
        B2_REG_BASIC(B2_BSDF_NONE, B2_LSS_MIS, B2_BASE_POWER, true),
#endif
#ifdef B2_HAVE_RIS
        B2_REG_BASIC(B2_BSDF_NONE, B2_LSS_RIS, B2_BASE_POWER, false),
        B2_REG_BASIC(B2_BSDF_NONE, B2_LSS_RIS, B2_BASE_POWER, true),
#endif
#ifdef B2_HAVE_RESTIR_DI
        B2_REG_BASIC(B2_BSDF_NONE, B2_LSS_RESTIR_DI, B2_BASE_POWER, false),
        B2_REG_RESTIR(B2_BSDF_NONE, B2_BASE_POWER, false),
        B2_REG_BASIC(B2_BSDF_NONE, B2_LSS_RESTIR_DI, B2_BASE_POWER, true),
        B2_REG_RESTIR(B2_BSDF_NONE, B2_BASE_POWER, true),
#endif
    };
    *count = (int)(sizeof(entries) / sizeof(entries[0]));
    return entries;
}
} // namespace B2_NS

#if defined(B2_FLAVOUR_PRECISE)
const B2KernelEntry* b2_kernel_table_precise(int* count) { return b2_precise::table(count); }
cudaError_t b2_counters_reset_precise()
{
    unsigned long long zero[3] = { 0, 0, 0 };
    return cudaMemcpyToSymbol(b2_precise::g_counters, zero, sizeof(zero));
}
cudaError_t b2_counters_read_precise(unsigned long long out[3]) { return cudaMemcpyFromSymbol(out, b2_precise::g_counters, 3 * sizeof(unsigned long long)); }
cudaError_t b2_faults_read_precise(unsigned int out[B2_FAULT_COUNT], int reset)
{
    cudaError_t e = cudaMemcpyFromSymbol(out, b2_precise::g_faults, B2_FAULT_COUNT * sizeof(unsigned int));
    if (e != cudaSuccess || !reset)
        return e;
    unsigned int zero[B2_FAULT_COUNT] = {};
    return cudaMemcpyToSymbol(b2_precise::g_faults, zero, sizeof(zero));
}
#else
const B2KernelEntry* b2_kernel_table_fast(int* count) { return b2_fast::table(count); }
cudaError_t b2_counters_reset_fast()
{
    unsigned long long zero[3] = { 0, 0, 0 };
    return cudaMemcpyToSymbol(b2_fast::g_counters, zero, sizeof(zero));
}
cudaError_t b2_counters_read_fast(unsigned long long out[3]) { return cudaMemcpyFromSymbol(out, b2_fast::g_counters, 3 * sizeof(unsigned long long)); }
cudaError_t b2_faults_read_fast(unsigned int out[B2_FAULT_COUNT], int reset)
{
    cudaError_t e = cudaMemcpyFromSymbol(out, b2_fast::g_faults, B2_FAULT_COUNT * sizeof(unsigned int));
    if (e != cudaSuccess || !reset)
        return e;
    unsigned int zero[B2_FAULT_COUNT] = {};
    return cudaMemcpyToSymbol(b2_fast::g_faults, zero, sizeof(zero));
}
#endif
