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

#include <atomic>
#include <mutex>

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
#define B2_BAND_ROWS B2PT_BAND_ROWS
#define B2_WORK_COUNTER_SLOTS 256
#define B2_MAX_DEVICES 64
/* phased MegaKernel (default): threads per block and resident blocks per SM it is compiled for (register cap 65536 / (T * N)) */
#ifndef B2_MK_THREADS
#define B2_MK_THREADS 256
#endif
#ifndef B2_MK_MIN_BLOCKS
#define B2_MK_MIN_BLOCKS 3
#endif
/* persistent MegaKernel without phases (kept for A/B measurements): 64-thread blocks, 12 resident per SM */
#define B2_MKP_THREADS 64
#define B2_MKP_MIN_BLOCKS 12
#define B2_MKP_BLOCK_SYNC 0

namespace B2_NS
{
__device__ unsigned long long g_counters[3];
__device__ unsigned int g_faults[B2_FAULT_COUNT];

/* Pixel rows are partitioned across GPUs in interleaved bands of B2_BAND_ROWS rows (band b belongs to rank b % world);
 * a launch covers the bands of one rank. `local_row` counts the rows of this rank's bands, top to bottom. */
__device__ __forceinline__ uint32_t b2_global_row(uint32_t local_row, const B2LaunchGeom& g)
{
    const uint32_t band_local = local_row / B2_BAND_ROWS;
    return (band_local * (uint32_t)g.world + (uint32_t)g.rank) * B2_BAND_ROWS + local_row % B2_BAND_ROWS;
}

__device__ __forceinline__ void b2_make_ctx(Ctx& ctx, const b2pt_render_data& rd, uint64_t* stacks, B2TraversalCounters* counters, unsigned int* faults, int threads = B2_BLOCK_THREADS)
{
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
    ctx.stack_stride = threads;
    ctx.counters = counters;
    ctx.faults = faults;
}

__device__ __forceinline__ void b2_flush_counters(const B2TraversalCounters& local_counters)
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

/* One thread per pixel, 8x8 tiles (CameraRays and the ReSTIR DI passes: a fixed, short amount of work per pixel). */
template <class V, int PASS>
__global__ void __launch_bounds__(B2_BLOCK_THREADS) pass_kernel(const __grid_constant__ b2pt_render_data rd, const B2LaunchGeom geom, int tiles_x)
{
    __shared__ uint64_t stacks[B2_STACK_ENTRIES * B2_BLOCK_THREADS];

    const int tile = blockIdx.x;
    const int tile_x = tile % tiles_x, tile_y = tile / tiles_x;
    /* lane l of warp w: x = l & 7, y = 4 w + (l >> 3) */
    const uint32_t x = tile_x * B2_TILE_W + (threadIdx.x & 7);
    const uint32_t y = b2_global_row(tile_y * B2_TILE_H + (threadIdx.x >> 3), geom);

    B2TraversalCounters local_counters;
    local_counters.rays = local_counters.nodes = local_counters.tris = 0;

    Ctx ctx;
    b2_make_ctx(ctx, rd, stacks, V::COUNTERS ? &local_counters : nullptr, g_faults);

    b2_run_pass_pixel<V>(PASS, ctx, x, y);

    if (V::COUNTERS)
        b2_flush_counters(local_counters);
}

/*
 * The MegaKernel as PERSISTENT warps with path regeneration. The grid is (SM count x resident blocks per SM). A warp
 * fetches work one 8x4-pixel tile at a time from a global counter (one atomicAdd per tile): the tile's primary-miss
 * pixels are finished on the spot by all of their lanes together (megakernel_primary_miss), its hit pixels go into a
 * small per-warp queue in shared memory. A lane whose path ended -- Russian roulette, absorption, bounce limit, escape
 * -- takes the next queued pixel at the top of the loop while the other lanes carry on with their bounce, so
 * `megakernel_bounce` runs with (nearly) all 32 lanes holding a live surface hit until the work runs out. In the
 * one-thread-per-pixel form the warp idled lanes until its longest path ended (6.2 of 32 lanes active on the C2
 * Cornell frame, profiles/r01_ncu_c2_megakernel.md); handing out single pixels instead of tiles still spent a third of
 * the bounce loop's lanes on primary misses (profiles/r01b_ncu_c2_megakernel.md).
 * Per-pixel results do not depend on the schedule: every pixel owns its RNG stream and its framebuffer word.
 */
#define B2_QUEUE_SLOTS 64
template <class V>
__global__ void __launch_bounds__(B2_MKP_THREADS, B2_MKP_MIN_BLOCKS) megakernel_persistent(const __grid_constant__ b2pt_render_data rd, const B2LaunchGeom geom, int tiles_x,
                                                                                             unsigned int total_tiles, unsigned int* work_counter)
{
    __shared__ uint64_t stacks[B2_STACK_ENTRIES * B2_MKP_THREADS];
    __shared__ uint32_t queues[B2_MKP_THREADS / 32][B2_QUEUE_SLOTS];

    B2TraversalCounters local_counters;
    local_counters.rays = local_counters.nodes = local_counters.tris = 0;

    Ctx ctx;
    b2_make_ctx(ctx, rd, stacks, V::COUNTERS ? &local_counters : nullptr, g_faults, B2_MKP_THREADS);

    const unsigned int lane = threadIdx.x & 31u;
    const unsigned int lanes_below = (1u << lane) - 1u;
    uint32_t* queue = queues[threadIdx.x >> 5];
    const int res_x = rd.render_settings.render_resolution.x, res_y = rd.render_settings.render_resolution.y;
    PathState s;
    bool alive = false;
    /* warp-uniform */
    bool exhausted = false;
    unsigned int q_head = 0, q_count = 0;

    while (true)
    {
        const unsigned int idle = __ballot_sync(0xFFFFFFFFu, !alive);
        const unsigned int n_idle = __popc(idle);
        while (q_count < n_idle && !exhausted)
        {
            unsigned int tile = 0;
            if (lane == 0)
                tile = atomicAdd(work_counter, 1u);
            tile = __shfl_sync(0xFFFFFFFFu, tile, 0);
            if (tile >= total_tiles)
            {
                exhausted = true;
                break;
            }
            const uint32_t x = (tile % (unsigned int)tiles_x) * 8u + (lane & 7u);
            const uint32_t y = b2_global_row((tile / (unsigned int)tiles_x) * 4u + (lane >> 3), geom);
            bool is_hit = false;
            if (x < (uint32_t)res_x && y < (uint32_t)res_y)
            {
                const uint32_t pixel_index = x + y * (uint32_t)res_x;
                if (rd.aux_buffers.pixel_active[pixel_index])
                {
                    is_hit = rd.g_buffer.first_hit_prim_index[pixel_index] != -1;
                    if (!is_hit)
                        megakernel_primary_miss<V>(ctx, x, y, pixel_index);
                }
            }
            const unsigned int hits = __ballot_sync(0xFFFFFFFFu, is_hit);
            if (is_hit)
                queue[(q_head + q_count + __popc(hits & lanes_below)) % B2_QUEUE_SLOTS] = x | (y << 16);
            q_count += __popc(hits);
            __syncwarp();
        }
        if (!alive)
        {
            const unsigned int r = __popc(idle & lanes_below);
            if (r < q_count)
            {
                const uint32_t packed = queue[(q_head + r) % B2_QUEUE_SLOTS];
                alive = megakernel_begin<V>(ctx, packed & 0xFFFFu, packed >> 16, s);
            }
        }
        {
            const unsigned int taken = n_idle < q_count ? n_idle : q_count;
            q_head = (q_head + taken) % B2_QUEUE_SLOTS;
            q_count -= taken;
        }
        __syncwarp();
#if B2_MKP_BLOCK_SYNC
        if (!__syncthreads_or(alive ? 1 : 0))
        {
            /* no live path in the whole block: either every warp ran out of work, or (rarely) the freshly fetched tiles
             * held no hit pixel -- in that case go round again (block-uniform decision) */
            if (__syncthreads_and((exhausted && q_count == 0) ? 1 : 0))
                break;
            continue;
        }
#else
        if (!__any_sync(0xFFFFFFFFu, alive))
        {
            if (exhausted && q_count == 0)
                break;
            continue;
        }
#endif
        if (alive)
        {
            if (!megakernel_bounce<V>(ctx, s))
            {
                megakernel_end(ctx, s);
                alive = false;
            }
        }
    }

    if (V::COUNTERS)
        b2_flush_counters(local_counters);
}

/*
 * The MegaKernel in PHASES (default). Same persistent warps, tile fetch and per-warp hit queue as megakernel_persistent
 * above, plus two things its ncu profile asked for (profiles/r01c_ncu_c2_megakernel.md: 6 of 32 lanes active inside the
 * traversal, `no_instruction` = 70 % of all stall samples because every warp of an SM sat in a different corner of a
 * 180 KB kernel while the L1.5 instruction cache holds 32 KB):
 *   1. the bounce is cut at its rays (b2_pathtrace.h: stage A / B / C, b2_nee.h) and the rays of the WHOLE BLOCK are traced
 *      together: every thread drops its ray into a shared-memory pool, then all threads drain the pool in one converged
 *      loop of traversal steps -- a lane whose ray is done takes the next ray from the pool while the other lanes keep
 *      stepping (warp-aggregated atomicAdd on the pool cursor), so the loop runs with full warps until the pool is empty
 *      no matter how many lanes asked for a ray or how long each ray is;
 *   2. block barriers around every trace keep the block's warps in the same stage of the code, so a stage's instructions
 *      are fetched once and reused by all of its warps.
 * A path's arithmetic, RNG draws and results are the same as in the other two forms; tests/ run all three against the same
 * fixtures.
 */
struct B2RayPool
{
    float4* ray_a;   /* origin.xyz, t_max                         [threads] */
    float4* ray_b;   /* direction.xyz, skip_prim | any-hit flag   [threads] */
    float4* hit_a;   /* t, u, v, prim                             [threads] */
    float4* hit_b;   /* ng.xyz                                    [threads] */
    int* list;       /* compacted slots of this phase's rays      [threads] */
    int* count;      /* [2]: rays in the pool, by phase parity */
    int* next;       /* [2]: drain cursor, by phase parity */
};

/* All threads of the block call this the same number of times. `ray.kind == B2_RAY_NONE`: the thread has no ray this phase
 * but still helps draining. Returns this thread's result (prim = -1 when it had no ray). */
__device__ __forceinline__ B2Hit b2_trace_pooled(const Ctx& ctx, const B2RayPool& pool, unsigned int& phase, const RayRequest& ray, B2TraversalCounters* counters)
{
    const unsigned int lane = threadIdx.x & 31u;
    const unsigned int lanes_below = (1u << lane) - 1u;
    const int pc = (int)(phase & 1u);
    phase++;

    /* deposit (warp-aggregated slot allocation) */
    const bool have = ray.kind != B2_RAY_NONE && ctx.bvh.node_count != 0;
    {
        const unsigned int havers = __ballot_sync(0xFFFFFFFFu, have);
        int base = 0;
        if (lane == 0 && havers != 0u)
            base = atomicAdd(&pool.count[pc], __popc(havers));
        base = __shfl_sync(0xFFFFFFFFu, base, 0);
        if (have)
        {
            pool.list[base + __popc(havers & lanes_below)] = (int)threadIdx.x;
            pool.ray_a[threadIdx.x] = make_float4(ray.origin.x, ray.origin.y, ray.origin.z, ray.t_max);
            pool.ray_b[threadIdx.x] = make_float4(ray.direction.x, ray.direction.y, ray.direction.z,
                                                  __int_as_float(ray.kind == B2_RAY_ANY_HIT ? 1 : 0));
            pool.hit_b[threadIdx.x] = make_float4(__int_as_float(ray.skip_prim), 0.0f, 0.0f, 0.0f); /* skip_prim travels in hit_b.x until the hit overwrites it */
        }
    }
    if (threadIdx.x == 0)
    {
        /* the other parity's counters are idle during this phase: reset them for the next one */
        pool.count[pc ^ 1] = 0;
        pool.next[pc ^ 1] = 0;
    }
    __syncthreads();

    /* drain: one converged loop of traversal steps over all rays of the block */
    const int count = pool.count[pc];
    B2Traversal t;
    bool busy = false;
    bool drained = false; /* warp-uniform */
    int slot = 0;
    while (true)
    {
        const unsigned int idle = __ballot_sync(0xFFFFFFFFu, !busy);
        if (idle != 0u && !drained)
        {
            const int wanted = __popc(idle);
            int base = 0;
            if (lane == 0)
                base = atomicAdd(&pool.next[pc], wanted);
            base = __shfl_sync(0xFFFFFFFFu, base, 0);
            if (!busy)
            {
                const int i = base + __popc(idle & lanes_below);
                if (i < count)
                {
                    slot = pool.list[i];
                    const float4 a = pool.ray_a[slot], b = pool.ray_b[slot];
                    const int skip_prim = __float_as_int(pool.hit_b[slot].x);
                    b2_traverse_begin(t, __float_as_int(b.w) != 0, v3(a.x, a.y, a.z), v3(b.x, b.y, b.z), a.w, skip_prim, ctx.stack);
                    busy = true;
                }
            }
            drained = base + wanted >= count;
        }
        if (!__any_sync(0xFFFFFFFFu, busy))
        {
            if (drained)
                break;
            continue;
        }
        if (busy)
        {
            if (!b2_traverse_step(t, ctx.bvh, ctx.stack_stride, ctx.faults))
            {
                pool.hit_a[slot] = make_float4(t.best.t, t.best.u, t.best.v, __int_as_float(t.best.prim));
                pool.hit_b[slot] = make_float4(t.best.ng.x, t.best.ng.y, t.best.ng.z, 0.0f);
                if (counters)
                {
                    counters->rays += 1;
                    counters->nodes += t.n_nodes;
                    counters->tris += t.n_tris;
                }
                busy = false;
            }
        }
    }
    __syncthreads();

    B2Hit result;
    result.t = -1.0f;
    result.u = result.v = 0.0f;
    result.prim = -1;
    result.ng = v3(0.0f);
    if (have)
    {
        const float4 a = pool.hit_a[threadIdx.x], b = pool.hit_b[threadIdx.x];
        result.t = a.x;
        result.u = a.y;
        result.v = a.z;
        result.prim = __float_as_int(a.w);
        result.ng = v3(b.x, b.y, b.z);
    }
    return result;
}

#define B2_PH_SMEM_BYTES (B2_STACK_ENTRIES * B2_MK_THREADS * 8 + (B2_MK_THREADS / 32) * B2_QUEUE_SLOTS * 4 + B2_MK_THREADS * (4 * 16 + 4) + 16)

template <class V>
__global__ void __launch_bounds__(B2_MK_THREADS, B2_MK_MIN_BLOCKS) megakernel_phased(const __grid_constant__ b2pt_render_data rd, const B2LaunchGeom geom, int tiles_x,
                                                                                         unsigned int total_tiles, unsigned int* work_counter)
{
    extern __shared__ __align__(16) unsigned char smem[];
    uint64_t* stacks = reinterpret_cast<uint64_t*>(smem);
    B2RayPool pool;
    pool.ray_a = reinterpret_cast<float4*>(smem + B2_STACK_ENTRIES * B2_MK_THREADS * 8);
    pool.ray_b = pool.ray_a + B2_MK_THREADS;
    pool.hit_a = pool.ray_b + B2_MK_THREADS;
    pool.hit_b = pool.hit_a + B2_MK_THREADS;
    pool.list = reinterpret_cast<int*>(pool.hit_b + B2_MK_THREADS);
    uint32_t* queues = reinterpret_cast<uint32_t*>(pool.list + B2_MK_THREADS);
    pool.count = reinterpret_cast<int*>(queues + (B2_MK_THREADS / 32) * B2_QUEUE_SLOTS);
    pool.next = pool.count + 2;
    if (threadIdx.x < 4)
        pool.count[threadIdx.x] = 0; /* count[0..1], next[0..1] */
    __syncthreads();

    B2TraversalCounters local_counters;
    local_counters.rays = local_counters.nodes = local_counters.tris = 0;
    B2TraversalCounters* counters = V::COUNTERS ? &local_counters : nullptr;

    Ctx ctx;
    b2_make_ctx(ctx, rd, stacks, counters, g_faults, B2_MK_THREADS);

    const unsigned int lane = threadIdx.x & 31u;
    const unsigned int lanes_below = (1u << lane) - 1u;
    uint32_t* queue = queues + (threadIdx.x >> 5) * B2_QUEUE_SLOTS;
    const int res_x = rd.render_settings.render_resolution.x, res_y = rd.render_settings.render_resolution.y;
    PathState s;
    BouncePhase<V> ph;
    RayRequest ray;
    request_none(ray);
    unsigned int phase = 0;
    bool alive = false;
    /* warp-uniform */
    bool exhausted = false;
    unsigned int q_head = 0, q_count = 0;

    while (true)
    {
        /* refill: idle lanes take queued hit pixels; the queue is topped up one 8x4 tile at a time */
        const unsigned int idle = __ballot_sync(0xFFFFFFFFu, !alive);
        const unsigned int n_idle = __popc(idle);
        while (q_count < n_idle && !exhausted)
        {
            unsigned int tile = 0;
            if (lane == 0)
                tile = atomicAdd(work_counter, 1u);
            tile = __shfl_sync(0xFFFFFFFFu, tile, 0);
            if (tile >= total_tiles)
            {
                exhausted = true;
                break;
            }
            const uint32_t x = (tile % (unsigned int)tiles_x) * 8u + (lane & 7u);
            const uint32_t y = b2_global_row((tile / (unsigned int)tiles_x) * 4u + (lane >> 3), geom);
            bool is_hit = false;
            if (x < (uint32_t)res_x && y < (uint32_t)res_y)
            {
                const uint32_t pixel_index = x + y * (uint32_t)res_x;
                if (rd.aux_buffers.pixel_active[pixel_index])
                {
                    is_hit = rd.g_buffer.first_hit_prim_index[pixel_index] != -1;
                    if (!is_hit)
                        megakernel_primary_miss<V>(ctx, x, y, pixel_index);
                }
            }
            const unsigned int hits = __ballot_sync(0xFFFFFFFFu, is_hit);
            if (is_hit)
                queue[(q_head + q_count + __popc(hits & lanes_below)) % B2_QUEUE_SLOTS] = x | (y << 16);
            q_count += __popc(hits);
            __syncwarp();
        }
        if (!alive)
        {
            const unsigned int r = __popc(idle & lanes_below);
            if (r < q_count)
            {
                const uint32_t packed = queue[(q_head + r) % B2_QUEUE_SLOTS];
                alive = megakernel_begin<V>(ctx, packed & 0xFFFFu, packed >> 16, s);
                request_none(ray); /* bounce 0 starts from the G-buffer hit */
            }
        }
        {
            const unsigned int taken = n_idle < q_count ? n_idle : q_count;
            q_head = (q_head + taken) % B2_QUEUE_SLOTS;
            q_count -= taken;
        }
        __syncwarp();
        if (!__syncthreads_or(alive ? 1 : 0))
        {
            /* no live path in the whole block: every warp ran out of work, or the fetched tiles held no hit pixel */
            if (__syncthreads_and((exhausted && q_count == 0) ? 1 : 0))
                break;
            continue;
        }

        /* main ray -> stage A -> ray #1 -> stage B -> ray #2 -> stage C */
        if (!alive)
            request_none(ray);
        const B2Hit main_hit = b2_trace_pooled(ctx, pool, phase, ray, counters);
        request_none(ray);
        if (alive)
        {
            if (!megakernel_stage_a<V>(ctx, s, s.payload.bounce > 0 ? &main_hit : nullptr, ph, ray))
            {
                megakernel_end(ctx, s);
                alive = false;
                request_none(ray);
            }
        }
        const B2Hit result1 = b2_trace_pooled(ctx, pool, phase, ray, counters);
        request_none(ray);
        if (alive)
            megakernel_stage_b<V>(ctx, s, ph, result1, ray);
        const B2Hit result2 = b2_trace_pooled(ctx, pool, phase, ray, counters);
        request_none(ray);
        if (alive)
        {
            if (megakernel_stage_c<V>(ctx, s, ph, result1, result2))
                request_closest(ray, s.ray_origin, s.ray_direction, s.hit.primitive_index);
            else
            {
                megakernel_end(ctx, s);
                alive = false;
            }
        }
    }

    if (V::COUNTERS)
        b2_flush_counters(local_counters);
}

static int owned_rows(int height, const B2LaunchGeom& g)
{
    int rows = 0;
    for (int band = g.rank; band * B2_BAND_ROWS < height; band += g.world)
        rows += (height - band * B2_BAND_ROWS) < B2_BAND_ROWS ? (height - band * B2_BAND_ROWS) : B2_BAND_ROWS;
    return rows;
}

template <class V, int PASS>
cudaError_t launch_pass(const b2pt_render_data* rd, const B2LaunchGeom* geom, cudaStream_t stream)
{
    const B2LaunchGeom g = *geom;
    const int rows = owned_rows(g.height, g);
    if (rows == 0)
        return cudaSuccess;
    const int tiles_x = (g.width + B2_TILE_W - 1) / B2_TILE_W;
    if (PASS == B2_PASS_MEGAKERNEL && g.persistent)
    {
        /* per instantiation and per device: resident blocks per SM, SM count, and a ring of work counters (a launch
         * zeroes its own slot on its stream, so launches in flight on different streams never share a counter) */
        struct PerDevice { int blocks_per_sm[2] = { 0, 0 }, sm_count = 0; unsigned int* ring = nullptr; };
        static PerDevice per_device[B2_MAX_DEVICES];
        static std::mutex init_mutex;
        static std::atomic<unsigned int> next_slot{ 0 };
        int device = 0;
        cudaError_t e = cudaGetDevice(&device);
        if (e != cudaSuccess)
            return e;
        if (device < 0 || device >= B2_MAX_DEVICES)
            return cudaErrorInvalidDevice;
        PerDevice& pd = per_device[device];
        {
            std::lock_guard<std::mutex> lock(init_mutex);
            if (!pd.ring)
            {
                e = cudaDeviceGetAttribute(&pd.sm_count, cudaDevAttrMultiProcessorCount, device);
                if (e == cudaSuccess) e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&pd.blocks_per_sm[0], megakernel_persistent<V>, B2_MKP_THREADS, 0);
                if (e == cudaSuccess) e = cudaFuncSetAttribute(megakernel_phased<V>, cudaFuncAttributeMaxDynamicSharedMemorySize, B2_PH_SMEM_BYTES);
                if (e == cudaSuccess) e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&pd.blocks_per_sm[1], megakernel_phased<V>, B2_MK_THREADS, B2_PH_SMEM_BYTES);
                if (e == cudaSuccess && pd.blocks_per_sm[0] < 1) pd.blocks_per_sm[0] = 1;
                if (e == cudaSuccess && pd.blocks_per_sm[1] < 1) pd.blocks_per_sm[1] = 1;
                if (e == cudaSuccess) e = cudaMalloc(&pd.ring, B2_WORK_COUNTER_SLOTS * sizeof(unsigned int));
                if (e != cudaSuccess)
                    return e;
            }
        }
        unsigned int* counter = pd.ring + (next_slot.fetch_add(1) % B2_WORK_COUNTER_SLOTS);
        e = cudaMemsetAsync(counter, 0, sizeof(unsigned int), stream);
        if (e != cudaSuccess)
            return e;
        const unsigned int total_tiles = (unsigned int)tiles_x * (unsigned int)((rows + 3) / 4);
        if (g.persistent == 2)
            megakernel_phased<V><<<pd.sm_count * pd.blocks_per_sm[1], B2_MK_THREADS, B2_PH_SMEM_BYTES, stream>>>(*rd, g, tiles_x, total_tiles, counter);
        else
            megakernel_persistent<V><<<pd.sm_count * pd.blocks_per_sm[0], B2_MKP_THREADS, 0, stream>>>(*rd, g, tiles_x, total_tiles, counter);
        return cudaGetLastError();
    }
    const int tiles_y = (rows + B2_TILE_H - 1) / B2_TILE_H;
    pass_kernel<V, PASS><<<tiles_x * tiles_y, B2_BLOCK_THREADS, 0, stream>>>(*rd, g, tiles_x);
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
