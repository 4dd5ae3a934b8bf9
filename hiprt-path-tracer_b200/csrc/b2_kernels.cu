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
#include <map>
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
#define B2_WORK_COUNTER_SLOTS 256
#define B2_MAX_DEVICES 64
/* phased MegaKernel (default): threads per block and resident blocks per SM it is compiled for (register cap 65536 / (T * N)) */
#ifndef B2_MK_THREADS
#define B2_MK_THREADS 512
#endif
#ifndef B2_MK_MIN_BLOCKS
#define B2_MK_MIN_BLOCKS 1
#endif
/* persistent MegaKernel without phases (kept for A/B measurements): 64-thread blocks, 12 resident per SM */
#define B2_MKP_THREADS 64
#define B2_MKP_MIN_BLOCKS 12
#define B2_MKP_BLOCK_SYNC 0

namespace
{
namespace B2_NS
{
static __device__ unsigned long long g_counters[3];
static __device__ unsigned int g_faults[B2_FAULT_COUNT];
/* CUDA kernels this object launched (bench.py: gpu_launches) */
static std::atomic<unsigned long long> g_cuda_launches{ 0 };

/* Pixel rows are partitioned across GPUs in interleaved bands of B2_BAND_ROWS rows (band b belongs to rank b % world);
 * a launch covers the bands of one rank. `local_row` counts the rows of this rank's bands, top to bottom. */
__device__ __forceinline__ uint32_t b2_global_row(uint32_t local_row, const B2LaunchGeom& g)
{
    if (g.row_map != nullptr) /* an explicit list of frame rows (own bands + halo rows): local rows beyond it map outside the frame */
        return local_row < (uint32_t)g.row_count ? (uint32_t)g.row_map[local_row] : 0xFFFFu;
    const uint32_t band_rows = (uint32_t)g.band_rows;
    const uint32_t band_local = local_row / band_rows;
    return (band_local * (uint32_t)g.world + (uint32_t)g.rank) * band_rows + local_row % band_rows;
}

__device__ __forceinline__ void b2_make_ctx(Ctx& ctx, const b2pt_render_data& rd, uint64_t* stacks, B2TraversalCounters* counters, unsigned int* faults, int threads = B2_BLOCK_THREADS,
                                            bool alpha_test = false)
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
    ctx.light_presampling = 0;
    ctx.alpha_test = alpha_test; /* kernels that were asked for alpha testing (B2PT_ALPHA_TEST) */
    ctx.bias_correction = 4;     /* RESTIR_DI_BIAS_CORRECTION_PAIRWISE_MIS; the pixel kernels of the reuse passes take the launch's */
    ctx.restir_options = 4 | 8;  /* B2_DI_OPT_DEFAULTS, likewise */
    B2_SET_RAY_BOUNCE(ctx, 0);
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
/* resident 64-thread blocks per SM the per-pixel pass kernels are compiled for (register cap 65536 / (64 * N)): their private
 * traversals are latency-bound, so occupancy is worth a few spills (measured: profiles/r02_results.md) */
#ifndef B2_PASS_MIN_BLOCKS
#define B2_PASS_MIN_BLOCKS 8
#endif
template <class V, int PASS>
__global__ void __launch_bounds__(B2_BLOCK_THREADS, B2_PASS_MIN_BLOCKS) pass_kernel(const __grid_constant__ b2pt_render_data rd, const B2LaunchGeom geom, int tiles_x)
{
    __shared__ uint64_t stacks[B2_STACK_ENTRIES * B2_BLOCK_THREADS];
    b2_block_constants_init(rd);

    const int tile = blockIdx.x;
    const int tile_x = tile % tiles_x, tile_y = tile / tiles_x;
    /* lane l of warp w: x = l & 7, y = 4 w + (l >> 3) */
    const uint32_t x = tile_x * B2_TILE_W + (threadIdx.x & 7);
    const uint32_t y = b2_global_row(tile_y * B2_TILE_H + (threadIdx.x >> 3), geom);

    B2TraversalCounters local_counters;
    local_counters.rays = local_counters.nodes = local_counters.tris = 0;

    Ctx ctx;
    b2_make_ctx(ctx, rd, stacks, V::COUNTERS ? &local_counters : nullptr, g_faults, B2_BLOCK_THREADS, geom.alpha_test != 0);
    ctx.light_presampling = geom.light_presampling;
    ctx.bias_correction = geom.bias_correction;
    ctx.restir_options = geom.restir_options;

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
 * Cornell frame, profiles/r01_ncu_c2_megakernel_first.md); handing out single pixels instead of tiles still spent a third of
 * the bounce loop's lanes on primary misses (profiles/r01_ncu_megakernel_history.md).
 * Per-pixel results do not depend on the schedule: every pixel owns its RNG stream and its framebuffer word.
 */
#define B2_QUEUE_SLOTS 64
template <class V>
__global__ void __launch_bounds__(B2_MKP_THREADS, B2_MKP_MIN_BLOCKS) megakernel_persistent(const __grid_constant__ b2pt_render_data rd, const B2LaunchGeom geom, int tiles_x,
                                                                                             unsigned int total_tiles, unsigned int* work_counter)
{
    __shared__ uint64_t stacks[B2_STACK_ENTRIES * B2_MKP_THREADS];
    __shared__ uint32_t queues[B2_MKP_THREADS / 32][B2_QUEUE_SLOTS];
    b2_block_constants_init(rd);

    B2TraversalCounters local_counters;
    local_counters.rays = local_counters.nodes = local_counters.tris = 0;

    Ctx ctx;
    b2_make_ctx(ctx, rd, stacks, V::COUNTERS ? &local_counters : nullptr, g_faults, B2_MKP_THREADS, geom.alpha_test != 0);

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
 * above, plus what its ncu profile asked for (profiles/: 6 of 32 lanes active inside the traversal, `no_instruction` = 70 %
 * of all stall samples because every warp of an SM sat in a different corner of a 180 KB kernel while the L1.5 instruction
 * cache holds 32 KB):
 *   1. the bounce is cut at its rays (b2_pathtrace.h: stage 1 / stage 2, b2_nee.h) and the rays of the WHOLE BLOCK are traced
 *      together: every thread drops up to three rays (the next main ray and the previous vertex's deferred NEE rays) into a
 *      shared-memory pool, then all threads drain the pool in one converged loop of traversal steps -- a lane whose ray is
 *      done takes the next ray from the pool while the other lanes keep stepping (warp-aggregated atomicAdd on the pool
 *      cursor), so the loop runs with full warps until the pool is empty, however many lanes asked for a ray and however
 *      long each ray is;
 *   2. block barriers around every trace keep the block's warps in the same stage of the code, so a stage's instructions
 *      are fetched once and reused by all of its warps.
 * One 512-thread block per SM (128 registers per thread: the path state stays in registers across the stages).
 * A path's arithmetic, RNG draws and results are the same as in the other two forms; tests/ run all three against the same
 * fixtures and against each other (bit-identical frames).
 */
#define B2_POOL_RAYS 5 /* main, deferred X, deferred Y, envmap light ray, envmap BSDF ray */
#ifndef B2_POOL_REFILL_MIN
#define B2_POOL_REFILL_MIN 1
#endif
struct B2RayPool
{
    float4* slot_a;  /* ray: origin.xyz, t_max            -> hit: t, u, v, prim      [B2_POOL_RAYS * threads] */
    float4* slot_b;  /* ray: direction.xyz, skip | flags  -> hit: ng.xyz             [B2_POOL_RAYS * threads] */
    int* list;       /* compacted slots of this phase's rays                        [B2_POOL_RAYS * threads] */
    int* count;      /* [2]: rays in the pool, by phase parity */
    int* next;       /* [2]: drain cursor, by phase parity */
};

/* A trace of the block's pooled rays is three steps, each called by ALL threads of the block the same number of times:
 *   b2_pool_deposit : every thread drops its rays of this phase into the pool (warp-aggregated slot allocation)
 *   b2_pool_drain   : barrier (which also carries the caller's block-wide vote), then one converged loop of traversal steps
 *                     over all rays of the block, barrier
 *   b2_pool_results : every thread reads the hits of the rays it deposited (prim = -1 for rays it did not ask for)
 * Deposit and drain are separate so that the rays a bounce ends with go straight from the shading code into shared memory
 * instead of living in registers across the top of the path loop (5 rays = 45 registers of a 128-register budget). `phase`
 * counts traces; its parity picks one of two counter pairs so that a deposit can reset the other pair without a barrier. */
__device__ __forceinline__ unsigned int b2_pool_deposit_one(const Ctx& ctx, const B2RayPool& pool, int pc, int k, const RayRequest& ray)
{
    const unsigned int lane = threadIdx.x & 31u;
    const unsigned int lanes_below = (1u << lane) - 1u;
    /* skip_prim shares a word with the any-hit flag: prim ids are < 2^30 */
    const bool have = ray.kind != B2_RAY_NONE && ctx.bvh.node_count != 0;
    const unsigned int havers = __ballot_sync(0xFFFFFFFFu, have);
    if (havers == 0u)
        return 0u;
    int base = 0;
    if (lane == 0)
        base = atomicAdd(&pool.count[pc], __popc(havers));
    base = __shfl_sync(0xFFFFFFFFu, base, 0);
    if (have)
    {
        const int slot = k * (int)blockDim.x + (int)threadIdx.x;
        pool.list[base + __popc(havers & lanes_below)] = slot;
        pool.slot_a[slot] = make_float4(ray.origin.x, ray.origin.y, ray.origin.z, ray.t_max);
        pool.slot_b[slot] = make_float4(ray.direction.x, ray.direction.y, ray.direction.z, __int_as_float((ray.skip_prim << 1) | (ray.kind == B2_RAY_ANY_HIT ? 1 : 0)));
    }
    return have ? (1u << k) : 0u;
}

/* `active_rays` (block-uniform): rays[k] with k >= active_rays are never requested in this launch (no envmap in the world) */
template <int N>
__device__ __forceinline__ unsigned int b2_pool_deposit(const Ctx& ctx, const B2RayPool& pool, unsigned int phase, const RayRequest (&rays)[N], int active_rays = N)
{
    static_assert(N <= B2_POOL_RAYS, "pool slots per thread");
    const int pc = (int)(phase & 1u);
    unsigned int have_mask = 0u;
#pragma unroll
    for (int k = 0; k < N; k++)
        if (k < active_rays)
            have_mask |= b2_pool_deposit_one(ctx, pool, pc, k, rays[k]);
    if (threadIdx.x == 0)
    {
        /* the other parity's counters are idle during this phase: reset them for the next one */
        pool.count[pc ^ 1] = 0;
        pool.next[pc ^ 1] = 0;
    }
    return have_mask;
}

__device__ __forceinline__ int b2_pool_drain(const Ctx& ctx, const B2RayPool& pool, unsigned int phase, B2TraversalCounters* counters, int stack_entries, int vote = 0)
{
    const unsigned int lane = threadIdx.x & 31u;
    const unsigned int lanes_below = (1u << lane) - 1u;
    const int pc = (int)(phase & 1u);
    /* the barrier between deposit and drain also carries the caller's block-wide vote (returned) */
    const int block_vote = __syncthreads_or(vote);

    /* one converged loop of traversal steps over all rays of the block */
    const int count = pool.count[pc];
    B2Traversal t;
    bool busy = false;
    bool drained = false; /* warp-uniform */
    int slot = 0;
    while (true)
    {
        const unsigned int idle = __ballot_sync(0xFFFFFFFFu, !busy);
        /* fetch when at least B2_POOL_REFILL_MIN lanes are idle (or none is busy): every fetch costs the whole warp a
         * ballot, an atomic and a ray setup, so single stragglers wait for company */
        if (idle != 0u && !drained && ((int)__popc(idle) >= B2_POOL_REFILL_MIN || idle == 0xFFFFFFFFu))
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
                    const float4 a = pool.slot_a[slot], b = pool.slot_b[slot];
                    const int packed = __float_as_int(b.w);
                    b2_traverse_begin(t, (packed & 1) != 0, v3(a.x, a.y, a.z), v3(b.x, b.y, b.z), a.w, packed >> 1, ctx.stack);
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
            if (!b2_traverse_step(t, ctx.bvh, ctx.stack_stride, ctx.faults, stack_entries))
            {
                pool.slot_a[slot] = make_float4(t.best.t, t.best.u, t.best.v, __int_as_float(t.best.prim));
                pool.slot_b[slot] = make_float4(t.best.ng.x, t.best.ng.y, t.best.ng.z, 0.0f);
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
    return block_vote;
}

template <int N>
__device__ __forceinline__ void b2_pool_results(const B2RayPool& pool, unsigned int have_mask, B2Hit (&results)[N])
{
#pragma unroll
    for (int k = 0; k < N; k++)
    {
        results[k].t = -1.0f;
        results[k].u = results[k].v = 0.0f;
        results[k].prim = -1;
        results[k].ng = v3(0.0f);
        if (have_mask & (1u << k))
        {
            const int own = k * (int)blockDim.x + (int)threadIdx.x;
            const float4 a = pool.slot_a[own], b = pool.slot_b[own];
            results[k].t = a.x;
            results[k].u = a.y;
            results[k].v = a.z;
            results[k].prim = __float_as_int(a.w);
            results[k].ng = v3(b.x, b.y, b.z);
        }
    }
}

/* shared memory of a phased block for a traversal stack of `entries` entries per thread: the stack is sized by the depth of the
 * BVH at hand (Cornell: 9 entries = 36 KB instead of 96 KB), which leaves more of the SM's 256 KB to the L1 that backs the
 * path state in local memory */
#define B2_PH_SMEM_BYTES_FOR(entries, pool_rays) ((entries) * B2_MK_THREADS * 8 + (B2_MK_THREADS / 32) * B2_QUEUE_SLOTS * 4 + (pool_rays) * B2_MK_THREADS * (2 * 16 + 4) + 16)
#define B2_PH_SMEM_BYTES B2_PH_SMEM_BYTES_FOR(B2_STACK_ENTRIES, B2_POOL_RAYS)

template <class V>
__global__ void __launch_bounds__(B2_MK_THREADS, B2_MK_MIN_BLOCKS) megakernel_phased(const __grid_constant__ b2pt_render_data rd, const B2LaunchGeom geom, int tiles_x,
                                                                                         unsigned int total_tiles, unsigned int* work_counter)
{
    extern __shared__ __align__(16) unsigned char smem[];
    uint64_t* stacks = reinterpret_cast<uint64_t*>(smem);
    B2RayPool pool;
    const int stack_entries = geom.stack_entries;
    /* rays per thread of a trace: 3 (main + the two NEE rays), 5 when the world has an envmap (its two MIS shadow rays) */
    const int pool_rays = geom.pool_rays;
    pool.slot_a = reinterpret_cast<float4*>(smem + stack_entries * B2_MK_THREADS * 8);
    pool.slot_b = pool.slot_a + pool_rays * B2_MK_THREADS;
    pool.list = reinterpret_cast<int*>(pool.slot_b + pool_rays * B2_MK_THREADS);
    uint32_t* queues = reinterpret_cast<uint32_t*>(pool.list + pool_rays * B2_MK_THREADS);
    pool.count = reinterpret_cast<int*>(queues + (B2_MK_THREADS / 32) * B2_QUEUE_SLOTS);
    pool.next = pool.count + 2;
    if (threadIdx.x < 4)
        pool.count[threadIdx.x] = 0; /* count[0..1], next[0..1] */
    b2_block_constants_init(rd); /* ends with the block barrier */

    B2TraversalCounters local_counters;
    local_counters.rays = local_counters.nodes = local_counters.tris = 0;
    B2TraversalCounters* counters = V::COUNTERS ? &local_counters : nullptr;

    Ctx ctx;
    b2_make_ctx(ctx, rd, stacks, counters, g_faults, B2_MK_THREADS, geom.alpha_test != 0);

    const unsigned int lane = threadIdx.x & 31u;
    const unsigned int lanes_below = (1u << lane) - 1u;
    uint32_t* queue = queues + (threadIdx.x >> 5) * B2_QUEUE_SLOTS;
    const int res_x = rd.render_settings.render_resolution.x, res_y = rd.render_settings.render_resolution.y;
    PathState s;
    BouncePhase<V> ph;
    ph.reset();
    /* rays of the NEXT trace A this thread has in the pool: bit 0 next main ray, 1 deferred X, 2 deferred Y, 3 envmap light ray,
     * 4 envmap BSDF ray (deposited at the end of the previous iteration, straight from stage 2) */
    unsigned int have_a = 0u;
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
                ph.reset();
                /* bounce 0 starts from the G-buffer hit: no main ray, nothing deferred (a lane is only refilled after its previous
                 * path ended, and a path never ends with rays in the pool) */
                have_a = 0u;
            }
        }
        {
            const unsigned int taken = n_idle < q_count ? n_idle : q_count;
            q_head = (q_head + taken) % B2_QUEUE_SLOTS;
            q_count -= taken;
        }
        __syncwarp();

        /* trace A: main ray + the previous vertex's deferred rays (already in the pool); its first barrier also votes "any live
         * path in the block?" */
        B2Hit results[5];
        const int any_alive = b2_pool_drain(ctx, pool, phase, counters, stack_entries, alive ? 1 : 0);
        phase++;
        if (!any_alive)
        {
            /* no live path in the whole block (so the pool was empty): every warp ran out of work, or the fetched tiles held no hit
             * pixel. The deposit that would have reset the other parity's counters is replayed empty. */
            RayRequest none[1];
            request_none(none[0]);
            if (__syncthreads_and((exhausted && q_count == 0) ? 1 : 0))
                break;
            b2_pool_deposit<1>(ctx, pool, phase, none);
            continue;
        }
        b2_pool_results<5>(pool, have_a, results);
        have_a = 0u;

        RayRequest sequential[1];
        request_none(sequential[0]);
        if (alive)
        {
            if (!megakernel_stage1<V>(ctx, s, ph, s.payload.bounce > 0 ? &results[0] : nullptr, results[1], results[2], results[3], results[4], sequential[0]))
            {
                megakernel_end(ctx, s);
                alive = false;
                request_none(sequential[0]);
            }
        }

        /* trace B: the sequential ray (RIS: closest hit along BSDF candidate 0) */
        B2Hit sequential_result[1];
        sequential_result[0].t = -1.0f;
        sequential_result[0].u = sequential_result[0].v = 0.0f;
        sequential_result[0].prim = -1;
        sequential_result[0].ng = v3(0.0f);
        if (Nee<V>::HAS_SEQUENTIAL_RAY)
        {
            const unsigned int have_b = b2_pool_deposit<1>(ctx, pool, phase, sequential);
            b2_pool_drain(ctx, pool, phase, counters, stack_entries);
            phase++;
            b2_pool_results<1>(pool, have_b, sequential_result);
        }

        /* stage 2 ends with the rays of the next trace A: they go straight into the pool */
        RayRequest rays[5];
#pragma unroll
        for (int k = 0; k < 5; k++)
            request_none(rays[k]);
        if (alive)
        {
            if (!megakernel_stage2<V>(ctx, s, ph, sequential_result[0], rays[0], rays[1], rays[2], rays[3], rays[4]))
            {
                megakernel_end(ctx, s);
                alive = false;
#pragma unroll
                for (int k = 0; k < 5; k++)
                    request_none(rays[k]);
            }
        }
        if (!Nee<V>::HAS_DEFERRED_Y)
            request_none(rays[2]);
        have_a = b2_pool_deposit<5>(ctx, pool, phase, rays, pool_rays);
    }

    if (V::COUNTERS)
        b2_flush_counters(local_counters);
}

#include "b2_wavefront.h"

/* ReSTIR_DI_LightsPresampling: one thread per presampled light (kernels/ReSTIR/DI/LightsPresampling.h:78-110); arguments in the
 * reference's order (LightPresamplingParameters, HIPRTRenderData) */
template <class V>
__global__ void __launch_bounds__(128) lights_presampling_kernel(const B2LightPresamplingParameters params, const __grid_constant__ b2pt_render_data rd, int uniform_strategy)
{
    restir_di_presample_light<V>(rd, params, blockIdx.x * blockDim.x + threadIdx.x, uniform_strategy != 0);
}

/* ReSTIR_Directional_Reuse_Compute: one thread per pixel, 8x8 tiles; arguments in the reference's order (render data, 32-bit masks
 * [unused: the DI instantiation keeps 64 direction bits], 64-bit masks, radii) */
__global__ void __launch_bounds__(B2_BLOCK_THREADS) directional_reuse_kernel(const __grid_constant__ b2pt_render_data rd, unsigned int*, unsigned long long* masks_ull, uint8_t* radius, int tiles_x)
{
    const int tile_x = blockIdx.x % tiles_x, tile_y = blockIdx.x / tiles_x;
    restir_directional_reuse_compute_pixel(rd, tile_x * B2_TILE_W + (threadIdx.x & 7), tile_y * B2_TILE_H + (threadIdx.x >> 3), masks_ull, radius);
}

static int owned_rows(int height, const B2LaunchGeom& g)
{
    if (g.row_map != nullptr)
        return g.row_count;
    int rows = 0;
    for (int band = g.rank; band * g.band_rows < height; band += g.world)
        rows += (height - band * g.band_rows) < g.band_rows ? (height - band * g.band_rows) : g.band_rows;
    return rows;
}

template <class V, int PASS>
cudaError_t launch_pass(const b2pt_render_data* rd, const B2LaunchGeom* geom, cudaStream_t stream)
{
    const B2LaunchGeom g = *geom;
    if (PASS == B2_PASS_RESTIR_DIRECTIONAL_REUSE)
    {
        if (!g.directional_masks_ull || !g.directional_radius)
            return cudaErrorInvalidValue;
        const int tx = (g.width + B2_TILE_W - 1) / B2_TILE_W, ty = (g.height + B2_TILE_H - 1) / B2_TILE_H;
        g_cuda_launches += 1;
        directional_reuse_kernel<<<tx * ty, B2_BLOCK_THREADS, 0, stream>>>(*rd, nullptr, (unsigned long long*)g.directional_masks_ull, (uint8_t*)g.directional_radius, tx);
        return cudaGetLastError();
    }
    if (PASS == B2_PASS_RESTIR_DI_PRESAMPLING)
    {
        if constexpr (V::DIRECT == B2_LSS_RESTIR_DI)
        {
            if (!g.presampling_parameters)
                return cudaErrorInvalidValue;
            const B2LightPresamplingParameters params = *reinterpret_cast<const B2LightPresamplingParameters*>(g.presampling_parameters);
            const int count = params.number_of_subsets * params.subset_size;
            if (count <= 0)
                return cudaSuccess;
            g_cuda_launches += 1;
            lights_presampling_kernel<V><<<(count + 127) / 128, 128, 0, stream>>>(params, *rd, g.light_presampling == 2 ? 1 : 0);
            return cudaGetLastError();
        }
        return cudaErrorInvalidValue;
    }
    const int rows = owned_rows(g.height, g);
    if (rows == 0)
        return cudaSuccess;
    const int tiles_x = (g.width + B2_TILE_W - 1) / B2_TILE_W;
    if (PASS == B2_PASS_MEGAKERNEL && g.persistent == 3)
        return launch_wavefront<V>(rd, g, rows, stream);
    if constexpr (V::DIRECT == B2_LSS_RESTIR_DI)
    {
        /* the ray-delimited ReSTIR DI passes follow the MegaKernel's form: frame-wide phases + the global ray queue on big scenes */
        const bool default_options = (g.restir_options & B2_DI_OPT_PHASE_CUT_MASK) == B2_DI_OPT_DEFAULTS;
        if (PASS == B2_PASS_RESTIR_DI_INITIAL_CANDIDATES && g.persistent == 3 && default_options)
            return launch_wavefront_pass<WfInitialCandidates<V>, V>(rd, g, rows, stream);
        /* (pairwise MIS with at most 6 neighbours; anything else runs the pass in one piece, restir_di_spatiotemporal_inplace_pixel) */
        if (PASS == B2_PASS_RESTIR_DI_SPATIOTEMPORAL && g.persistent == 3 && g.bias_correction == 4 && default_options && restir_di_spatiotemporal_supported<V>(*rd))
            return launch_wavefront_pass<WfSpatiotemporalReuse<V>, V>(rd, g, rows, stream);
    }
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
        g_cuda_launches += 1;
        if (g.persistent == 2)
            megakernel_phased<V><<<pd.sm_count * pd.blocks_per_sm[1], B2_MK_THREADS, B2_PH_SMEM_BYTES_FOR(g.stack_entries, g.pool_rays), stream>>>(*rd, g, tiles_x, total_tiles, counter);
        else
            megakernel_persistent<V><<<pd.sm_count * pd.blocks_per_sm[0], B2_MKP_THREADS, 0, stream>>>(*rd, g, tiles_x, total_tiles, counter);
        return cudaGetLastError();
    }
    const int tiles_y = (rows + B2_TILE_H - 1) / B2_TILE_H;
    g_cuda_launches += 1;
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

/* ReSTIR DI with another ReSTIR_DI_LaterBouncesSamplingStrategy: only the MegaKernel reads it (CameraRays and the ReSTIR DI passes are unit 4's and 5's) */
#define B2_REG_LATER(BSDF, BASE, LATER, COUNTERS) Reg<BSDF, B2_LSS_RESTIR_DI, BASE, LATER, B2_ESS_ALIAS_TABLE, COUNTERS>::entry<B2_PASS_MEGAKERNEL>()

#define B2_REG_RESTIR(BSDF, BASE, COUNTERS)                                                                                       \
    Reg<BSDF, B2_LSS_RESTIR_DI, BASE, B2_LATER_RIS, B2_ESS_ALIAS_TABLE, COUNTERS>::entry<B2_PASS_RESTIR_DI_PRESAMPLING>(),        \
    Reg<BSDF, B2_LSS_RESTIR_DI, BASE, B2_LATER_RIS, B2_ESS_ALIAS_TABLE, COUNTERS>::entry<B2_PASS_RESTIR_DIRECTIONAL_REUSE>(),     \
    Reg<BSDF, B2_LSS_RESTIR_DI, BASE, B2_LATER_RIS, B2_ESS_ALIAS_TABLE, COUNTERS>::entry<B2_PASS_RESTIR_DI_INITIAL_CANDIDATES>(), \
    Reg<BSDF, B2_LSS_RESTIR_DI, BASE, B2_LATER_RIS, B2_ESS_ALIAS_TABLE, COUNTERS>::entry<B2_PASS_RESTIR_DI_TEMPORAL>(),           \
    Reg<BSDF, B2_LSS_RESTIR_DI, BASE, B2_LATER_RIS, B2_ESS_ALIAS_TABLE, COUNTERS>::entry<B2_PASS_RESTIR_DI_SPATIAL>(),            \
    Reg<BSDF, B2_LSS_RESTIR_DI, BASE, B2_LATER_RIS, B2_ESS_ALIAS_TABLE, COUNTERS>::entry<B2_PASS_RESTIR_DI_SPATIOTEMPORAL>()

/* This translation unit is compiled once per (flavour, B2_UNIT, B2_UNIT_COUNTERS) -- build.py -- so that the variants build in
 * parallel; every object registers its entries (and its own copy of the counter / fault words) with b2_capi.cu at load time. */
#ifndef B2_UNIT
#error "define B2_UNIT (1 = c1, 2 = c2, 3 = c3, 4 = c4 MegaKernel + CameraRays, 5 = c4 ReSTIR DI passes, 6 / 7 / 8 = c4 MegaKernel with later bounces = one light / BSDF / MIS)"
#endif
#ifndef B2_UNIT_COUNTERS
#define B2_UNIT_COUNTERS 0
#endif
#define B2_UC (B2_UNIT_COUNTERS != 0)

static const B2KernelEntry* table(int* count)
{
    static const B2KernelEntry entries[] = {
#if B2_UNIT == 1
        B2_REG_BASIC(B2_BSDF_LAMBERTIAN, B2_LSS_ONE_LIGHT, B2_BASE_UNIFORM, B2_UC),
#elif B2_UNIT == 2
        B2_REG_BASIC(B2_BSDF_NONE, B2_LSS_MIS, B2_BASE_POWER, B2_UC),
#elif B2_UNIT == 3
        B2_REG_BASIC(B2_BSDF_NONE, B2_LSS_RIS, B2_BASE_POWER, B2_UC),
#elif B2_UNIT == 4
        B2_REG_BASIC(B2_BSDF_NONE, B2_LSS_RESTIR_DI, B2_BASE_POWER, B2_UC),
#elif B2_UNIT == 5
        B2_REG_RESTIR(B2_BSDF_NONE, B2_BASE_POWER, B2_UC),
#elif B2_UNIT == 6
        B2_REG_LATER(B2_BSDF_NONE, B2_BASE_POWER, B2_LATER_ONE_LIGHT, B2_UC),
#elif B2_UNIT == 7
        B2_REG_LATER(B2_BSDF_NONE, B2_BASE_POWER, B2_LATER_BSDF, B2_UC),
#elif B2_UNIT == 8
        B2_REG_LATER(B2_BSDF_NONE, B2_BASE_POWER, B2_LATER_MIS, B2_UC),
#else
#error "unknown B2_UNIT"
#endif
    };
    *count = (int)(sizeof(entries) / sizeof(entries[0]));
    return entries;
}

static cudaError_t unit_counters_reset()
{
    unsigned long long zero[3] = { 0, 0, 0 };
    return cudaMemcpyToSymbol(g_counters, zero, sizeof(zero));
}
static cudaError_t unit_counters_read(unsigned long long out[3]) { return cudaMemcpyFromSymbol(out, g_counters, 3 * sizeof(unsigned long long)); }
static cudaError_t unit_faults_read(unsigned int out[B2_FAULT_COUNT], int reset)
{
    cudaError_t e = cudaMemcpyFromSymbol(out, g_faults, B2_FAULT_COUNT * sizeof(unsigned int));
    if (e != cudaSuccess || !reset)
        return e;
    unsigned int zero[B2_FAULT_COUNT] = {};
    return cudaMemcpyToSymbol(g_faults, zero, sizeof(zero));
}

static unsigned long long unit_launches_read() { return g_cuda_launches.load(); }

struct Registrar
{
    Registrar()
    {
        B2KernelUnit u;
        u.entries = table(&u.count);
#if defined(B2_FLAVOUR_PRECISE)
        u.precise = 1;
#else
        u.precise = 0;
#endif
        u.counters_reset = &unit_counters_reset;
        u.counters_read = &unit_counters_read;
        u.faults_read = &unit_faults_read;
        u.launches_read = &unit_launches_read;
        b2_register_kernel_unit(&u);
    }
};
static Registrar g_registrar;
} // namespace B2_NS
} // anonymous namespace: every object of this file keeps its symbols to itself
