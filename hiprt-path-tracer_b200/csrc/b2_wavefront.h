/*
 * b2_wavefront.h -- the MegaKernel pass as a WAVEFRONT of small kernels (included by b2_kernels.cu inside its namespaces).
 *
 * Same per-pixel code as the other forms (megakernel_begin / megakernel_stage1 / megakernel_stage2 / megakernel_end of
 * b2_pathtrace.h, i.e. the reference's Megakernel.h:22-144 cut at its rays) -- what changes is where a path lives between its
 * rays: in the phased form a path stays in the registers of one thread of a persistent 512-thread block for its whole life,
 * which caps the SM at 16 warps (ncu, profiles/r02_ncu_precise.md: issue slots 17 % busy, shading at 6.5 of 32 lanes because the
 * lanes of a warp hold 32 different materials, traversal waiting on dependent node fetches with 16 warps to hide them). Here the
 * path state is parked in HBM between stages (B200: 180 GB, 7 TB/s; ~0.7 KB per path) and every stage gets the launch shape it
 * wants:
 *   wf_trace   : ALL rays of the frame's current bounce in one queue, drained by persistent 64-register warps (32 warps / SM,
 *                traversal stack in shared memory) with the same converged step loop + dynamic refill as the block pool
 *   wf_stage1/2: one thread per live path, paths SORTED BY THE MATERIAL they are about to shade (counting sort on the material
 *                index of the main hit), so that the lanes of a warp run the same lobes of the principled BSDF
 * Paths never move between threads mid-stage, RNG streams and arithmetic are per path: frames are bit-identical to the other
 * forms (tests/test_gpu_parity.py runs all of them against the same fixtures and against each other).
 *
 * Per launch (one sample of every owned pixel): wf_begin, then nb_bounces + 2 rounds of
 *   wf_trace(A) -> wf_sort_count -> wf_sort_scan -> wf_sort_scatter -> wf_stage1 -> [wf_trace(B): RIS variants] -> wf_stage2
 * all on the caller's stream, sized by device-side counters (no host round trip; a kernel whose list is empty returns at once).
 */
#ifndef B2_WAVEFRONT_H
#define B2_WAVEFRONT_H

#define B2_WF_RAYS 5           /* ray slots per path: main, deferred X, deferred Y, envmap light, envmap BSDF (slot 0 doubles as the sequential ray) */
#define B2_WF_STAGE_THREADS 128
#ifndef B2_WF_STAGE1_MIN_BLOCKS
#define B2_WF_STAGE1_MIN_BLOCKS 3 /* resident 128-thread blocks per SM the shading kernels are compiled for (65536 / (128 N) registers) */
#endif
#ifndef B2_WF_STAGE2_MIN_BLOCKS
#define B2_WF_STAGE2_MIN_BLOCKS 3
#endif
#ifndef B2_WF_TRACE_THREADS
#define B2_WF_TRACE_THREADS 256
#endif
#ifndef B2_WF_TRACE_MIN_BLOCKS
#define B2_WF_TRACE_MIN_BLOCKS 4 /* 64 registers, 32 warps / SM */
#endif
#ifndef B2_WF_REFILL_LANES
#define B2_WF_REFILL_LANES 8 /* idle lanes of a warp that trigger a refill from the queue */
#endif
#ifndef B2_WF_TRI_LANES
#define B2_WF_TRI_LANES 8 /* lanes of a warp that must hold untested leaf triangles before the warp runs a triangle phase */
#endif
#define B2_WF_SORT_BUCKETS 1024 /* material index & 1023; the last bucket also takes paths whose main ray missed */

enum { WF_C_ALIVE = 0, WF_C_MID = 1, WF_C_NEXT = 2, WF_C_RAYS = 3, WF_C_CURSOR = 4, WF_C_COUNT = 8 };

template <class V>
struct alignas(16) WfPath
{
    PathState s;
    BouncePhase<V> ph;
    unsigned int have; /* rays of this path in flight: bits 0..4 = slots of trace A, bit 8 = the sequential ray (slot 0) of trace B */
};

/* the 16-byte pieces of a record that lie wholly inside the unpacked material of its payload */
template <class V>
struct WfMaterialPieces
{
    static constexpr int BEGIN_BYTE = (int)(offsetof(WfPath<V>, s) + offsetof(PathState, payload) + offsetof(Payload, material));
    static constexpr int BEGIN = (BEGIN_BYTE + 15) / 16;
    static constexpr int END = (BEGIN_BYTE + (int)sizeof(Mat)) / 16;
};

struct WfBuffers
{
    unsigned char* paths;   /* n_paths records of WfPath<V> */
    float4* slot_a;         /* n_paths * B2_WF_RAYS: ray (origin, t_max) -> hit (t, u, v, prim) */
    float4* slot_b;         /* ray (direction, skip | any-hit flag) -> hit (ng, 0) */
    int* ray_list;          /* slots to trace in the next wf_trace */
    int* list[3];           /* path ids: live paths (two, ping-pong) and the material-sorted copy */
    unsigned int* keys;     /* sort bucket per entry of the live list */
    unsigned int* counters; /* WF_C_* */
    unsigned int* buckets;  /* 2 * B2_WF_SORT_BUCKETS: histogram, then running offsets */
    int n_paths;
};

__device__ __forceinline__ void wf_path_pixel(int path, const B2LaunchGeom& g, uint32_t& x, uint32_t& y)
{
    x = (uint32_t)(path % g.width);
    y = b2_global_row((uint32_t)(path / g.width), g);
}

/* warp-aggregated append of `value` to list[counter++] by the lanes with `take` set; every lane of the warp must call it */
__device__ __forceinline__ void wf_append(int* list, unsigned int* counter, bool take, int value)
{
    const unsigned int lane = threadIdx.x & 31u;
    const unsigned int takers = __ballot_sync(0xFFFFFFFFu, take);
    if (takers == 0u)
        return;
    unsigned int base = 0;
    if (lane == (unsigned int)(__ffs(takers) - 1))
        base = atomicAdd(counter, (unsigned int)__popc(takers));
    base = __shfl_sync(0xFFFFFFFFu, base, __ffs(takers) - 1);
    if (take)
        list[base + __popc(takers & ((1u << lane) - 1u))] = value;
}

/* fourth word of a queued ray: (skip primitive << 3) | (B2_ALPHA_RAY_* << 1) | any-hit */
__device__ __forceinline__ float wf_ray_word(const RayRequest& ray, bool alpha_tested)
{
    const int alpha_ray = !alpha_tested ? B2_ALPHA_RAY_NONE : (ray.main_path ? B2_ALPHA_RAY_MAIN : B2_ALPHA_RAY_NEE);
    return __int_as_float((ray.skip_prim << 3) | (alpha_ray << 1) | (ray.kind == B2_RAY_ANY_HIT ? 1 : 0));
}

/* drops the thread's ray k into its slot and the trace list; returns the have-bit. `bounce` = bounce of the vertex the ray belongs
 * to (FilterFunctionPayload::bounce of the reference's trace call) */
__device__ __forceinline__ unsigned int wf_deposit(const Ctx& ctx, const WfBuffers& b, int path, int k, const RayRequest& ray, bool thread_active, int bounce)
{
    const bool have = thread_active && ray.kind != B2_RAY_NONE && ctx.bvh.node_count != 0;
    const int slot = path * B2_WF_RAYS + k;
    if (have)
    {
        b.slot_a[slot] = make_float4(ray.origin.x, ray.origin.y, ray.origin.z, ray.t_max);
        b.slot_b[slot] = make_float4(ray.direction.x, ray.direction.y, ray.direction.z, wf_ray_word(ray, b2_ctx_alpha_tested_at(ctx, bounce)));
    }
    wf_append(b.ray_list, b.counters + WF_C_RAYS, have, slot);
    return have ? (1u << k) : 0u;
}

__device__ __forceinline__ void wf_result(const WfBuffers& b, int path, int k, bool have, B2Hit& out)
{
    out.t = -1.0f;
    out.u = out.v = 0.0f;
    out.prim = -1;
    out.ng = v3(0.0f);
    if (have)
    {
        const float4 a = b.slot_a[path * B2_WF_RAYS + k], c = b.slot_b[path * B2_WF_RAYS + k];
        out.t = a.x;
        out.u = a.y;
        out.v = a.z;
        out.prim = __float_as_int(a.w);
        out.ng = v3(c.x, c.y, c.z);
    }
}


/* ---- a warp moves the records of its 32 paths between HBM and the threads' private copies in full 128-byte lines -------------
 * A thread copying its own ~0.6 KB record touches one 32-byte sector per lane and instruction (ncu on the 10M-triangle scene,
 * profiles/r02_ncu_wavefront.md: 134 load + 161 store sectors per path and stage, `lg_throttle` the second stall reason, 24 GB
 * through L2 in a 5.8 ms launch). Here 8 lanes fetch 128 contiguous bytes of ONE record, the warp 4 records per instruction, and
 * the 16-byte pieces change hands through a padded shared-memory tile (32 records x 128 B per warp). */
#ifndef B2_WF_SORT_ENDING_APART
#define B2_WF_SORT_ENDING_APART 1
#endif
#define B2_WF_COPY_TILE_UINT4 (32 * 9) /* per warp: 32 records x (8 + 1 pad) uint4 */
#define B2_WF_COPY_SMEM_BYTES ((B2_WF_STAGE_THREADS / 32) * B2_WF_COPY_TILE_UINT4 * 16)

/* pieces [SKIP_BEGIN, SKIP_END) of a record are left alone: the unpacked material (224 of 624 bytes) is dead between stage 2 and the
 * next stage 1, which gathers the material of the new hit */
template <int BYTES, int SKIP_BEGIN = 0, int SKIP_END = 0>
__device__ __forceinline__ void wf_records_load(uint4* __restrict__ mine, const unsigned char* __restrict__ records, int path, bool active, uint4* __restrict__ tile)
{
    static_assert(BYTES % 16 == 0, "records are whole uint4s");
    constexpr int PIECES = BYTES / 16;
    const unsigned int lane = threadIdx.x & 31u;
    const unsigned int actives = __ballot_sync(0xFFFFFFFFu, active);
#pragma unroll
    for (int chunk = 0; chunk * 8 < PIECES; chunk++)
    {
        if (chunk * 8 >= SKIP_BEGIN && chunk * 8 + 8 <= SKIP_END)
            continue; /* compile-time: the loop is unrolled */
#pragma unroll
        for (int it = 0; it < 8; it++)
        {
            const int owner = it * 4 + (int)(lane >> 3), part = chunk * 8 + (int)(lane & 7u);
            const int owner_path = __shfl_sync(0xFFFFFFFFu, path, owner);
            if (((actives >> owner) & 1u) && part < PIECES && !(part >= SKIP_BEGIN && part < SKIP_END))
                tile[owner * 9 + (lane & 7u)] = __ldcs(reinterpret_cast<const uint4*>(records + (size_t)owner_path * BYTES) + part);
        }
        __syncwarp();
        if (active)
        {
#pragma unroll
            for (int j = 0; j < 8; j++)
                if (chunk * 8 + j < PIECES && !(chunk * 8 + j >= SKIP_BEGIN && chunk * 8 + j < SKIP_END))
                    mine[chunk * 8 + j] = tile[lane * 9 + j];
        }
        __syncwarp();
    }
}

template <int BYTES, int SKIP_BEGIN = 0, int SKIP_END = 0>
__device__ __forceinline__ void wf_records_store(const uint4* __restrict__ mine, unsigned char* __restrict__ records, int path, bool write, uint4* __restrict__ tile)
{
    constexpr int PIECES = BYTES / 16;
    const unsigned int lane = threadIdx.x & 31u;
    const unsigned int writers = __ballot_sync(0xFFFFFFFFu, write);
    if (writers == 0u)
        return;
#pragma unroll
    for (int chunk = 0; chunk * 8 < PIECES; chunk++)
    {
        if (chunk * 8 >= SKIP_BEGIN && chunk * 8 + 8 <= SKIP_END)
            continue;
        if (write)
        {
#pragma unroll
            for (int j = 0; j < 8; j++)
                if (chunk * 8 + j < PIECES && !(chunk * 8 + j >= SKIP_BEGIN && chunk * 8 + j < SKIP_END))
                    tile[lane * 9 + j] = mine[chunk * 8 + j];
        }
        __syncwarp();
#pragma unroll
        for (int it = 0; it < 8; it++)
        {
            const int owner = it * 4 + (int)(lane >> 3), part = chunk * 8 + (int)(lane & 7u);
            const int owner_path = __shfl_sync(0xFFFFFFFFu, path, owner);
            if (((writers >> owner) & 1u) && part < PIECES && !(part >= SKIP_BEGIN && part < SKIP_END))
                __stcs(reinterpret_cast<uint4*>(records + (size_t)owner_path * BYTES) + part, tile[owner * 9 + (lane & 7u)]);
        }
        __syncwarp();
    }
}

/* Megakernel.h:33-79 for every owned pixel: primary-miss pixels are finished here, hit pixels become live paths */
template <class V>
__global__ void __launch_bounds__(B2_WF_STAGE_THREADS) wf_begin(const __grid_constant__ b2pt_render_data rd, const B2LaunchGeom geom, const WfBuffers b)
{
    extern __shared__ __align__(16) unsigned char wf_smem[];
    b2_block_constants_init(rd);
    Ctx ctx;
    b2_make_ctx(ctx, rd, reinterpret_cast<uint64_t*>(wf_smem), nullptr, g_faults, B2_WF_STAGE_THREADS, geom.alpha_test != 0);
    const int path = blockIdx.x * blockDim.x + threadIdx.x;
    bool alive = false;
    if (path < b.n_paths)
    {
        uint32_t x, y;
        wf_path_pixel(path, geom, x, y);
        const int res_x = rd.render_settings.render_resolution.x, res_y = rd.render_settings.render_resolution.y;
        if (x < (uint32_t)res_x && y < (uint32_t)res_y)
        {
            const uint32_t pixel_index = x + y * (uint32_t)res_x;
            if (rd.aux_buffers.pixel_active[pixel_index])
            {
                if (rd.g_buffer.first_hit_prim_index[pixel_index] == -1)
                    megakernel_primary_miss<V>(ctx, x, y, pixel_index);
                else
                {
                    WfPath<V>* record = reinterpret_cast<WfPath<V>*>(b.paths) + path;
                    PathState s;
                    alive = megakernel_begin<V>(ctx, x, y, s);
                    if (alive)
                    {
                        record->s = s;
                        record->ph.reset();
                        record->have = 0u;
                    }
                }
            }
        }
    }
    wf_append(b.list[0], b.counters + WF_C_ALIVE, alive, path);
}

/* sort bucket of every live path = the material it shades next: the material of the main ray's hit (bounce > 0: slot 0 holds the
 * traced result by now) or of the G-buffer's primary hit (bounce 0) */
template <class V>
__global__ void __launch_bounds__(256) wf_sort_count(const __grid_constant__ b2pt_render_data rd, const B2LaunchGeom geom, const WfBuffers b, int list_in)
{
    __shared__ unsigned int hist[B2_WF_SORT_BUCKETS];
    for (int i = threadIdx.x; i < B2_WF_SORT_BUCKETS; i += blockDim.x)
        hist[i] = 0u;
    __syncthreads();
    const unsigned int count = b.counters[WF_C_ALIVE];
    for (unsigned int i = blockIdx.x * blockDim.x + threadIdx.x; i < count; i += gridDim.x * blockDim.x)
    {
        const int path = b.list[list_in][i];
        const WfPath<V>* record = reinterpret_cast<const WfPath<V>*>(b.paths) + path;
        int prim;
        if (record->have & 1u)
            prim = __float_as_int(b.slot_a[path * B2_WF_RAYS].w);
        else
            prim = record->s.hit.primitive_index;
        /* paths that only come back to collect their last vertex's NEE (no main ray: they end in stage 1 after a few instructions)
         * and paths whose main ray left the scene get buckets of their own, so that they do not idle lanes of shading warps */
        unsigned int key;
        if (B2_WF_SORT_ENDING_APART && record->ph.ending)
            key = B2_WF_SORT_BUCKETS - 2;
        else
            key = prim >= 0 ? ((unsigned int)rd.buffers.material_indices[prim] % (B2_WF_SORT_BUCKETS - 2)) : (B2_WF_SORT_BUCKETS - 1);
        b.keys[i] = key;
        atomicAdd(&hist[key], 1u);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < B2_WF_SORT_BUCKETS; i += blockDim.x)
        if (hist[i])
            atomicAdd(&b.buckets[i], hist[i]);
}

/* one block: exclusive prefix of the histogram into the running offsets; also resets what the next stage appends to */
__global__ void __launch_bounds__(B2_WF_SORT_BUCKETS) wf_sort_scan(const WfBuffers b)
{
    __shared__ unsigned int scan[B2_WF_SORT_BUCKETS];
    const unsigned int mine = b.buckets[threadIdx.x];
    scan[threadIdx.x] = mine;
    __syncthreads();
    for (int offset = 1; offset < B2_WF_SORT_BUCKETS; offset <<= 1)
    {
        const unsigned int add = threadIdx.x >= (unsigned int)offset ? scan[threadIdx.x - offset] : 0u;
        __syncthreads();
        scan[threadIdx.x] += add;
        __syncthreads();
    }
    b.buckets[B2_WF_SORT_BUCKETS + threadIdx.x] = scan[threadIdx.x] - mine;
    b.buckets[threadIdx.x] = 0u; /* histogram cleared for the next round */
    if (threadIdx.x == 0)
    {
        /* trace A is over: stage 1 starts with empty lists */
        b.counters[WF_C_MID] = 0u;
        b.counters[WF_C_RAYS] = 0u;
        b.counters[WF_C_CURSOR] = 0u;
    }
}

__global__ void __launch_bounds__(256) wf_sort_scatter(const WfBuffers b, int list_in, int list_out)
{
    /* A block owns a contiguous chunk of the list: it counts its keys in shared memory, reserves one range per non-empty bucket
     * with ONE global atomic, and places its entries with shared-memory atomics -- a frame has tens of materials, so per-entry
     * global atomics on the bucket cursors serialise (measured: 0.38 ms per round, profiles/r02_launches_c4s.md). */
    __shared__ unsigned int local[B2_WF_SORT_BUCKETS];
    const unsigned int count = b.counters[WF_C_ALIVE];
    const unsigned int chunk = (count + gridDim.x - 1) / gridDim.x;
    const unsigned int begin = blockIdx.x * chunk, end = min(begin + chunk, count);
    if (begin >= end)
        return;
    for (int i = threadIdx.x; i < B2_WF_SORT_BUCKETS; i += blockDim.x)
        local[i] = 0u;
    __syncthreads();
    for (unsigned int i = begin + threadIdx.x; i < end; i += blockDim.x)
        atomicAdd(&local[b.keys[i]], 1u);
    __syncthreads();
    for (int i = threadIdx.x; i < B2_WF_SORT_BUCKETS; i += blockDim.x)
    {
        const unsigned int n = local[i];
        local[i] = n ? atomicAdd(&b.buckets[B2_WF_SORT_BUCKETS + i], n) : 0u; /* first position of this block's range in bucket i */
    }
    __syncthreads();
    for (unsigned int i = begin + threadIdx.x; i < end; i += blockDim.x)
    {
        const unsigned int position = atomicAdd(&local[b.keys[i]], 1u);
        b.list[list_out][position] = b.list[list_in][i];
    }
}

/* between trace B and stage 2 / at the end of a round: one thread moves the counters on */
__global__ void wf_advance(const WfBuffers b, int after_stage)
{
    if (after_stage == 1)
    {
        /* trace B is over: stage 2 appends the next round's rays and live paths */
        b.counters[WF_C_RAYS] = 0u;
        b.counters[WF_C_CURSOR] = 0u;
        b.counters[WF_C_NEXT] = 0u;
    }
    else
        b.counters[WF_C_ALIVE] = b.counters[WF_C_NEXT];
}

/* All rays of the queue, drained by persistent warps: a lane whose ray finished takes the next one while the others keep
 * stepping (the block pool's loop, fed from HBM instead of shared memory). */
template <bool COUNTERS>
__global__ void __launch_bounds__(B2_WF_TRACE_THREADS, B2_WF_TRACE_MIN_BLOCKS) wf_trace(const B2BvhDevice* __restrict__ bvh_descriptor, const WfBuffers b, int stack_entries,
                                                                                        const __grid_constant__ B2AlphaTest alpha)
{
    extern __shared__ __align__(16) unsigned char wf_smem[];
    B2BvhDevice bvh;
    bvh.nodes = nullptr;
    bvh.tris = nullptr;
    bvh.node_count = bvh.tri_count = 0;
    if (bvh_descriptor != nullptr)
        bvh = *bvh_descriptor;
    uint64_t* stack = reinterpret_cast<uint64_t*>(wf_smem) + threadIdx.x;
    const unsigned int lane = threadIdx.x & 31u;
    const unsigned int lanes_below = (1u << lane) - 1u;
    const int count = (int)b.counters[WF_C_RAYS];
    B2TraversalCounters local_counters;
    local_counters.rays = local_counters.nodes = local_counters.tris = 0;

    B2Traversal t;
    bool busy = false;
    bool drained = false; /* warp-uniform */
    int slot = 0;
    /* leaf triangles found by the lane's last node step, waiting for the warp's next triangle phase */
    uint32_t tri_base = 0, tri_bits = 0;
    while (true)
    {
        const unsigned int idle = __ballot_sync(0xFFFFFFFFu, !busy);
        if (idle != 0u && !drained && ((int)__popc(idle) >= B2_WF_REFILL_LANES || idle == 0xFFFFFFFFu))
        {
            const int wanted = __popc(idle);
            int base = 0;
            if (lane == 0)
                base = (int)atomicAdd(b.counters + WF_C_CURSOR, (unsigned int)wanted);
            base = __shfl_sync(0xFFFFFFFFu, base, 0);
            if (!busy)
            {
                const int i = base + __popc(idle & lanes_below);
                if (i < count)
                {
                    slot = b.ray_list[i];
                    const float4 a = b.slot_a[slot], c = b.slot_b[slot];
                    const int packed = __float_as_int(c.w);
                    b2_traverse_begin(t, (packed & 1) != 0, v3(a.x, a.y, a.z), v3(c.x, c.y, c.z), a.w, packed >> 3, stack, (packed >> 1) & 3);
                    tri_bits = 0;
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
        /* node phase: every lane without pending triangles takes its next node (or pops). The per-ray order node -> triangles ->
         * pop of b2_traverse_step is kept: a lane with pending triangles just waits, so results and counters do not change. */
        bool finished = false;
        if (busy && tri_bits == 0u)
        {
            if (!b2_traverse_node(t, bvh, B2_WF_TRACE_THREADS, g_faults, stack_entries, tri_base, tri_bits))
                finished = true;
            else if (tri_bits == 0u)
                finished = !b2_traverse_pop(t, B2_WF_TRACE_THREADS, stack_entries);
        }
        /* triangle phase: the exact tests run with 2 - 3 lanes when every lane does them right after its node (ncu, profiles/r02_ncu_
         * wavefront_before.md: 17 of 32 lanes per instruction over the kernel); here they wait until B2_WF_TRI_LANES lanes have
         * some, or until no lane of the warp can take a node */
        const unsigned int pending = __ballot_sync(0xFFFFFFFFu, busy && !finished && tri_bits != 0u);
        if (pending != 0u)
        {
            const unsigned int runnable = __ballot_sync(0xFFFFFFFFu, busy && !finished && tri_bits == 0u);
            if ((int)__popc(pending) >= B2_WF_TRI_LANES || runnable == 0u)
            {
                if (busy && !finished && tri_bits != 0u)
                {
                    const bool any_hit_done = b2_traverse_triangles(t, bvh, tri_base, tri_bits, alpha.enabled ? &alpha : nullptr);
                    tri_bits = 0u;
                    finished = any_hit_done || !b2_traverse_pop(t, B2_WF_TRACE_THREADS, stack_entries);
                }
            }
        }
        if (finished)
        {
            b.slot_a[slot] = make_float4(t.best.t, t.best.u, t.best.v, __int_as_float(t.best.prim));
            b.slot_b[slot] = make_float4(t.best.ng.x, t.best.ng.y, t.best.ng.z, 0.0f);
            if (COUNTERS)
            {
                local_counters.rays += 1;
                local_counters.nodes += t.n_nodes;
                local_counters.tris += t.n_tris;
            }
            busy = false;
        }
    }
    if (COUNTERS)
        b2_flush_counters(local_counters);
}

/* stage 1 of every live path, in material order: resolve the previous vertex's NEE, take the main hit, begin this vertex's NEE */
template <class V>
__global__ void __launch_bounds__(B2_WF_STAGE_THREADS, B2_WF_STAGE1_MIN_BLOCKS) wf_stage1(const __grid_constant__ b2pt_render_data rd, const B2LaunchGeom geom, const WfBuffers b, int list_in, int list_out, int first_round)
{
    extern __shared__ __align__(16) unsigned char wf_smem[];
    b2_block_constants_init(rd);
    B2TraversalCounters local_counters;
    local_counters.rays = local_counters.nodes = local_counters.tris = 0;
    /* The stages trace nothing themselves except in rare cases (a main hit that is a skipped nested-dielectric boundary, RIS with more
     * than one BSDF candidate): those traversals keep their stack in LOCAL memory, so that the block's shared memory is the copy tile
     * alone and the L1 -- which these kernels need for their 1.1 - 1.3 KB stack frames -- keeps the rest of the 256 KB */
    uint64_t local_stack[B2_STACK_ENTRIES];
    Ctx ctx;
    b2_make_ctx(ctx, rd, local_stack, V::COUNTERS ? &local_counters : nullptr, g_faults, 1, geom.alpha_test != 0);
    ctx.stack = local_stack;
    const unsigned int i = blockIdx.x * blockDim.x + threadIdx.x;
    const unsigned int count = b.counters[WF_C_ALIVE];
    if (blockIdx.x * blockDim.x >= count)
        return; /* block-uniform */
    const bool active = i < count;
    const int path = active ? b.list[list_in][i] : 0;
    bool alive = false;
    unsigned int have_b = 0u;
    uint4* tile = reinterpret_cast<uint4*>(wf_smem) + (threadIdx.x >> 5) * B2_WF_COPY_TILE_UINT4;
    WfPath<V> p;
    /* round 0 shades the primary hit with the material wf_begin unpacked from the G-buffer; every later round gathers the material of
     * the main ray's hit (trace_main_path_ray), so what the record holds there is dead */
    if (first_round)
        wf_records_load<(int)sizeof(WfPath<V>)>(reinterpret_cast<uint4*>(&p), b.paths, path, active, tile);
    else
        wf_records_load<(int)sizeof(WfPath<V>), WfMaterialPieces<V>::BEGIN, WfMaterialPieces<V>::END>(reinterpret_cast<uint4*>(&p), b.paths, path, active, tile);
    if (active)
    {
        B2Hit results[B2_WF_RAYS];
#pragma unroll
        for (int k = 0; k < B2_WF_RAYS; k++)
            wf_result(b, path, k, (p.have >> k) & 1u, results[k]);
        RayRequest sequential;
        request_none(sequential);
        alive = megakernel_stage1<V>(ctx, p.s, p.ph, p.s.payload.bounce > 0 ? &results[0] : nullptr, results[1], results[2], results[3], results[4], sequential);
        if (!alive)
        {
            megakernel_end(ctx, p.s);
            request_none(sequential);
        }
        p.have = 0u;
        if (alive)
        {
            /* the sequential ray takes slot 0 (the main ray's result has been consumed) */
            if (Nee<V>::HAS_SEQUENTIAL_RAY && sequential.kind != B2_RAY_NONE && ctx.bvh.node_count != 0)
            {
                const int slot = path * B2_WF_RAYS;
                b.slot_a[slot] = make_float4(sequential.origin.x, sequential.origin.y, sequential.origin.z, sequential.t_max);
                b.slot_b[slot] = make_float4(sequential.direction.x, sequential.direction.y, sequential.direction.z,
                                             wf_ray_word(sequential, b2_ctx_alpha_tested_at(ctx, p.s.payload.bounce)));
                have_b = 1u;
                p.have = 1u << 8;
            }
        }
    }
    wf_records_store<(int)sizeof(WfPath<V>)>(reinterpret_cast<const uint4*>(&p), b.paths, path, alive, tile);
    if (Nee<V>::HAS_SEQUENTIAL_RAY)
        wf_append(b.ray_list, b.counters + WF_C_RAYS, have_b != 0u, path * B2_WF_RAYS);
    wf_append(b.list[list_out], b.counters + WF_C_MID, alive, path);
    if (V::COUNTERS)
        b2_flush_counters(local_counters);
}

/* stage 2: rest of the NEE sampling, envmap NEE, next direction; ends with the rays of the next trace A */
template <class V>
__global__ void __launch_bounds__(B2_WF_STAGE_THREADS, B2_WF_STAGE2_MIN_BLOCKS) wf_stage2(const __grid_constant__ b2pt_render_data rd, const B2LaunchGeom geom, const WfBuffers b, int list_in, int list_out)
{
    extern __shared__ __align__(16) unsigned char wf_smem[];
    b2_block_constants_init(rd);
    B2TraversalCounters local_counters;
    local_counters.rays = local_counters.nodes = local_counters.tris = 0;
    uint64_t local_stack[B2_STACK_ENTRIES];
    Ctx ctx;
    b2_make_ctx(ctx, rd, local_stack, V::COUNTERS ? &local_counters : nullptr, g_faults, 1, geom.alpha_test != 0);
    ctx.stack = local_stack;
    const unsigned int i = blockIdx.x * blockDim.x + threadIdx.x;
    const unsigned int count = b.counters[WF_C_MID];
    if (blockIdx.x * blockDim.x >= count)
        return; /* block-uniform */
    const bool active = i < count;
    const int path = active ? b.list[list_in][i] : 0;
    bool alive = false;
    RayRequest rays[B2_WF_RAYS];
#pragma unroll
    for (int k = 0; k < B2_WF_RAYS; k++)
        request_none(rays[k]);
    uint4* tile = reinterpret_cast<uint4*>(wf_smem) + (threadIdx.x >> 5) * B2_WF_COPY_TILE_UINT4;
    WfPath<V> p;
    wf_records_load<(int)sizeof(WfPath<V>)>(reinterpret_cast<uint4*>(&p), b.paths, path, active, tile);
    if (active)
    {
        B2Hit sequential_result;
        wf_result(b, path, 0, (p.have >> 8) & 1u, sequential_result);
        alive = megakernel_stage2<V>(ctx, p.s, p.ph, sequential_result, rays[0], rays[1], rays[2], rays[3], rays[4]);
        if (!alive)
        {
            megakernel_end(ctx, p.s);
#pragma unroll
            for (int k = 0; k < B2_WF_RAYS; k++)
                request_none(rays[k]);
        }
        if (!Nee<V>::HAS_DEFERRED_Y)
            request_none(rays[2]);
    }
    unsigned int have = 0u;
#pragma unroll
    for (int k = 0; k < B2_WF_RAYS; k++)
        if (k < geom.pool_rays)
            have |= wf_deposit(ctx, b, path, k, rays[k], active && alive, k == 0 ? p.s.payload.bounce : p.ph.pending_bounce);
    p.have = have;
    wf_records_store<(int)sizeof(WfPath<V>), WfMaterialPieces<V>::BEGIN, WfMaterialPieces<V>::END>(reinterpret_cast<const uint4*>(&p), b.paths, path, active && alive, tile);
    wf_append(b.list[list_out], b.counters + WF_C_NEXT, alive, path);
    if (V::COUNTERS)
        b2_flush_counters(local_counters);
}

/* ---- the ReSTIR DI passes whose per-pixel code is cut at its rays (b2_restir_di.h: InitialCandidates, fused spatiotemporal reuse) ----
 * phase A over every owned pixel -> wf_trace -> phase B -> wf_trace -> phase C. One thread per pixel in 8x8 tiles like pass_kernel
 * (the neighbour gathers of a reuse pass stay local); a pixel's state between phases (40 - 200 B) is parked in HBM. */
template <class V>
struct WfInitialCandidates
{
    typedef DiInitialState State;
    static constexpr int RAYS = 1;
    static __device__ __forceinline__ void phase_a(const Ctx& ctx, uint32_t x, uint32_t y, State& st, RayRequest (&rays)[RAYS]) { restir_di_initial_phase_a<V>(ctx, x, y, st, rays[0]); }
    static __device__ __forceinline__ void phase_b(const Ctx& ctx, uint32_t x, uint32_t y, State& st, const B2Hit (&results)[RAYS], RayRequest& final_ray)
    {
        restir_di_initial_phase_b<V>(ctx, x, y, st, results[0], final_ray);
    }
    static __device__ __forceinline__ void phase_c(const Ctx& ctx, uint32_t x, uint32_t y, State& st, const B2Hit& final_hit) { restir_di_initial_phase_c<V>(ctx, x, y, st, final_hit); }
};

template <class V>
struct WfSpatiotemporalReuse
{
    typedef DiSpatiotemporalState State;
    /* slot k < 8: the canonical-sample ray of entry k; slot 8: the center-side ray of the temporal entry (always entry 0) */
    static constexpr int RAYS = B2_ST_MAX_ENTRIES + 1;
    static __device__ __forceinline__ void phase_a(const Ctx& ctx, uint32_t x, uint32_t y, State& st, RayRequest (&rays)[RAYS])
    {
        RayRequest all[B2_ST_RAYS];
        restir_di_spatiotemporal_phase_a<V>(ctx, x, y, st, all);
#pragma unroll
        for (int k = 0; k < B2_ST_MAX_ENTRIES; k++)
            rays[k] = all[2 * k + 1];
        rays[B2_ST_MAX_ENTRIES] = all[0];
    }
    static __device__ __forceinline__ void phase_b(const Ctx& ctx, uint32_t x, uint32_t y, State& st, const B2Hit (&results)[RAYS], RayRequest& final_ray)
    {
        B2Hit all[B2_ST_RAYS];
#pragma unroll
        for (int k = 0; k < B2_ST_MAX_ENTRIES; k++)
        {
            all[2 * k + 1] = results[k];
            all[2 * k] = results[k];
            all[2 * k].prim = -1; /* no center-side ray but entry 0's */
        }
        all[0] = results[B2_ST_MAX_ENTRIES];
        restir_di_spatiotemporal_phase_b<V>(ctx, x, y, st, all, final_ray);
    }
    static __device__ __forceinline__ void phase_c(const Ctx& ctx, uint32_t x, uint32_t y, State& st, const B2Hit& final_hit) { restir_di_spatiotemporal_phase_c<V>(ctx, x, y, st, final_hit); }
};

struct WfPassBuffers
{
    unsigned char* states;
    float4* slot_a;
    float4* slot_b;
    int* ray_list;
    unsigned int* counters;
    unsigned int* have; /* per pixel: bit k = ray k of phase A is in flight; bit 31 = the final ray */
    const int* pixel_list; /* the launch's pixels (slots) sorted by the material of their primary hit; thread i works on pixel_list[i] */
    int n_pixels;
};

/* Pixel of a thread. The phases of a pass run over the launch's pixels SORTED BY THE MATERIAL OF THEIR PRIMARY HIT (pixels without a
 * hit last): like the paths of the MegaKernel's shading stages, the lanes of a warp then evaluate the same lobes of the principled BSDF
 * (the stages lose 20 % when their paths come in pixel order instead, profiles/r02_results.md). slot = the pixel's index among the
 * launch's pixels (row-major over the launch's rows), -1 for a thread beyond the list. */
__device__ __forceinline__ int wf_list_pixel(const B2LaunchGeom& g, const WfPassBuffers& b, uint32_t& x, uint32_t& y)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    x = y = 0;
    if (i >= b.n_pixels)
        return -1;
    const int slot = b.pixel_list[i];
    x = (uint32_t)(slot % g.width);
    y = b2_global_row((uint32_t)(slot / g.width), g);
    return slot;
}

/* counting sort of the launch's pixels on the material of their primary hit (same three steps as the path sort above) */
__global__ void __launch_bounds__(256) wf_pixel_sort_count(const __grid_constant__ b2pt_render_data rd, const B2LaunchGeom geom, unsigned int* __restrict__ keys, unsigned int* __restrict__ buckets, int n_pixels)
{
    __shared__ unsigned int hist[B2_WF_SORT_BUCKETS];
    for (int i = threadIdx.x; i < B2_WF_SORT_BUCKETS; i += blockDim.x)
        hist[i] = 0u;
    __syncthreads();
    const int res_x = rd.render_settings.render_resolution.x, res_y = rd.render_settings.render_resolution.y;
    for (int slot = blockIdx.x * blockDim.x + threadIdx.x; slot < n_pixels; slot += gridDim.x * blockDim.x)
    {
        const uint32_t x = (uint32_t)(slot % geom.width), y = b2_global_row((uint32_t)(slot / geom.width), geom);
        unsigned int key = B2_WF_SORT_BUCKETS - 1;
        if (x < (uint32_t)res_x && y < (uint32_t)res_y)
        {
            const int prim = rd.g_buffer.first_hit_prim_index[x + y * (uint32_t)res_x];
            if (prim >= 0)
                key = (unsigned int)rd.buffers.material_indices[prim] % (B2_WF_SORT_BUCKETS - 2);
        }
        keys[slot] = key;
        atomicAdd(&hist[key], 1u);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < B2_WF_SORT_BUCKETS; i += blockDim.x)
        if (hist[i])
            atomicAdd(&buckets[i], hist[i]);
}

__global__ void __launch_bounds__(256) wf_pixel_sort_scatter(const unsigned int* __restrict__ keys, unsigned int* __restrict__ buckets, int* __restrict__ pixel_list, int n_pixels)
{
    __shared__ unsigned int local[B2_WF_SORT_BUCKETS];
    const unsigned int count = (unsigned int)n_pixels;
    const unsigned int chunk = (count + gridDim.x - 1) / gridDim.x;
    const unsigned int begin = blockIdx.x * chunk, end = min(begin + chunk, count);
    if (begin >= end)
        return;
    for (int i = threadIdx.x; i < B2_WF_SORT_BUCKETS; i += blockDim.x)
        local[i] = 0u;
    __syncthreads();
    for (unsigned int i = begin + threadIdx.x; i < end; i += blockDim.x)
        atomicAdd(&local[keys[i]], 1u);
    __syncthreads();
    for (int i = threadIdx.x; i < B2_WF_SORT_BUCKETS; i += blockDim.x)
    {
        const unsigned int n = local[i];
        local[i] = n ? atomicAdd(&buckets[B2_WF_SORT_BUCKETS + i], n) : 0u;
    }
    __syncthreads();
    for (unsigned int i = begin + threadIdx.x; i < end; i += blockDim.x)
        pixel_list[atomicAdd(&local[keys[i]], 1u)] = (int)i;
}

template <class P, class V>
__global__ void __launch_bounds__(B2_BLOCK_THREADS) wf_pass_a(const __grid_constant__ b2pt_render_data rd, const B2LaunchGeom geom, int tiles_x, int rows, const WfPassBuffers b)
{
    __shared__ uint64_t stacks[B2_STACK_ENTRIES * B2_BLOCK_THREADS];
    b2_block_constants_init(rd);
    B2TraversalCounters local_counters;
    local_counters.rays = local_counters.nodes = local_counters.tris = 0;
    Ctx ctx;
    b2_make_ctx(ctx, rd, stacks, V::COUNTERS ? &local_counters : nullptr, g_faults, B2_BLOCK_THREADS, geom.alpha_test != 0);
    ctx.light_presampling = geom.light_presampling;
    uint32_t x, y;
    const int slot = wf_list_pixel(geom, b, x, y);
    RayRequest rays[P::RAYS];
#pragma unroll
    for (int k = 0; k < P::RAYS; k++)
        request_none(rays[k]);
    typename P::State st;
    st.active = 0;
    if (slot >= 0)
    {
        P::phase_a(ctx, x, y, st, rays);
        reinterpret_cast<typename P::State*>(b.states)[slot] = st;
    }
    unsigned int have = 0u;
#pragma unroll
    for (int k = 0; k < P::RAYS; k++)
    {
        const bool take = slot >= 0 && st.active && rays[k].kind != B2_RAY_NONE && ctx.bvh.node_count != 0;
        const int ray_slot = slot * P::RAYS + k;
        if (take)
        {
            b.slot_a[ray_slot] = make_float4(rays[k].origin.x, rays[k].origin.y, rays[k].origin.z, rays[k].t_max);
            b.slot_b[ray_slot] = make_float4(rays[k].direction.x, rays[k].direction.y, rays[k].direction.z, wf_ray_word(rays[k], b2_ctx_alpha_tested_at(ctx, 0)));
            have |= 1u << k;
        }
        wf_append(b.ray_list, b.counters + WF_C_RAYS, take, ray_slot);
    }
    if (slot >= 0)
        b.have[slot] = have;
    if (V::COUNTERS)
        b2_flush_counters(local_counters);
}

template <class P, class V>
__global__ void __launch_bounds__(B2_BLOCK_THREADS) wf_pass_b(const __grid_constant__ b2pt_render_data rd, const B2LaunchGeom geom, int tiles_x, int rows, const WfPassBuffers b)
{
    __shared__ uint64_t stacks[B2_STACK_ENTRIES * B2_BLOCK_THREADS];
    b2_block_constants_init(rd);
    B2TraversalCounters local_counters;
    local_counters.rays = local_counters.nodes = local_counters.tris = 0;
    Ctx ctx;
    b2_make_ctx(ctx, rd, stacks, V::COUNTERS ? &local_counters : nullptr, g_faults, B2_BLOCK_THREADS, geom.alpha_test != 0);
    ctx.light_presampling = geom.light_presampling;
    uint32_t x, y;
    const int slot = wf_list_pixel(geom, b, x, y);
    RayRequest final_ray;
    request_none(final_ray);
    bool active = false;
    if (slot >= 0)
    {
        typename P::State st = reinterpret_cast<const typename P::State*>(b.states)[slot];
        active = st.active != 0;
        if (active)
        {
            const unsigned int have = b.have[slot];
            B2Hit results[P::RAYS];
#pragma unroll
            for (int k = 0; k < P::RAYS; k++)
            {
                results[k].t = -1.0f;
                results[k].u = results[k].v = 0.0f;
                results[k].prim = -1;
                results[k].ng = v3(0.0f);
                if ((have >> k) & 1u)
                {
                    const float4 ra = b.slot_a[slot * P::RAYS + k], rb = b.slot_b[slot * P::RAYS + k];
                    results[k].t = ra.x;
                    results[k].u = ra.y;
                    results[k].v = ra.z;
                    results[k].prim = __float_as_int(ra.w);
                    results[k].ng = v3(rb.x, rb.y, rb.z);
                }
            }
            P::phase_b(ctx, x, y, st, results, final_ray);
            reinterpret_cast<typename P::State*>(b.states)[slot] = st;
        }
    }
    const bool take = active && final_ray.kind != B2_RAY_NONE && ctx.bvh.node_count != 0;
    const int ray_slot = slot * P::RAYS;
    if (take)
    {
        b.slot_a[ray_slot] = make_float4(final_ray.origin.x, final_ray.origin.y, final_ray.origin.z, final_ray.t_max);
        b.slot_b[ray_slot] = make_float4(final_ray.direction.x, final_ray.direction.y, final_ray.direction.z, wf_ray_word(final_ray, b2_ctx_alpha_tested_at(ctx, 0)));
    }
    wf_append(b.ray_list, b.counters + WF_C_RAYS, take, ray_slot);
    if (slot >= 0)
        b.have[slot] = take ? (1u << 31) : 0u;
    if (V::COUNTERS)
        b2_flush_counters(local_counters);
}

template <class P, class V>
__global__ void __launch_bounds__(B2_BLOCK_THREADS) wf_pass_c(const __grid_constant__ b2pt_render_data rd, const B2LaunchGeom geom, int tiles_x, int rows, const WfPassBuffers b)
{
    Ctx ctx;
    b2_make_ctx(ctx, rd, nullptr, nullptr, g_faults, B2_BLOCK_THREADS, geom.alpha_test != 0);
    uint32_t x, y;
    const int slot = wf_list_pixel(geom, b, x, y);
    if (slot < 0)
        return;
    typename P::State st = reinterpret_cast<const typename P::State*>(b.states)[slot];
    if (!st.active)
        return;
    B2Hit final_hit;
    final_hit.t = -1.0f;
    final_hit.u = final_hit.v = 0.0f;
    final_hit.prim = -1;
    final_hit.ng = v3(0.0f);
    if (b.have[slot] >> 31)
        final_hit.prim = __float_as_int(b.slot_a[slot * P::RAYS].w);
    P::phase_c(ctx, x, y, st, final_hit);
}

__global__ void wf_reset_rays(unsigned int* counters)
{
    counters[WF_C_RAYS] = 0u;
    counters[WF_C_CURSOR] = 0u;
}

/* host side: scratch buffers per (device, stream) -- launches on different streams never share state --, grown on demand and kept
 * for the life of the library, and the launch sequences */
struct WfScratch
{
    void* base = nullptr;
    size_t bytes = 0;
    unsigned int* counters = nullptr;
    unsigned int* buckets = nullptr;
};

static cudaError_t wf_scratch_get(cudaStream_t stream, size_t total, WfScratch& out, int& out_sm_count)
{
    static std::map<std::pair<int, cudaStream_t>, WfScratch> scratch;
    static int sm_count[B2_MAX_DEVICES] = {};
    static std::mutex scratch_mutex;
    int device = 0;
    cudaError_t e = cudaGetDevice(&device);
    if (e != cudaSuccess)
        return e;
    if (device < 0 || device >= B2_MAX_DEVICES)
        return cudaErrorInvalidDevice;
    std::lock_guard<std::mutex> lock(scratch_mutex);
    if (!sm_count[device])
    {
        e = cudaDeviceGetAttribute(&sm_count[device], cudaDevAttrMultiProcessorCount, device);
        if (e != cudaSuccess)
            return e;
    }
    out_sm_count = sm_count[device];
    WfScratch& sc = scratch[std::make_pair(device, stream)];
    if (!sc.counters)
    {
        e = cudaMalloc((void**)&sc.counters, WF_C_COUNT * sizeof(unsigned int));
        if (e == cudaSuccess) e = cudaMalloc((void**)&sc.buckets, 2 * B2_WF_SORT_BUCKETS * sizeof(unsigned int));
        if (e != cudaSuccess)
            return e;
    }
    if (sc.bytes < total)
    {
        /* work of this stream may still be using the smaller buffer */
        e = cudaStreamSynchronize(stream);
        if (e != cudaSuccess)
            return e;
        cudaFree(sc.base);
        sc.base = nullptr;
        sc.bytes = 0;
        e = cudaMalloc(&sc.base, total);
        if (e != cudaSuccess)
            return e;
        sc.bytes = total;
    }
    out = sc;
    return cudaSuccess;
}

template <class P, class V>
cudaError_t launch_wavefront_pass(const b2pt_render_data* rd, const B2LaunchGeom& g, int rows, cudaStream_t stream)
{
    const size_t n_pixels = (size_t)rows * g.width;
    auto align = [](size_t v) { return (v + 255) & ~(size_t)255; };
    const size_t states_bytes = align(n_pixels * sizeof(typename P::State));
    const size_t slots_bytes = align(n_pixels * P::RAYS * sizeof(float4));
    const size_t ray_list_bytes = align(n_pixels * P::RAYS * sizeof(int));
    const size_t have_bytes = align(n_pixels * sizeof(unsigned int));
    const size_t list_bytes = align(n_pixels * sizeof(int));
    WfScratch sc;
    int sms = 0;
    cudaError_t e = wf_scratch_get(stream, states_bytes + 2 * slots_bytes + ray_list_bytes + have_bytes + 2 * list_bytes, sc, sms);
    if (e != cudaSuccess)
        return e;
    WfPassBuffers b;
    unsigned char* p = static_cast<unsigned char*>(sc.base);
    b.states = p; p += states_bytes;
    b.slot_a = reinterpret_cast<float4*>(p); p += slots_bytes;
    b.slot_b = reinterpret_cast<float4*>(p); p += slots_bytes;
    b.ray_list = reinterpret_cast<int*>(p); p += ray_list_bytes;
    b.have = reinterpret_cast<unsigned int*>(p); p += have_bytes;
    int* pixel_list = reinterpret_cast<int*>(p); p += list_bytes;
    unsigned int* pixel_keys = reinterpret_cast<unsigned int*>(p);
    b.pixel_list = pixel_list;
    b.n_pixels = (int)n_pixels;
    b.counters = sc.counters;
    WfBuffers tb; /* what wf_trace looks at */
    tb.paths = nullptr;
    tb.slot_a = b.slot_a;
    tb.slot_b = b.slot_b;
    tb.ray_list = b.ray_list;
    tb.list[0] = tb.list[1] = tb.list[2] = nullptr;
    tb.keys = nullptr;
    tb.counters = sc.counters;
    tb.buckets = sc.buckets;
    tb.n_paths = (int)n_pixels;

    e = cudaMemsetAsync(sc.counters, 0, WF_C_COUNT * sizeof(unsigned int), stream);
    if (e == cudaSuccess) e = cudaMemsetAsync(sc.buckets, 0, 2 * B2_WF_SORT_BUCKETS * sizeof(unsigned int), stream);
    if (e != cudaSuccess)
        return e;
    const int tiles_x = (g.width + B2_TILE_W - 1) / B2_TILE_W;
    const int pass_blocks = (int)((n_pixels + B2_BLOCK_THREADS - 1) / B2_BLOCK_THREADS);
    const int sort_blocks = sms * 4;
    const int trace_blocks = sms * B2_WF_TRACE_MIN_BLOCKS;
    const size_t trace_smem = (size_t)g.stack_entries * B2_WF_TRACE_THREADS * 8;
    const B2BvhDevice* bvh = reinterpret_cast<const B2BvhDevice*>(rd->GPU_BVH);
    const B2AlphaTest alpha = b2_alpha_test_of(*rd, g.alpha_test != 0);
    g_cuda_launches += 9;
    const bool prof = b2_wfprof_enabled();
    if (prof) b2_wfprof_mark(stream, B2_WF_K_SORT);
    wf_pixel_sort_count<<<sort_blocks, 256, 0, stream>>>(*rd, g, pixel_keys, sc.buckets, (int)n_pixels);
    wf_sort_scan<<<1, B2_WF_SORT_BUCKETS, 0, stream>>>(tb);
    wf_pixel_sort_scatter<<<sort_blocks, 256, 0, stream>>>(pixel_keys, sc.buckets, pixel_list, (int)n_pixels);
    if (prof) b2_wfprof_mark(stream, B2_WF_K_PASS_A);
    wf_pass_a<P, V><<<pass_blocks, B2_BLOCK_THREADS, 0, stream>>>(*rd, g, tiles_x, rows, b);
    if (prof) b2_wfprof_mark(stream, B2_WF_K_TRACE_PASS);
    wf_trace<V::COUNTERS><<<trace_blocks, B2_WF_TRACE_THREADS, trace_smem, stream>>>(bvh, tb, g.stack_entries, alpha);
    if (prof) b2_wfprof_mark(stream, B2_WF_K_ADVANCE);
    wf_reset_rays<<<1, 1, 0, stream>>>(sc.counters);
    if (prof) b2_wfprof_mark(stream, B2_WF_K_PASS_B);
    wf_pass_b<P, V><<<pass_blocks, B2_BLOCK_THREADS, 0, stream>>>(*rd, g, tiles_x, rows, b);
    if (prof) b2_wfprof_mark(stream, B2_WF_K_TRACE_PASS);
    wf_trace<V::COUNTERS><<<trace_blocks, B2_WF_TRACE_THREADS, trace_smem, stream>>>(bvh, tb, g.stack_entries, alpha);
    if (prof) b2_wfprof_mark(stream, B2_WF_K_PASS_C);
    wf_pass_c<P, V><<<pass_blocks, B2_BLOCK_THREADS, 0, stream>>>(*rd, g, tiles_x, rows, b);
    if (prof) b2_wfprof_mark(stream, B2_WF_K_END);
    return cudaGetLastError();
}

template <class V>
cudaError_t launch_wavefront(const b2pt_render_data* rd, const B2LaunchGeom& g, int rows, cudaStream_t stream)
{
    const size_t n_paths = (size_t)rows * g.width;
    auto align = [](size_t v) { return (v + 255) & ~(size_t)255; };
    const size_t paths_bytes = align(n_paths * sizeof(WfPath<V>));
    const size_t slots_bytes = align(n_paths * B2_WF_RAYS * sizeof(float4));
    const size_t ray_list_bytes = align(n_paths * B2_WF_RAYS * sizeof(int));
    const size_t list_bytes = align(n_paths * sizeof(int));
    WfScratch sc;
    int sms = 0;
    cudaError_t e = wf_scratch_get(stream, paths_bytes + 2 * slots_bytes + ray_list_bytes + 4 * list_bytes, sc, sms);
    if (e != cudaSuccess)
        return e;
    WfBuffers b;
    unsigned char* p = static_cast<unsigned char*>(sc.base);
    b.paths = p; p += paths_bytes;
    b.slot_a = reinterpret_cast<float4*>(p); p += slots_bytes;
    b.slot_b = reinterpret_cast<float4*>(p); p += slots_bytes;
    b.ray_list = reinterpret_cast<int*>(p); p += ray_list_bytes;
    for (int i = 0; i < 3; i++)
    {
        b.list[i] = reinterpret_cast<int*>(p);
        p += list_bytes;
    }
    b.keys = reinterpret_cast<unsigned int*>(p);
    b.counters = sc.counters;
    b.buckets = sc.buckets;
    b.n_paths = (int)n_paths;

    e = cudaMemsetAsync(sc.counters, 0, WF_C_COUNT * sizeof(unsigned int), stream);
    if (e == cudaSuccess) e = cudaMemsetAsync(sc.buckets, 0, 2 * B2_WF_SORT_BUCKETS * sizeof(unsigned int), stream);
    if (e != cudaSuccess)
        return e;

    const int stage_blocks = (int)((n_paths + B2_WF_STAGE_THREADS - 1) / B2_WF_STAGE_THREADS);
    const int sort_blocks = sms * 4;
    const int trace_blocks = sms * B2_WF_TRACE_MIN_BLOCKS;
    const size_t begin_smem = (size_t)g.stack_entries * B2_WF_STAGE_THREADS * 8;
    const size_t stage_smem = B2_WF_COPY_SMEM_BYTES; /* the record-copy tiles of the block's warps */
    const size_t trace_smem = (size_t)g.stack_entries * B2_WF_TRACE_THREADS * 8;
    const B2BvhDevice* bvh = reinterpret_cast<const B2BvhDevice*>(rd->GPU_BVH);
    const B2AlphaTest alpha = b2_alpha_test_of(*rd, g.alpha_test != 0);
    if (stage_smem > 48 * 1024)
    {
        e = cudaFuncSetAttribute(wf_stage1<V>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)stage_smem);
        if (e == cudaSuccess) e = cudaFuncSetAttribute(wf_stage2<V>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)stage_smem);
        if (e != cudaSuccess)
            return e;
    }

    const bool prof = b2_wfprof_enabled();
    if (prof) b2_wfprof_mark(stream, B2_WF_K_BEGIN);
    wf_begin<V><<<stage_blocks, B2_WF_STAGE_THREADS, begin_smem, stream>>>(*rd, g, b);
    int nb_bounces = rd->render_settings.nb_bounces;
    if (nb_bounces < 0)
        nb_bounces = 0;
    g_cuda_launches += 1 + (unsigned long long)(nb_bounces + 2) * (7 + (Nee<V>::HAS_SEQUENTIAL_RAY ? 1 : 0)) + (unsigned long long)(nb_bounces + 1);
    /* lists: [0] live paths of the round -> sorted into [2] -> stage-1 survivors appended to [1] -> stage-2 survivors appended to
     * [0] again (its old content has been consumed by the scatter before stage 1 started) */
    const int cur = 0, other = 1;
    for (int round = 0; round < nb_bounces + 2; round++)
    {
        if (round > 0)
        {
            if (prof) b2_wfprof_mark(stream, B2_WF_K_TRACE);
            wf_trace<V::COUNTERS><<<trace_blocks, B2_WF_TRACE_THREADS, trace_smem, stream>>>(bvh, b, g.stack_entries, alpha);
        }
        if (prof) b2_wfprof_mark(stream, B2_WF_K_SORT);
        wf_sort_count<V><<<sort_blocks, 256, 0, stream>>>(*rd, g, b, cur);
        wf_sort_scan<<<1, B2_WF_SORT_BUCKETS, 0, stream>>>(b);
        wf_sort_scatter<<<sort_blocks, 256, 0, stream>>>(b, cur, 2);
        if (prof) b2_wfprof_mark(stream, B2_WF_K_STAGE1);
        wf_stage1<V><<<stage_blocks, B2_WF_STAGE_THREADS, stage_smem, stream>>>(*rd, g, b, 2, other, round == 0); /* survivors -> list[other] */
        if (Nee<V>::HAS_SEQUENTIAL_RAY)
        {
            if (prof) b2_wfprof_mark(stream, B2_WF_K_TRACE);
            wf_trace<V::COUNTERS><<<trace_blocks, B2_WF_TRACE_THREADS, trace_smem, stream>>>(bvh, b, g.stack_entries, alpha);
        }
        if (prof) b2_wfprof_mark(stream, B2_WF_K_ADVANCE);
        wf_advance<<<1, 1, 0, stream>>>(b, 1);
        if (prof) b2_wfprof_mark(stream, B2_WF_K_STAGE2);
        wf_stage2<V><<<stage_blocks, B2_WF_STAGE_THREADS, stage_smem, stream>>>(*rd, g, b, other, cur); /* survivors -> list[cur] */
        if (prof) b2_wfprof_mark(stream, B2_WF_K_ADVANCE);
        wf_advance<<<1, 1, 0, stream>>>(b, 2);
    }
    if (prof) b2_wfprof_mark(stream, B2_WF_K_END);
    return cudaGetLastError();
}

#endif
