/*
 * tools/traverse_selftest.cu -- GPU unit test of the CW8 traversal in isolation (no shading): traces the same rays
 * with the device code (product configuration: 64-thread blocks, shared-memory stack) and with the host build of
 * the very same function, and compares hit primitive / t bit for bit; reports guard trips and prints a step trace of
 * the first ray that trips the guard on the device.
 *   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -Iinclude -Ihiprt-path-tracer_b200/csrc \
 *        tools/traverse_selftest.cu hiprt-path-tracer_b200/csrc/b2_bvh_build.cpp -Xcompiler -fopenmp -o tools/traverse_selftest
 *   tools/traverse_selftest data/cornell_pbr_16x9.b2scene [n_rays]
 */
#include <cstdio>
#include <cstdlib>
#include <vector>

#include <cuda_runtime.h>

__device__ int g_trace_thread = -1;
#define B2_TRAVERSE_TRACE(guard, node, base, bits, sp) b2_selftest_trace(guard, node, base, bits, sp)
__host__ __device__ inline void b2_selftest_trace(unsigned guard, unsigned node, unsigned base, unsigned bits, int sp)
{
#if defined(__CUDA_ARCH__)
    if ((int)(blockIdx.x * blockDim.x + threadIdx.x) == g_trace_thread && guard < 80)
        printf("  step %u node %u group_base %u group_bits %08x sp %d\n", guard, node, base, bits, sp);
#endif
}

#include "b2_bvh_build.h"
#include "b2pt_scene_blob.h"

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { std::fprintf(stderr, "%s: %s\n", #x, cudaGetErrorString(e_)); std::exit(1); } } while (0)

struct Ray { float ox, oy, oz, dx, dy, dz; };
struct Res { float t; int prim; };

template <int MODE>
__global__ void __launch_bounds__(64) trace_kernel(B2BvhDevice bvh, const Ray* rays, Res* out, int n, unsigned int* faults)
{
    __shared__ uint64_t stacks[B2_STACK_ENTRIES * 64];
    uint64_t local_stack[B2_STACK_ENTRIES];
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n)
        return;
    Ray r = rays[i];
    B2Hit h;
    if (MODE == 0)
        h = b2_traverse<false>(bvh, v3(r.ox, r.oy, r.oz), v3(r.dx, r.dy, r.dz), 3.0e38f, -1, stacks + threadIdx.x, 64, nullptr, faults + 3 * i);
    else
        h = b2_traverse<false>(bvh, v3(r.ox, r.oy, r.oz), v3(r.dx, r.dy, r.dz), 3.0e38f, -1, local_stack, 1, nullptr, faults + 3 * i);
    out[i].t = h.t;
    out[i].prim = h.prim;
}

int main(int argc, char** argv)
{
    if (argc < 2) { std::fprintf(stderr, "usage: %s scene.b2scene [n_rays]\n", argv[0]); return 2; }
    int n = argc > 2 ? std::atoi(argv[2]) : 1 << 16;
    B2Scene scene;
    std::string err = b2scene_read(argv[1], scene);
    if (!err.empty()) { std::fprintf(stderr, "%s\n", err.c_str()); return 1; }
    B2BvhHost host;
    if (!b2_build_cw8(scene.indices.data(), scene.n_triangles(), scene.positions.data(), scene.n_vertices(), host, err)) { std::fprintf(stderr, "%s\n", err.c_str()); return 1; }
    std::printf("nodes %zu tris %zu depth %d\n", host.nodes.size(), host.tris.size(), host.stats.max_depth);

    /* rays: from the camera through a jittered grid, then from random points in the scene box in random directions */
    float lo[3] = { 1e30f, 1e30f, 1e30f }, hi[3] = { -1e30f, -1e30f, -1e30f };
    for (size_t v = 0; v < scene.n_vertices(); v++)
        for (int a = 0; a < 3; a++) { lo[a] = std::min(lo[a], scene.positions[3 * v + a]); hi[a] = std::max(hi[a], scene.positions[3 * v + a]); }
    std::vector<Ray> rays(n);
    uint32_t s = 12345u;
    auto rnd = [&]() { s ^= s << 13; s ^= s >> 17; s ^= s << 5; return (float)(s >> 8) / 16777216.0f; };
    for (int i = 0; i < n; i++)
    {
        Ray r;
        if (i < n / 2)
        {
            r.ox = scene.camera.position[0]; r.oy = scene.camera.position[1]; r.oz = scene.camera.position[2];
            float tx = lo[0] + (hi[0] - lo[0]) * rnd(), ty = lo[1] + (hi[1] - lo[1]) * rnd(), tz = lo[2] + (hi[2] - lo[2]) * rnd();
            r.dx = tx - r.ox; r.dy = ty - r.oy; r.dz = tz - r.oz;
        }
        else
        {
            r.ox = lo[0] + (hi[0] - lo[0]) * rnd(); r.oy = lo[1] + (hi[1] - lo[1]) * rnd(); r.oz = lo[2] + (hi[2] - lo[2]) * rnd();
            r.dx = rnd() * 2 - 1; r.dy = rnd() * 2 - 1; r.dz = rnd() * 2 - 1;
            if (i % 7 == 0) r.dx = 0.0f;
            if (i % 11 == 0) r.dy = 0.0f;
        }
        float l = std::sqrt(r.dx * r.dx + r.dy * r.dy + r.dz * r.dz);
        r.dx /= l; r.dy /= l; r.dz /= l;
        rays[i] = r;
    }

    /* host */
    std::vector<Res> ref(n);
    B2BvhDevice hb{ host.nodes.data(), host.tris.data(), (uint32_t)host.nodes.size(), (uint32_t)host.tris.size() };
    unsigned int host_faults[B2_FAULT_COUNT] = { 0, 0, 0 };
#pragma omp parallel for
    for (int i = 0; i < n; i++)
    {
        uint64_t st[B2_STACK_ENTRIES];
        B2Hit h = b2_traverse<false>(hb, v3(rays[i].ox, rays[i].oy, rays[i].oz), v3(rays[i].dx, rays[i].dy, rays[i].dz), 3.0e38f, -1, st, 1, nullptr, host_faults);
        ref[i].t = h.t; ref[i].prim = h.prim;
    }
    std::printf("host faults %u %u %u\n", host_faults[0], host_faults[1], host_faults[2]);

    B2Node* d_nodes; B2Tri* d_tris; Ray* d_rays; Res* d_out; unsigned int* d_faults;
    CK(cudaMalloc(&d_nodes, host.nodes.size() * sizeof(B2Node)));
    CK(cudaMalloc(&d_tris, host.tris.size() * sizeof(B2Tri)));
    CK(cudaMalloc(&d_rays, n * sizeof(Ray)));
    CK(cudaMalloc(&d_out, n * sizeof(Res)));
    CK(cudaMalloc(&d_faults, (size_t)n * 12));
    CK(cudaMemcpy(d_nodes, host.nodes.data(), host.nodes.size() * sizeof(B2Node), cudaMemcpyHostToDevice));
    CK(cudaMemcpy(d_tris, host.tris.data(), host.tris.size() * sizeof(B2Tri), cudaMemcpyHostToDevice));
    CK(cudaMemcpy(d_rays, rays.data(), n * sizeof(Ray), cudaMemcpyHostToDevice));
    B2BvhDevice db{ d_nodes, d_tris, (uint32_t)host.nodes.size(), (uint32_t)host.tris.size() };

    int rc = 0;
    for (int mode = 0; mode < 2; mode++)
    {
        CK(cudaMemset(d_faults, 0, (size_t)n * 12));
        if (mode == 0) trace_kernel<0><<<(n + 63) / 64, 64>>>(db, d_rays, d_out, n, d_faults);
        else trace_kernel<1><<<(n + 63) / 64, 64>>>(db, d_rays, d_out, n, d_faults);
        CK(cudaDeviceSynchronize());
        std::vector<Res> got(n);
        unsigned int faults[3] = { 0, 0, 0 };
        std::vector<unsigned int> tf((size_t)n * 3);
        CK(cudaMemcpy(got.data(), d_out, n * sizeof(Res), cudaMemcpyDeviceToHost));
        CK(cudaMemcpy(tf.data(), d_faults, (size_t)n * 12, cudaMemcpyDeviceToHost));
        int mismatch = 0, first = -1;
        for (int i = 0; i < n; i++)
        {
            for (int k = 0; k < 3; k++) faults[k] += tf[3 * (size_t)i + k];
            if (tf[3 * (size_t)i] && (first < 0 || !tf[3 * (size_t)first])) first = i; /* prefer a guard-tripping ray */
            if (got[i].prim != ref[i].prim || std::memcmp(&got[i].t, &ref[i].t, 4) != 0) { if (first < 0) first = i; mismatch++; }
        }
        std::printf("mode %d (%s stack): faults %u %u %u, mismatches vs host %d of %d (first %d)\n", mode, mode == 0 ? "shared" : "local", faults[0], faults[1], faults[2], mismatch, n, first);
        if (faults[0] || mismatch) rc = 1;
        if (first >= 0)
        {
            std::printf(" ray %d org %.9g %.9g %.9g dir %.9g %.9g %.9g host prim %d t %.9g device prim %d t %.9g\n", first, rays[first].ox, rays[first].oy, rays[first].oz,
                        rays[first].dx, rays[first].dy, rays[first].dz, ref[first].prim, ref[first].t, got[first].prim, got[first].t);
            CK(cudaMemcpyToSymbol(g_trace_thread, &first, sizeof(int)));
            if (mode == 0) trace_kernel<0><<<(n + 63) / 64, 64>>>(db, d_rays, d_out, n, d_faults);
            else trace_kernel<1><<<(n + 63) / 64, 64>>>(db, d_rays, d_out, n, d_faults);
            CK(cudaDeviceSynchronize());
            int off = -1;
            CK(cudaMemcpyToSymbol(g_trace_thread, &off, sizeof(int)));
        }
    }
    std::printf(rc ? "SELFTEST FAILED\n" : "SELFTEST OK\n");
    return rc;
}
