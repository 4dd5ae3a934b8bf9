/* b2_kernel_table.h -- registry of precompiled kernel variants (one table per codegen flavour). */
#ifndef B2_KERNEL_TABLE_H
#define B2_KERNEL_TABLE_H

#include "b2pt.h"
#include "b2_bvh.h"
#include "b2_frame_loop.h"

#include <cuda_runtime.h>

/* what a launch covers: the pixel rows of `rank`'s interleaved B2PT_BAND_ROWS-row bands of a width x height frame */
struct B2LaunchGeom
{
    int width, height;
    int rank, world;
    int band_rows;      /* interleaved bands of this many rows: band b belongs to rank b % world */
    const int* row_map; /* device array of row_count frame rows (ascending) this launch covers instead of rank's bands, or null */
    int row_count;
    int pool_rays;      /* phased MegaKernel: ray-pool slots per thread (3; 5 when the world has an envmap: its two MIS shadow rays) */
    int stack_entries; /* phased MegaKernel: traversal-stack entries per thread (BVH depth + 1 when known, else B2_STACK_ENTRIES) */
    int light_presampling; /* ReSTIR DI initial candidates: take light candidates from the presampled buffer (1; 2 = presampled uniformly, ReSTIR_DI_LightPresamplingStrategy) */
    int bias_correction;   /* ReSTIR_DI_BiasCorrectionWeights of the ReSTIR DI reuse passes (ReSTIRDIOptions.h:9-16; 4 = pairwise MIS, the default) */
    int restir_options;    /* ReSTIR DI passes: B2_DI_OPT_* bits (InitialTargetFunctionVisibility, SpatialTargetFunctionVisibility, DoVisibilityReuse, BiasCorrectionUseVisibility) */
    int alpha_test;        /* the scene has materials that are not fully opaque: candidate hits go through the alpha test (b2_bvh.h) */
    const void* presampling_parameters; /* ReSTIR_DI_LightsPresampling only: host pointer to its LightPresamplingParameters argument */
    void* directional_masks_ull; /* ReSTIR_Directional_Reuse_Compute only: its two output buffers (device pointers) */
    void* directional_radius;
    int persistent; /* MegaKernel only: 0 = one thread per pixel, 1 = persistent warps, 2 = persistent + phased (pooled rays), 3 = wavefront */
};

struct B2KernelEntry
{
    int pass;          /* B2Pass */
    int bsdf_override; /* variant key: the reference's macro values */
    int direct;
    int base;
    int later;
    int envmap;
    int counters;
    cudaError_t (*launch)(const b2pt_render_data* rd, const B2LaunchGeom* geom, cudaStream_t stream);
};

/* One object of b2_kernels.cu (a flavour x a group of variants, build.py) = one unit; units register themselves with
 * b2_capi.cu when the library is loaded. Each unit owns a private copy of the traversal counters and the fault words. */
struct B2KernelUnit
{
    const B2KernelEntry* entries;
    int count;
    int precise;
    cudaError_t (*counters_reset)();
    cudaError_t (*counters_read)(unsigned long long out[3]);
    cudaError_t (*faults_read)(unsigned int out[B2_FAULT_COUNT], int reset);
    unsigned long long (*launches_read)();
};
void b2_register_kernel_unit(const B2KernelUnit* unit);

/* Per-kernel timing of the wavefront launch sequences (b2_wavefront.h), off by default. When b2pt_wavefront_profile is on, every
 * kernel of a sequence is preceded by a CUDA event on the launching stream tagged with the kernel's kind; the time between two
 * consecutive events is attributed to the first one's kind (b2_capi.cu). */
enum B2WfKind { B2_WF_K_TRACE = 0, B2_WF_K_STAGE1, B2_WF_K_STAGE2, B2_WF_K_SORT, B2_WF_K_BEGIN, B2_WF_K_ADVANCE, B2_WF_K_PASS_A, B2_WF_K_PASS_B, B2_WF_K_PASS_C, B2_WF_K_TRACE_PASS, B2_WF_K_END, B2_WF_K_COUNT = B2_WF_K_END };
bool b2_wfprof_enabled();
void b2_wfprof_mark(cudaStream_t stream, int kind);

#endif
