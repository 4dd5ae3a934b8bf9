/*
 * b2_kernels_body.h -- the precompiled variant table and the pass -> per-pixel function dispatch.
 *
 * The reference selects strategies with -D macros at JIT time (src/Compiler/GPUKernelCompilerOptions.cpp:19-105);
 * the B200 build precompiles the combinations BASELINE.json's configs use, named after those configs:
 *   c1  BSDFOverride=BSDF_LAMBERTIAN, DirectLightSamplingStrategy=LSS_ONE_LIGHT,         BaseStrategy=LSS_BASE_UNIFORM
 *   c2  BSDF_NONE (principled),       LSS_MIS_LIGHT_BSDF,                                LSS_BASE_POWER
 *   c3  BSDF_NONE,                    LSS_RIS_BSDF_AND_LIGHT (the shipped default),      LSS_BASE_POWER
 *   c4  BSDF_NONE,                    LSS_RESTIR_DI (+ RIS at later bounces),            LSS_BASE_POWER
 * every other macro at the reference's default.
 */
#ifndef B2_KERNELS_BODY_H
#define B2_KERNELS_BODY_H

#include "b2_frame_loop.h"
#include "b2_pathtrace.h"
#include "b2_restir_di.h"

typedef B2Variant<B2_BSDF_LAMBERTIAN, B2_LSS_ONE_LIGHT, B2_BASE_UNIFORM> B2VariantC1;
typedef B2Variant<B2_BSDF_NONE, B2_LSS_MIS, B2_BASE_POWER> B2VariantC2;
typedef B2Variant<B2_BSDF_NONE, B2_LSS_RIS, B2_BASE_POWER> B2VariantC3;
typedef B2Variant<B2_BSDF_NONE, B2_LSS_RESTIR_DI, B2_BASE_POWER, B2_LATER_RIS> B2VariantC4;

template <class V>
B2_HD void b2_run_pass_pixel(int pass, const Ctx& ctx, uint32_t x, uint32_t y)
{
    switch (pass)
    {
    case B2_PASS_CAMERA_RAYS: camera_rays_pixel<V>(ctx, x, y); break;
    case B2_PASS_MEGAKERNEL: megakernel_pixel<V>(ctx, x, y); break;
    default:
        if constexpr (V::DIRECT == B2_LSS_RESTIR_DI)
            restir_di_pass_pixel<V>(pass, ctx, x, y);
        break;
    }
}

#endif
