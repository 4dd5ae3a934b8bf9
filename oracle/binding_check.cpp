/*
 * oracle/binding_check.cpp -- TEST INFRASTRUCTURE ONLY, runs without a GPU. The reference's GPUKernel class as bound to libb2pt.so
 * (hiprt-path-tracer_b200/binding/GPUKernel_b2pt.cpp, compiled against the reference's own src/Compiler/GPUKernel.h) together with the
 * reference's real GPUKernelCompilerOptions (src/Compiler/GPUKernelCompilerOptions.cpp, compiled from where it lies):
 *   - every macro name the library understands must be a name the reference knows (GPUKernelCompilerOptions::ALL_MACROS_NAMES):
 *     a misspelt name on either side would silently never be passed;
 *   - every macro the library serves at one value only must have that value as the reference's default, i.e. a freshly constructed
 *     GPUKernelCompilerOptions must describe a variant that exists;
 *   - set_macro_value / synchronize_options_with / add_additional_macro_for_compilation behave through the binding as render passes use
 *     them (shared ints between a renderer's global options and its kernels).
 * Prints one JSON line; exit code 0 when everything holds.
 */
#include "Compiler/GPUKernel.h"
#include "HostDeviceCommon/KernelOptions/KernelOptions.h"

#include "b2pt.h"

#include <cstdio>

int main()
{
    int failures = 0;
    GPUKernel kernel("Device/kernels/Megakernel.h", "MegaKernel");
    kernel.parse_option_macros_used();

    const char* const* names = nullptr;
    int count = 0;
    if (b2pt_kernel_option_names(&names, &count) != B2PT_OK || count == 0)
    {
        std::fprintf(stderr, "b2pt_kernel_option_names failed: %s\n", b2pt_last_error_string());
        return 1;
    }
    int fixed = 0, switched = 0;
    for (int i = 0; i < count; i++)
    {
        const std::string name = names[i];
        if (!kernel.uses_macro(name))
        {
            std::fprintf(stderr, "the binding does not report '%s' as used\n", name.c_str());
            failures++;
        }
        if (GPUKernelCompilerOptions::ALL_MACROS_NAMES.find(name) == GPUKernelCompilerOptions::ALL_MACROS_NAMES.end() || kernel.get_kernel_options().get_pointer_to_macro_value(name) == nullptr)
        {
            std::fprintf(stderr, "'%s' is not an option macro of the reference\n", name.c_str());
            failures++;
            continue;
        }
        int value = 0;
        if (b2pt_kernel_option_fixed_value(name.c_str(), &value) == B2PT_OK)
        {
            fixed++;
            if (kernel.get_kernel_options().get_macro_value(name) != value)
            {
                std::fprintf(stderr, "'%s': the reference's default is %d, the kernels were built with %d\n", name.c_str(), kernel.get_kernel_options().get_macro_value(name), value);
                failures++;
            }
        }
        else
            switched++;
    }

    /* a renderer's global options shared with its kernels (GPURenderer.cpp:482-501 -> GPUKernel::synchronize_options_with) */
    auto global_options = std::make_shared<GPUKernelCompilerOptions>();
    GPUKernel camera_rays("Device/kernels/CameraRays.h", "CameraRays");
    camera_rays.synchronize_options_with(global_options, { GPUKernelCompilerOptions::BSDF_OVERRIDE });
    global_options->set_macro_value(GPUKernelCompilerOptions::DIRECT_LIGHT_SAMPLING_STRATEGY, LSS_RESTIR_DI);
    global_options->set_macro_value(GPUKernelCompilerOptions::BSDF_OVERRIDE, BSDF_LAMBERTIAN);
    if (camera_rays.get_kernel_options().get_macro_value(GPUKernelCompilerOptions::DIRECT_LIGHT_SAMPLING_STRATEGY) != LSS_RESTIR_DI)
    {
        std::fprintf(stderr, "a synchronised option did not follow the global options\n");
        failures++;
    }
    if (camera_rays.get_kernel_options().get_macro_value(GPUKernelCompilerOptions::BSDF_OVERRIDE) == BSDF_LAMBERTIAN)
    {
        std::fprintf(stderr, "an excluded option followed the global options\n");
        failures++;
    }
    camera_rays.add_additional_macro_for_compilation("ReSTIR_DI_DoLightPresampling", 1);
    const std::vector<std::string> additional = camera_rays.get_additional_compiler_macros();
    if (additional.size() != 1 || additional[0] != "-DReSTIR_DI_DoLightPresampling=1")
    {
        std::fprintf(stderr, "additional macros are not reported as -D<name>=<value>\n");
        failures++;
    }
    if (camera_rays.has_been_compiled() || camera_rays.compute_execution_time() != 0.0f)
    {
        std::fprintf(stderr, "a kernel that was never compiled / launched reports otherwise\n");
        failures++;
    }

    /* ---- the option surface, walked on the host: what GPUKernel::compile would pass for each combination must name a variant ---- */
    int combinations = 0;
    auto resolves = [&](const GPUKernelCompilerOptions& options, const char* entry) {
        std::vector<std::string> used;
        for (int i = 0; i < count; i++)
            if (options.get_pointer_to_macro_value(names[i]) != nullptr)
                used.push_back(names[i]);
        std::vector<b2pt_option> list;
        for (const std::string& n : used)
            list.push_back(b2pt_option{ n.c_str(), options.get_macro_value(n) });
        combinations++;
        return b2pt_kernel_query(entry, list.data(), (int)list.size()) == B2PT_OK;
    };
    auto expect = [&](bool ok, const GPUKernelCompilerOptions& options, const char* entry, const char* what) {
        if (resolves(options, entry) != ok)
        {
            std::fprintf(stderr, "%s / %s: %s (%s)\n", what, entry, ok ? "no variant" : "a variant was found where none is built", b2pt_last_error_string());
            failures++;
        }
    };
    const char* basic_entries[] = { "CameraRays", "MegaKernel" };
    const char* restir_entries[] = { "CameraRays", "MegaKernel", "ReSTIR_DI_InitialCandidates", "ReSTIR_DI_TemporalReuse", "ReSTIR_DI_SpatialReuse", "ReSTIR_DI_SpatiotemporalReuse",
                                     "ReSTIR_Directional_Reuse_Compute" };
    {
        GPUKernelCompilerOptions c3; /* the reference's defaults are BASELINE configs[2] */
        for (const char* e : basic_entries) expect(true, c3, e, "reference defaults (c3)");
        GPUKernelCompilerOptions c1;
        c1.set_macro_value(GPUKernelCompilerOptions::BSDF_OVERRIDE, BSDF_LAMBERTIAN);
        c1.set_macro_value(GPUKernelCompilerOptions::DIRECT_LIGHT_SAMPLING_STRATEGY, LSS_ONE_LIGHT);
        c1.set_macro_value(GPUKernelCompilerOptions::DIRECT_LIGHT_SAMPLING_BASE_STRATEGY, LSS_BASE_UNIFORM);
        for (const char* e : basic_entries) expect(true, c1, e, "c1");
        GPUKernelCompilerOptions c2;
        c2.set_macro_value(GPUKernelCompilerOptions::DIRECT_LIGHT_SAMPLING_STRATEGY, LSS_MIS_LIGHT_BSDF);
        for (const char* e : basic_entries) expect(true, c2, e, "c2");

        GPUKernelCompilerOptions c4;
        c4.set_macro_value(GPUKernelCompilerOptions::DIRECT_LIGHT_SAMPLING_STRATEGY, LSS_RESTIR_DI);
        for (const char* e : restir_entries) expect(true, c4, e, "c4");
        for (int weights = 0; weights <= 7; weights++)
        {
            GPUKernelCompilerOptions o = c4.deep_copy();
            o.set_macro_value(GPUKernelCompilerOptions::RESTIR_DI_BIAS_CORRECTION_WEIGHTS, weights);
            for (const char* e : { "ReSTIR_DI_TemporalReuse", "ReSTIR_DI_SpatialReuse", "ReSTIR_DI_SpatiotemporalReuse" })
                expect(weights != RESTIR_DI_BIAS_CORRECTION_1_OVER_M || std::string(e) != "ReSTIR_DI_SpatialReuse", o, e, "ReSTIR_DI_BiasCorrectionWeights"); /* 1/M: no spatial kernel */
        }
        for (int later = 0; later <= 3; later++)
        {
            GPUKernelCompilerOptions o = c4.deep_copy();
            o.set_macro_value(GPUKernelCompilerOptions::RESTIR_DI_LATER_BOUNCES_SAMPLING_STRATEGY, later);
            for (const char* e : restir_entries) expect(true, o, e, "ReSTIR_DI_LaterBouncesSamplingStrategy");
        }
        for (int bits = 0; bits < 16; bits++)
        {
            GPUKernelCompilerOptions o = c4.deep_copy();
            o.set_macro_value(GPUKernelCompilerOptions::RESTIR_DI_INITIAL_TARGET_FUNCTION_VISIBILITY, bits & 1);
            o.set_macro_value(GPUKernelCompilerOptions::RESTIR_DI_SPATIAL_TARGET_FUNCTION_VISIBILITY, (bits >> 1) & 1);
            o.set_macro_value(GPUKernelCompilerOptions::RESTIR_DI_DO_VISIBILITY_REUSE, (bits >> 2) & 1);
            o.set_macro_value(GPUKernelCompilerOptions::RESTIR_DI_BIAS_CORRECTION_USE_VISIBILITY, (bits >> 3) & 1);
            for (const char* e : restir_entries) expect(true, o, e, "ReSTIR DI visibility macros");
        }
        {
            GPUKernelCompilerOptions o = c4.deep_copy();
            o.set_macro_value(GPUKernelCompilerOptions::RESTIR_DI_DO_LIGHT_PRESAMPLING, KERNEL_OPTION_TRUE);
            for (const char* e : restir_entries) expect(true, o, e, "ReSTIR_DI_DoLightPresampling");
            expect(true, o, "ReSTIR_DI_LightsPresampling", "ReSTIR_DI_DoLightPresampling");
            o.set_macro_value("ReSTIR_DI_LightPresamplingStrategy", LSS_BASE_UNIFORM);
            expect(true, o, "ReSTIR_DI_LightsPresampling", "ReSTIR_DI_LightPresamplingStrategy");
            o.set_macro_value("ReSTIR_DI_LightPresamplingStrategy", LSS_BASE_REGIR);
            expect(false, o, "ReSTIR_DI_LightsPresampling", "ReSTIR_DI_LightPresamplingStrategy = ReGIR");
        }
        {
            GPUKernelCompilerOptions o = c4.deep_copy();
            o.set_macro_value(GPUKernelCompilerOptions::RESTIR_DI_DO_OPTIMAL_VISIBILITY_SAMPLING, KERNEL_OPTION_TRUE);
            for (const char* e : restir_entries) expect(true, o, e, "ReSTIR_DI_DoOptimalVisibilitySampling");
            GPUKernelCompilerOptions regir;
            regir.set_macro_value(GPUKernelCompilerOptions::DIRECT_LIGHT_SAMPLING_BASE_STRATEGY, LSS_BASE_REGIR);
            expect(false, regir, "MegaKernel", "ReGIR base strategy");
            GPUKernelCompilerOptions nee_plus_plus;
            nee_plus_plus.set_macro_value("DirectLightUseNEEPlusPlus", KERNEL_OPTION_TRUE);
            expect(false, nee_plus_plus, "MegaKernel", "NEE++");
            expect(false, c4, "ReSTIR_GI_SpatialReuse", "an entry this library does not serve");
        }
    }

    std::printf("{\"binding_check\": true, \"combinations_walked\": %d, \"library_macros\": %d, \"fixed\": %d, \"switched\": %d, \"reference_macros\": %zu, \"failures\": %d}\n", combinations, count, fixed, switched,
                GPUKernelCompilerOptions::ALL_MACROS_NAMES.size(), failures);
    return failures == 0 ? 0 : 1;
}
