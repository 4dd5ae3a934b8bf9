/*
 * binding/GPUKernel_b2pt.cpp -- the reference's GPUKernel class on the B200 library.
 *
 * This is the file a maintainer of the reference compiles INSTEAD OF src/Compiler/GPUKernel.cpp: every member function the reference's
 * own, unmodified header (src/Compiler/GPUKernel.h:21-176) declares, implemented on include/b2pt.h. Nothing above GPUKernel changes:
 * render passes keep calling kernel.get_kernel_options().set_macro_value(...), kernel.compile(ctx), kernel.launch_asynchronous(8, 8,
 * w, h, { &render_data }, stream) and kernel.compute_execution_time() (MegaKernelRenderPass.cpp:20-60, ReSTIRDIRenderPass.cpp:21-140).
 *
 * Two conventions carry the library's handles through the reference's types without touching them:
 *   - HIPRTOrochiCtx::orochi_ctx (an opaque oroCtx) holds the b2pt_ctx* the application created with b2pt_ctx_create;
 *   - GPUKernel::m_kernel_function (an opaque oroFunction) holds the b2pt_kernel* of the precompiled variant; the two oroEvent_t
 *     members stay null, the library times its launches itself (b2pt_kernel_last_ms).
 * "Compilation" is a table lookup: the macros b2pt_kernel_get understands (b2pt_kernel_option_names) are read from the kernel's
 * GPUKernelCompilerOptions and passed by name; a combination that was not built ahead of time is reported the way a failed JIT
 * compilation is (message + has_been_compiled() == false).
 *
 * Only compiled where the reference's headers are present (oracle/Makefile: boundary_gpukernel_*), by the recipe committed there.
 */
#include "Compiler/GPUKernel.h"

#include "b2pt.h"

#include <cstdio>
#include <cstdlib>

const std::vector<std::string> GPUKernel::COMMON_ADDITIONAL_KERNEL_INCLUDE_DIRS = {};

namespace
{
b2pt_kernel* handle_of(oroFunction function) { return reinterpret_cast<b2pt_kernel*>(function); }

void report(const char* what, const std::string& kernel_name)
{
    std::fprintf(stderr, "GPUKernel[%s]: %s: %s\n", kernel_name.c_str(), what, b2pt_last_error_string());
}
} // namespace

GPUKernel::GPUKernel() {}

GPUKernel::GPUKernel(const std::string& kernel_file_path, const std::string& kernel_function_name) : GPUKernel()
{
    m_kernel_file_path = kernel_file_path;
    m_kernel_function_name = kernel_function_name;
}

std::string GPUKernel::get_kernel_file_path() const { return m_kernel_file_path; }
std::string GPUKernel::get_kernel_function_name() const { return m_kernel_function_name; }
void GPUKernel::set_kernel_file_path(const std::string& kernel_file_path) { m_kernel_file_path = kernel_file_path; }
void GPUKernel::set_kernel_function_name(const std::string& kernel_function_name) { m_kernel_function_name = kernel_function_name; }

/* GPUKernel.cpp:72-83: hiprtBuildTraceKernels with -D<macro>=<value> for every macro the kernel uses; here: the variant table */
void GPUKernel::compile(std::shared_ptr<HIPRTOrochiCtx> hiprt_ctx, std::vector<hiprtFuncNameSet>, bool, bool silent)
{
    if (m_option_macro_invalidated)
        parse_option_macros_used();

    std::vector<std::string> names(m_used_option_macros.begin(), m_used_option_macros.end());
    std::vector<b2pt_option> options;
    for (const std::string& name : names)
        if (m_compiler_options.get_pointer_to_macro_value(name) != nullptr) /* (has_macro() only answers for custom macros, GPUKernelCompilerOptions.cpp:387-393) */
            options.push_back(b2pt_option{ name.c_str(), m_compiler_options.get_macro_value(name) });
    /* per-kernel additions (add_additional_macro_for_compilation) the library has a name for */
    for (const auto& macro : m_additional_compilation_macros)
        if (m_used_option_macros.count(macro.first))
            options.push_back(b2pt_option{ macro.first.c_str(), macro.second });

    if (m_kernel_function != nullptr)
    {
        b2pt_kernel_destroy(handle_of(m_kernel_function));
        m_kernel_function = nullptr;
    }
    b2pt_kernel* kernel = nullptr;
    b2pt_ctx* ctx = reinterpret_cast<b2pt_ctx*>(hiprt_ctx->orochi_ctx);
    if (b2pt_kernel_get(ctx, m_kernel_function_name.c_str(), options.data(), (int)options.size(), &kernel) != B2PT_OK)
    {
        if (!silent)
            report("no precompiled variant", m_kernel_function_name);
        return;
    }
    m_kernel_function = reinterpret_cast<oroFunction>(kernel);
}

void GPUKernel::launch(int tile_size_x, int tile_size_y, int res_x, int res_y, void** launch_args, oroStream_t stream)
{
    if (m_kernel_function == nullptr)
    {
        std::fprintf(stderr, "GPUKernel[%s]: launched before a successful compile()\n", m_kernel_function_name.c_str());
        std::exit(1);
    }
    if (b2pt_kernel_launch(handle_of(m_kernel_function), tile_size_x, tile_size_y, res_x, res_y, launch_args, reinterpret_cast<b2pt_stream>(stream)) != B2PT_OK)
    {
        report("launch failed", m_kernel_function_name);
        std::exit(1);
    }
    m_launched_at_least_once = true;
}

void GPUKernel::launch_3D(int, int, int, int, int, int, void**, oroStream_t)
{
    /* the reference's 3D launches belong to ReGIR's grid fill, which this library does not serve */
    std::fprintf(stderr, "GPUKernel[%s]: 3D launches are not served by libb2pt\n", m_kernel_function_name.c_str());
    std::exit(1);
}

void GPUKernel::launch_synchronous(int tile_size_x, int tile_size_y, int res_x, int res_y, void** launch_args, float* execution_time_out)
{
    launch(tile_size_x, tile_size_y, res_x, res_y, launch_args, nullptr);
    float ms = 0.0f;
    if (b2pt_kernel_last_ms(handle_of(m_kernel_function), &ms) != B2PT_OK) /* synchronises on the launch's stop event */
        report("timing failed", m_kernel_function_name);
    if (execution_time_out != nullptr)
        *execution_time_out = ms;
}

void GPUKernel::launch_asynchronous(int tile_size_x, int tile_size_y, int res_x, int res_y, void** launch_args, oroStream_t stream)
{
    launch(tile_size_x, tile_size_y, res_x, res_y, launch_args, stream);
}

void GPUKernel::launch_asynchronous_3D(int tile_size_x, int tile_size_y, int tile_size_z, int res_x, int res_y, int res_z, void** launch_args, oroStream_t stream)
{
    if (tile_size_z == 1 && res_z == 1)
        launch(tile_size_x, tile_size_y, res_x, res_y, launch_args, stream);
    else
        launch_3D(tile_size_x, tile_size_y, tile_size_z, res_x, res_y, res_z, launch_args, stream);
}

void GPUKernel::add_additional_macro_for_compilation(const std::string& name, int value) { m_additional_compilation_macros[name] = value; }

std::vector<std::string> GPUKernel::get_additional_compiler_macros() const
{
    std::vector<std::string> macros;
    for (const auto& macro : m_additional_compilation_macros)
        macros.push_back("-D" + macro.first + "=" + std::to_string(macro.second));
    return macros;
}

/* GPUKernel.cpp:176-180 parses the kernel's source for the option macros it reads; the precompiled kernels answer for themselves */
void GPUKernel::parse_option_macros_used()
{
    const char* const* names = nullptr;
    int count = 0;
    m_used_option_macros.clear();
    if (b2pt_kernel_option_names(&names, &count) == B2PT_OK)
        for (int i = 0; i < count; i++)
            m_used_option_macros.insert(names[i]);
    m_option_macro_invalidated = false;
}

bool GPUKernel::uses_macro(const std::string& macro_name) const { return m_used_option_macros.find(macro_name) != m_used_option_macros.end(); }

/* register / shared-memory attributes of a JIT-compiled function: the precompiled variants do not expose them */
int GPUKernel::get_kernel_attribute(oroFunction_attribute) const { return 0; }
int GPUKernel::get_kernel_attribute(oroFunction, oroFunction_attribute) { return 0; }

GPUKernelCompilerOptions& GPUKernel::get_kernel_options() { return m_compiler_options; }
const GPUKernelCompilerOptions& GPUKernel::get_kernel_options() const { return m_compiler_options; }

/* GPUKernel.cpp:115-136: this kernel's options become the SAME ints as the other set's (shared pointers), so that a change in the
 * renderer's global options is seen by every kernel at its next compile() */
void GPUKernel::synchronize_options_with(std::shared_ptr<GPUKernelCompilerOptions> other_options, const std::unordered_set<std::string>& options_excluded)
{
    for (const auto& macro : other_options->get_options_macro_map())
        if (options_excluded.find(macro.first) == options_excluded.end())
            m_compiler_options.set_pointer_to_macro(macro.first, other_options->get_pointer_to_macro_value(macro.first));
    for (const auto& macro : other_options->get_custom_macro_map())
        if (options_excluded.find(macro.first) == options_excluded.end())
            m_compiler_options.set_pointer_to_macro(macro.first, other_options->get_pointer_to_macro_value(macro.first));
}

float GPUKernel::compute_execution_time()
{
    if (!m_launched_at_least_once)
        return 0.0f;
    float ms = 0.0f;
    if (b2pt_kernel_last_ms(handle_of(m_kernel_function), &ms) != B2PT_OK)
        report("timing failed", m_kernel_function_name);
    m_last_execution_time = ms;
    return ms;
}

float GPUKernel::get_last_execution_time() { return m_last_execution_time; }
bool GPUKernel::has_been_compiled() const { return m_kernel_function != nullptr; }
bool GPUKernel::is_precompiled() const { return m_is_precompiled_kernel; }
void GPUKernel::set_precompiled(bool precompiled) { m_is_precompiled_kernel = precompiled; }
