/*
 * oracle/ref_stubs.cpp -- TEST INFRASTRUCTURE ONLY.
 * The reference's Utils.cpp cannot be compiled here (it includes OpenImageDenoise, which the reference's CMake
 * downloads). The oracle needs exactly two of its symbols; they are provided here:
 *   Utils::debugbreak           -- no-op (reference: raises SIGTRAP for the debugger)
 *   Utils::compute_alias_table  -- include/b2pt_alias_table.h, a restatement of Utils.cpp:63-131
 * (the global g_imgui_logger object is defined by the reference.s ImGuiLogger.cpp itself).
 */
#include "UI/ImGui/ImGuiLogger.h"
#include "Utils/Utils.h"

#include "b2pt_alias_table.h"


void Utils::debugbreak() {}

void Utils::compute_alias_table(const std::vector<float>& input, float in_input_total_sum, std::vector<float>& out_probas, std::vector<int>& out_alias)
{
    b2pt_build_alias_table(input, in_input_total_sum, out_probas, out_alias);
}
