/* oracle/stubs -- TEST INFRASTRUCTURE ONLY. Minimal stand-in for OpenImageDenoise (which the reference's CMake downloads) so that
 * the reference's own src/Utils/Utils.cpp compiles where it lies: the oracle uses Utils::compute_alias_table and the FLIP /
 * RMSE helpers of that file; the denoiser entry point (Utils::oidn_denoise) is never called and does nothing here. */
#pragma once
#include <cstddef>
namespace oidn
{
enum class DeviceType { Default, CPU };
enum class Format { Float3 };
enum class Error { None };
struct BufferRef { void* getData() { return nullptr; } };
struct FilterRef
{
    void setImage(const char*, BufferRef, Format, int, int) {}
    void set(const char*, bool) {}
    void commit() {}
    void execute() {}
};
struct DeviceRef
{
    void* getHandle() { return nullptr; }
    void commit() {}
    BufferRef newBuffer(size_t) { return BufferRef(); }
    FilterRef newFilter(const char*) { return FilterRef(); }
    Error getError(const char*&) { return Error::None; }
};
inline DeviceRef newDevice(int = 0) { return DeviceRef(); }
} // namespace oidn
inline int oidnGetNumPhysicalDevices() { return 0; }
inline int oidnGetPhysicalDeviceInt(int, const char*) { return 0; }
