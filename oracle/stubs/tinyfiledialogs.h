/* oracle/stubs -- TEST INFRASTRUCTURE ONLY: stand-in for libtinyfiledialogs (file picker), never called by the oracle. */
#pragma once
inline const char* tinyfd_openFileDialog(const char*, const char*, int, const char* const*, const char*, int) { return nullptr; }
