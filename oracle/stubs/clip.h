/* oracle/stubs -- TEST INFRASTRUCTURE ONLY: stand-in for thirdparties/clip (clipboard), never called by the oracle. */
#pragma once
namespace clip
{
struct image_spec { unsigned long width, height, bits_per_pixel, bytes_per_row, red_mask, green_mask, blue_mask, alpha_mask, red_shift, green_shift, blue_shift, alpha_shift; };
struct image { image(const void*, const image_spec&) {} };
inline bool set_image(const image&) { return false; }
} // namespace clip
