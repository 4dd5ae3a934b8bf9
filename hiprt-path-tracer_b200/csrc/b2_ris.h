/* b2_ris.h -- placeholder until the RIS strategy lands (filled in by a later milestone of this round). */
#ifndef B2_RIS_H
#define B2_RIS_H
#include "b2_envmap.h"
template <class V>
B2_HD Rgb sample_lights_RIS(const Ctx& ctx, Payload& payload, const HitInfo& hit, V3 view_direction, Rng& rng);
#endif
