/* b2_restir_di_shading.h -- placeholder until ReSTIR DI lands (filled in by a later milestone of this round). */
#ifndef B2_RESTIR_DI_SHADING_H
#define B2_RESTIR_DI_SHADING_H
#include "b2_envmap.h"
template <class V>
B2_HD Rgb sample_light_ReSTIR_DI(const Ctx& ctx, Payload& payload, const HitInfo& hit, V3 view_direction, Rng& rng, int x, int y);
#endif
