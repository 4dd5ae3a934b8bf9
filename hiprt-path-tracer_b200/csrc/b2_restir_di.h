/* b2_restir_di.h -- placeholder until ReSTIR DI lands (filled in by a later milestone of this round). */
#ifndef B2_RESTIR_DI_H
#define B2_RESTIR_DI_H
#include "b2_pathtrace.h"
template <class V>
B2_HD void restir_di_pass_pixel(int pass, const Ctx& ctx, uint32_t x, uint32_t y);
#endif
