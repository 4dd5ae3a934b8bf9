/* b2_principled.h -- placeholder until the principled BSDF lands (filled in by a later milestone of this round). */
#ifndef B2_PRINCIPLED_H
#define B2_PRINCIPLED_H
#include "b2_bsdf.h"
#endif
