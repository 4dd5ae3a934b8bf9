/*
 * div_selftest.cu -- b2_div3_exact (b2_math.h: three divisions by one divisor through one refined reciprocal) against the
 * compiler's IEEE division, bit for bit, on the GPU. Built by build.py with the precise flavour's flags; run by
 * tests/test_gpu_parity.py (pytest -m gpu):  tools/div_selftest [pairs_per_thread]
 * Operand families: raw random bit patterns (every exponent, denormals, inf, NaN), shading-range values, exact and signed
 * zeros, mantissas of all ones / one bit, quotients next to rounding boundaries (c = small odd integers).
 */
#define B2_FLAVOUR_PRECISE
#include "b2_math.h"

#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

__device__ unsigned long long g_mismatch[8];
__device__ unsigned long long g_fast[8];
__device__ float g_example[8][4];

__device__ __forceinline__ bool same_bits(float a, float b)
{
    if (isnan(a) && isnan(b))
        return true;
    return __float_as_uint(a) == __float_as_uint(b);
}

__device__ float make_operand(int family, Rng& rng, bool divisor)
{
    const uint32_t u = rng.next_u32();
    switch (family)
    {
    case 0: return __uint_as_float(u);                                                    /* anything */
    case 1: return (divisor ? 1.0e-3f : -2.0f) + (float)(u >> 8) * (1.0f / 16777216.0f) * (divisor ? 1.0e3f : 4.0f); /* shading range */
    case 2: return (u & 3u) == 0u ? 0.0f : (u & 3u) == 1u ? -0.0f : __uint_as_float((u & 0x807FFFFFu) | 0x3F800000u); /* zeros, [1, 2) */
    case 3: return __uint_as_float((u & 0xFF800000u) | ((u & 1u) ? 0x007FFFFFu : (1u << ((u >> 1) % 23u))));          /* edge mantissas */
    case 4: return divisor ? (float)(2u * (u % 64u) + 1u) : (float)(u % (1u << 24));      /* integer quotients near ties */
    case 5: return __uint_as_float((u & 0x807FFFFFu) | ((60u + (u >> 24) % 135u) << 23)); /* wide exponent range */
    case 6: return __uint_as_float((u & 0x807FFFFFu) | ((divisor ? 100u + (u >> 24) % 56u : (u >> 24) % 255u) << 23)); /* tiny / huge numerators */
    default: return __uint_as_float(u & 0x807FFFFFu);                                     /* denormals */
    }
}

__global__ void check(int pairs_per_thread, unsigned int seed)
{
    Rng rng = make_rng(wang_hash(seed + blockIdx.x * blockDim.x + threadIdx.x) | 1u);
    for (int i = 0; i < pairs_per_thread; i++)
    {
        const int family = i & 7;
        const float ax = make_operand(family, rng, false), ay = make_operand(family, rng, false), az = make_operand((family + 2) & 7, rng, false);
        const float c = make_operand(family, rng, true);
        const V3 q = b2_div3_exact(ax, ay, az, c);
        const float ex = __fdiv_rn(ax, c), ey = __fdiv_rn(ay, c), ez = __fdiv_rn(az, c);
        if (!same_bits(q.x, ex) || !same_bits(q.y, ey) || !same_bits(q.z, ez))
        {
            if (atomicAdd(&g_mismatch[family], 1ull) == 0ull)
            {
                const bool bx = !same_bits(q.x, ex), by = !same_bits(q.y, ey);
                g_example[family][0] = bx ? ax : by ? ay : az;
                g_example[family][1] = c;
                g_example[family][2] = bx ? q.x : by ? q.y : q.z;
                g_example[family][3] = bx ? ex : by ? ey : ez;
            }
        }
    }
}

int main(int argc, char** argv)
{
    const int pairs = argc > 1 ? std::atoi(argv[1]) : 4096;
    const int blocks = 148 * 8, threads = 256;
    check<<<blocks, threads>>>(pairs, 0xB200u);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess)
    {
        std::printf("{\"error\": \"%s\"}\n", cudaGetErrorString(e));
        return 2;
    }
    unsigned long long mismatch[8];
    float example[8][4];
    cudaMemcpyFromSymbol(mismatch, g_mismatch, sizeof(mismatch));
    cudaMemcpyFromSymbol(example, g_example, sizeof(example));
    unsigned long long total = 0;
    for (int f = 0; f < 8; f++)
        total += mismatch[f];
    std::printf("{\"triples\": %llu, \"mismatches\": %llu, \"per_family\": [", (unsigned long long)blocks * threads * pairs, total);
    for (int f = 0; f < 8; f++)
        std::printf("%llu%s", mismatch[f], f < 7 ? ", " : "]");
    for (int f = 0; f < 8; f++)
        if (mismatch[f])
            std::printf(", \"example_%d\": [%.9g, %.9g, %.9g, %.9g]", f, example[f][0], example[f][1], example[f][2], example[f][3]);
    std::printf("}\n");
    return total == 0 ? 0 : 1;
}
