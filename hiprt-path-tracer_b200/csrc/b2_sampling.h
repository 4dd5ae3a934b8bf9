/*
 * b2_sampling.h -- small sampling / frame utilities shared by the BSDFs and the light samplers.
 * Restates /root/reference/src/Device/includes/Sampling.h:50-57,59-101,103-146 and ONB.h:18-77 with the same
 * operation order (the sampled directions feed every later intersection, so they must round identically).
 */
#ifndef B2_SAMPLING_H
#define B2_SAMPLING_H

#include "b2_math.h"

B2_HD float balance_heuristic(float pdf_a, float nb_a, float pdf_b, float nb_b)
{
    if (pdf_a == 0.0f)
        return 0.0f;
    return pdf_a / (nb_a * pdf_a + nb_b * pdf_b);
}
B2_HD float balance_heuristic(float pdf_a, float pdf_b) { return balance_heuristic(pdf_a, 1.0f, pdf_b, 1.0f); }

B2_HD V2 sample_in_disk_uv(float radius, V2 uv)
{
    float r_sqrt_v = radius * sqrtf(uv.y);
    b2w sin_w, cos_w;
    b2w_sincos(B2_TWO_PI * uv.x, sin_w, cos_w);
    float x = (float)((b2w)r_sqrt_v * cos_w);
    float y = (float)((b2w)r_sqrt_v * sin_w);
    return v2(x, y);
}

B2_HD V3 refract_ray(V3 ray_direction, V3 surface_normal, float relative_eta)
{
    float NoI = dot(ray_direction, surface_normal);
    float sin_theta_i_2 = 1.0f - NoI * NoI;
    float root_term = 1.0f - sin_theta_i_2 / (relative_eta * relative_eta);
    float cos_theta_t = sqrtf(root_term);
    return -ray_direction / relative_eta + (NoI / relative_eta - cos_theta_t) * surface_normal;
}

B2_HD V3 cosine_weighted_sample_around_normal_world_space(V3 normal, Rng& rng)
{
    float rand_1 = rng.next_float();
    float rand_2 = 2.0f * rng.next_float() - 1.0f;
    if (rand_1 < 1.0e-8f && rand_2 < -0.999999f && normal.z > 0.999999f)
    {
        rand_1 += 1.0e-7f;
        rand_2 += 1.0e-7f;
    }
    float theta = B2_TWO_PI * rand_1;
    float s = sqrtf(1.0f - rand_2 * rand_2);
    b2w sin_w, cos_w;
    b2w_sincos(theta, sin_w, cos_w);
    V3 sphere_point = v3(s * (float)cos_w, s * (float)sin_w, rand_2);
    return normalize(normal + sphere_point);
}

B2_HD V3 cosine_weighted_sample_z_up_frame(Rng& rng)
{
    float r1 = rng.next_float();
    float r2 = rng.next_float();
    float phi = B2_TWO_PI * r1;
    float cos_theta = sqrtf(r2);
    float sin_theta = sqrtf(1 - cos_theta * cos_theta);
    b2w sin_w, cos_w;
    b2w_sincos(phi, sin_w, cos_w);
    return normalize(v3((float)(cos_w * (b2w)sin_theta), (float)(sin_w * (b2w)sin_theta), cos_theta));
}

/* Duff et al. style branchless basis, ONB.h:18-30 */
B2_HD void build_ONB(V3 N, V3& T, V3& B)
{
    if (N.z < -0.99998796f)
    {
        T = v3(0.0f, -1.0f, 0.0f);
        B = v3(-1.0f, 0.0f, 0.0f);
        return;
    }
    float nxa = -N.x / (1.0f + N.z);
    T = v3(1.0f + N.x * nxa, nxa * N.y, -N.x);
    B = v3(T.y, 1.0f - N.y * N.y / (1.0f + N.z), -N.y);
}

B2_HD void build_rotated_ONB(V3 N, V3& T, V3& B, float basis_rotation)
{
    V3 up = fabsf(N.z) < 0.9999999f ? v3(0.0f, 0.0f, 1.0f) : v3(1.0f, 0.0f, 0.0f);
    T = normalize(cross(up, N));
    b2w sin_w, cos_w;
    b2w_sincos(basis_rotation, sin_w, cos_w);
    T = T * (float)cos_w + cross(N, T) * (float)sin_w + N * dot(N, T) * (float)((b2w)1.0f - cos_w);
    B = cross(N, T);
}

B2_HD V3 local_to_world_frame(V3 T, V3 B, V3 N, V3 V) { return normalize(V.x * T + V.y * B + V.z * N); }
B2_HD V3 local_to_world_frame(V3 N, V3 V)
{
    V3 T, B;
    build_ONB(N, T, B);
    return normalize(V.x * T + V.y * B + V.z * N);
}
B2_HD V3 world_to_local_frame(V3 T, V3 B, V3 N, V3 V) { return normalize(v3(dot(V, T), dot(V, B), dot(V, N))); }
B2_HD V3 world_to_local_frame(V3 N, V3 V)
{
    V3 T, B;
    build_ONB(N, T, B);
    return normalize(v3(dot(V, T), dot(V, B), dot(V, N)));
}

#endif
