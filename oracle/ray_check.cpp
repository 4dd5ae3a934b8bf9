/*
 * oracle/ray_check.cpp -- TEST INFRASTRUCTURE ONLY. One ray through the reference's CPU acceleration structure (BVH::intersect: the
 * octree of src/Renderer/BVH.cpp with the 7-plane bounding volumes of BoundingVolume.h) and, beside it, through EVERY triangle of the
 * scene with the reference's own Triangle::intersect (src/Renderer/Triangle.h:20-63).
 * Why it exists: the only pixels of a 480x270 light-presampling run where the product and the (zero-initialised) reference disagree go back
 * to ONE shadow ray whose closest hit the reference's octree does not return -- it answers a triangle 2e-4 further along the ray, beyond
 * the ray's t_max, and calls the sample visible; the brute-force loop over the same triangles with the same test finds the nearer one, as
 * the product's traversal does. The reference's traversal has false negatives; the product's has none, which is what
 * tools/traverse_selftest.cu holds it to. tests/test_oracle.py keeps this ray as a witness.
 *
 *   ray_check --scene X.b2scene --luts L.b2luts --ray ox oy oz dx dy dz [--skip-prim N] [--t-max T]
 */
#define main ref_harness_main
#include "ref_harness.cpp"
#undef main

int main(int argc, char** argv)
{
    float r[6] = { 0, 0, 0, 0, 0, 1 };
    int skip_prim = -1;
    float t_max = 3.0e38f;
    std::vector<char*> rest = { argv[0] };
    for (int i = 1; i < argc; i++)
    {
        std::string k = argv[i];
        if (k == "--ray" && i + 6 < argc) { for (int j = 0; j < 6; j++) r[j] = std::strtof(argv[++i], nullptr); }
        else if (k == "--skip-prim" && i + 1 < argc) skip_prim = std::atoi(argv[++i]);
        else if (k == "--t-max" && i + 1 < argc) t_max = std::strtof(argv[++i], nullptr);
        else rest.push_back(argv[i]);
    }
    Args args = parse_args((int)rest.size(), rest.data());
    args.width = args.height = 8;
    Oracle o(args);
    o.rd.render_settings.nb_bounces = 1;
    o.rd.render_settings.samples_per_frame = 1;
    o.setup_brdfs_data(args.luts.c_str());
    o.set_envmap(args.envmap.c_str());
    o.set_scene(args.scene.c_str());

    hiprtRay ray;
    ray.origin = make_float3(r[0], r[1], r[2]);
    ray.direction = make_float3(r[3], r[4], r[5]);
    Xorshift32Generator rng(1);
    hiprtHit octree_hit = intersect_scene_cpu(o.rd, o.rd.cpu_only.bvh, ray, skip_prim, rng);

    int brute_prim = -1, brute_hits = 0;
    float brute_t = 3.0e38f;
    const int n_tri = (int)o.scene.indices.size() / 3;
    for (int i = 0; i < n_tri; i++)
    {
        if (i == skip_prim)
            continue;
        auto vertex = [&](int corner) {
            const float* p = &o.scene.positions[3 * (size_t)o.scene.indices[3 * (size_t)i + corner]];
            return make_float3(p[0], p[1], p[2]);
        };
        Triangle triangle(vertex(0), vertex(1), vertex(2));
        hiprtHit h;
        if (triangle.intersect(ray, h))
        {
            brute_hits++;
            if (h.t < brute_t)
            {
                brute_t = h.t;
                brute_prim = i;
            }
        }
    }
    std::printf("{\"ray_check\": true, \"octree_prim\": %d, \"octree_t\": %.9g, \"brute_force_prim\": %d, \"brute_force_t\": %.9g, \"triangles_hit\": %d, "
                "\"octree_occluded\": %s, \"brute_force_occluded\": %s}\n",
                octree_hit.hasHit() ? (int)octree_hit.primID : -1, octree_hit.hasHit() ? octree_hit.t : -1.0f, brute_prim, brute_prim >= 0 ? brute_t : -1.0f, brute_hits,
                (octree_hit.hasHit() && octree_hit.t < t_max) ? "true" : "false", (brute_prim >= 0 && brute_t < t_max) ? "true" : "false");
    return 0;
}
