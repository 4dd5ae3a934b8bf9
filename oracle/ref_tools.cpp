/*
 * oracle/ref_tools.cpp -- TEST INFRASTRUCTURE ONLY. Small commands over the reference's REAL src/Utils/Utils.cpp (compiled from
 * where it lies with the stand-in headers of oracle/stubs/ for OpenImageDenoise / clip / tinyfiledialogs):
 *   ref_tools alias-check N SEED      random weight lists through Utils::compute_alias_table (Utils.cpp:63-131) and through
 *                                     include/b2pt_alias_table.h (what the product and ref_render_* use): must agree bit for bit
 *   ref_tools alias LIST.bin OUT      alias table of a float32 list by the reference's function -> OUT_probas.bin / OUT_alias.bin
 *   ref_tools read-hdr FILE.hdr OUT   the reference's .hdr loader (Image32Bit::read_image_hdr(path, 4, flipY = true)) -> raw fp32 RGBA
 *   ref_tools flip A.bin B.bin W H N  A, B = fp32 RGB accumulation buffers of N samples: the reference's tone mapper
 *                                     (Utils::tonemap_hdr_image, gamma 2.2, exposure 1) then its FLIP (thirdparties/nvidia-FLIP
 *                                     through Utils::compute_image_weighted_median_FLIP) and RMSE (Utils::compute_image_rmse)
 */
#include "Image/Image.h"
#include "UI/ImGui/ImGuiLogger.h"
#include "Utils/Utils.h"

#include "b2pt_alias_table.h"

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <random>
#include <string>
#include <vector>

static std::vector<float> read_floats(const char* path)
{
    FILE* f = std::fopen(path, "rb");
    if (!f) { std::fprintf(stderr, "cannot read %s\n", path); std::exit(1); }
    std::fseek(f, 0, SEEK_END);
    long bytes = std::ftell(f);
    std::fseek(f, 0, SEEK_SET);
    std::vector<float> v((size_t)bytes / 4);
    if (std::fread(v.data(), 4, v.size(), f) != v.size()) { std::fprintf(stderr, "short read %s\n", path); std::exit(1); }
    std::fclose(f);
    return v;
}

static int alias_check(int n_lists, unsigned seed)
{
    std::mt19937 gen(seed);
    long long mismatches = 0, entries = 0;
    for (int l = 0; l < n_lists; l++)
    {
        const int family = l % 6;
        const size_t n = family == 5 ? 1 + gen() % 4 : 1 + gen() % 5000;
        std::vector<float> w(n);
        std::uniform_real_distribution<float> uni(0.0f, 1.0f);
        for (size_t i = 0; i < n; i++)
        {
            float u = uni(gen);
            switch (family)
            {
            case 0: w[i] = u; break;
            case 1: w[i] = std::exp(20.0f * u - 10.0f); break;           /* nine decades, like emissive power */
            case 2: w[i] = (gen() % 4 == 0) ? 0.0f : u; break;           /* zeros */
            case 3: w[i] = 1.0f; break;                                   /* all equal */
            case 4: w[i] = (i == 0) ? 1.0e6f : 1.0e-3f * u; break;        /* one dominant entry */
            default: w[i] = u; break;
            }
        }
        float sum = 0.0f;
        for (float x : w)
            sum += x;
        std::vector<float> p_ref, p_mine;
        std::vector<int> a_ref, a_mine;
        Utils::compute_alias_table(w, sum, p_ref, a_ref);
        b2pt_build_alias_table(w, sum, p_mine, a_mine);
        if (p_ref.size() != p_mine.size() || a_ref.size() != a_mine.size())
        {
            mismatches += (long long)n;
            continue;
        }
        for (size_t i = 0; i < p_ref.size(); i++)
        {
            entries++;
            if (std::memcmp(&p_ref[i], &p_mine[i], 4) != 0 || a_ref[i] != a_mine[i])
                mismatches++;
        }
    }
    std::printf("{\"tool\": \"alias-check\", \"lists\": %d, \"entries\": %lld, \"mismatches\": %lld}\n", n_lists, entries, mismatches);
    return mismatches == 0 ? 0 : 1;
}

int main(int argc, char** argv)
{
    if (argc >= 4 && !std::strcmp(argv[1], "alias-check"))
        return alias_check(std::atoi(argv[2]), (unsigned)std::strtoul(argv[3], nullptr, 10));
    if (argc >= 4 && !std::strcmp(argv[1], "alias"))
    {
        std::vector<float> w = read_floats(argv[2]);
        float sum = 0.0f;
        for (float x : w)
            sum += x;
        std::vector<float> probas;
        std::vector<int> alias;
        Utils::compute_alias_table(w, sum, probas, alias);
        std::string out = argv[3];
        FILE* f = std::fopen((out + "_probas.bin").c_str(), "wb");
        std::fwrite(probas.data(), 4, probas.size(), f);
        std::fclose(f);
        f = std::fopen((out + "_alias.bin").c_str(), "wb");
        std::fwrite(alias.data(), 4, alias.size(), f);
        std::fclose(f);
        std::printf("{\"tool\": \"alias\", \"entries\": %zu, \"sum\": %.9g}\n", w.size(), sum);
        return 0;
    }
    if (argc >= 7 && !std::strcmp(argv[1], "flip"))
    {
        const int w = std::atoi(argv[4]), h = std::atoi(argv[5]), samples = std::atoi(argv[6]);
        std::vector<float> a = read_floats(argv[2]), b = read_floats(argv[3]);
        if (a.size() != (size_t)w * h * 3 || b.size() != a.size()) { std::fprintf(stderr, "image size mismatch\n"); return 2; }
        /* HDR statistics on the per-sample averages */
        double se = 0.0, ref_sq = 0.0, mean_a = 0.0, mean_b = 0.0;
        for (size_t i = 0; i < a.size(); i++)
        {
            const double x = a[i] / samples, y = b[i] / samples;
            se += (x - y) * (x - y);
            ref_sq += x * x;
            mean_a += x;
            mean_b += y;
        }
        /* the reference's display transform, then its FLIP + RMSE on the LDR images */
        auto to_ldr = [&](const std::vector<float>& hdr) {
            std::vector<unsigned char> u8 = Utils::tonemap_hdr_image(hdr, samples, 2.2f, 1.0f);
            std::vector<float> f(u8.size());
            for (size_t i = 0; i < u8.size(); i++)
                f[i] = u8[i] / 255.0f;
            return Image32Bit(f, w, h, 3);
        };
        Image32Bit la = to_ldr(a), lb = to_ldr(b);
        float* error_map = nullptr;
        const float flip = Utils::compute_image_weighted_median_FLIP(la, lb, &error_map);
        const float rmse_ldr = Utils::compute_image_rmse(la, lb);
        std::printf("{\"tool\": \"flip\", \"width\": %d, \"height\": %d, \"samples\": %d, \"mean_flip\": %.6f, \"rmse_ldr\": %.6f, \"rmse_hdr\": %.6g, \"rel_rmse_hdr\": %.6g, "
                    "\"mean_a\": %.6g, \"mean_b\": %.6g}\n",
                    w, h, samples, flip, rmse_ldr, std::sqrt(se / a.size()), std::sqrt(se / (ref_sq > 0 ? ref_sq : 1.0)), mean_a / a.size(), mean_b / a.size());
        return 0;
    }
    if (argc >= 4 && !std::strcmp(argv[1], "read-hdr"))
    {
        /* the reference's own loader (stb_image through Image32Bit::read_image_hdr), 4 channels, flipped like RendererEnvmap does */
        Image32Bit image = Image32Bit::read_image_hdr(argv[2], 4, true);
        FILE* f = std::fopen(argv[3], "wb");
        std::fwrite(image.data().data(), 4, image.data().size(), f);
        std::fclose(f);
        std::printf("{\"tool\": \"read-hdr\", \"width\": %d, \"height\": %d, \"channels\": %d}\n", image.width, image.height, image.channels);
        return 0;
    }
    std::fprintf(stderr, "usage: ref_tools alias-check N SEED | alias LIST.bin OUT | flip A.bin B.bin W H SAMPLES\n");
    return 2;
}
