/*
 * oracle/dump_luts.cpp -- TEST/BUILD TOOL. Reads the reference's baked directional-albedo tables
 * (/root/reference/data/BRDFsData/**, Radiance .hdr) with the reference's OWN loader
 * (Image32Bit::read_image_hdr(path, 1, flipY=true), as CPURenderer::setup_brdfs_data does,
 * /root/reference/src/Renderer/CPURenderer.cpp:155-194) and writes the decoded floats into one container:
 *
 *   "B2LUTS01", uint32 count, then per table: char name[32], int32 w, h, d, float data[d][h][w]
 *
 * The container is DATA the hot path reads (the reference ships these LUTs as data files too); it is consumed
 * by the product (uploaded as linear fp32 arrays) and by the oracle harness, so both see identical floats.
 * File names follow GPUBakerConstants (Renderer/Baker/GPUBakerConstants.h:43-100) for the default
 * HeightUncorrelated masking-shadowing flavour (HostDeviceCommon/BSDFsData.h:61).
 */
#include "Device/includes/BSDFs/SheenLTCFittedParameters.h"
#include "Image/Image.h"
#include "Renderer/Baker/GPUBakerConstants.h"

#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

static void write_table(FILE* f, const char* name, const std::vector<Image32Bit>& slices)
{
    char padded[32] = { 0 };
    std::strncpy(padded, name, 31);
    int32_t dims[3] = { slices[0].width, slices[0].height, (int32_t)slices.size() };
    std::fwrite(padded, 1, 32, f);
    std::fwrite(dims, 4, 3, f);
    for (const Image32Bit& s : slices)
    {
        if (s.width != dims[0] || s.height != dims[1] || s.channels != 1) { std::fprintf(stderr, "bad slice in %s\n", name); std::exit(1); }
        std::fwrite(s.data().data(), 4, s.data().size(), f);
    }
    std::printf("%s: %dx%dx%d\n", name, dims[0], dims[1], dims[2]);
}

int main(int argc, char** argv)
{
    if (argc != 3) { std::fprintf(stderr, "usage: %s <reference data/BRDFsData dir> <out.b2luts>\n", argv[0]); return 2; }
    std::string root = argv[1];
    GGXMaskingShadowingFlavor flavor = GGXMaskingShadowingFlavor::HeightUncorrelated;

    FILE* f = std::fopen(argv[2], "wb");
    if (!f) { std::fprintf(stderr, "cannot write %s\n", argv[2]); return 1; }
    uint32_t count = 6;
    std::fwrite("B2LUTS01", 1, 8, f);
    std::fwrite(&count, 4, 1, f);

    auto load_stack = [&](const std::string& dir, const std::string& suffix, int n) {
        std::vector<Image32Bit> out(n);
        for (int i = 0; i < n; i++)
        {
            out[i] = Image32Bit::read_image_hdr(root + dir + std::to_string(i) + suffix, 1, true);
            if (out[i].width == 0) { std::fprintf(stderr, "missing %s%d%s\n", dir.c_str(), i, suffix.c_str()); std::exit(1); }
        }
        return out;
    };

    write_table(f, "ggx_conductor", { Image32Bit::read_image_hdr(root + "/GGX/" + GPUBakerConstants::get_GGX_conductor_directional_albedo_texture_filename(flavor), 1, true) });
    write_table(f, "glossy_dielectric", load_stack("/GlossyDielectrics/", GPUBakerConstants::get_glossy_dielectric_directional_albedo_texture_filename(flavor), GPUBakerConstants::GLOSSY_DIELECTRIC_TEXTURE_SIZE_IOR));
    write_table(f, "ggx_glass", load_stack("/GGX/Glass/", GPUBakerConstants::get_GGX_glass_directional_albedo_texture_filename(flavor), GPUBakerConstants::GGX_GLASS_DIRECTIONAL_ALBEDO_TEXTURE_SIZE_IOR));
    write_table(f, "ggx_glass_inv", load_stack("/GGX/Glass/", GPUBakerConstants::get_GGX_glass_directional_albedo_inv_texture_filename(flavor), GPUBakerConstants::GGX_GLASS_DIRECTIONAL_ALBEDO_TEXTURE_SIZE_IOR));
    write_table(f, "ggx_thin_glass", load_stack("/GGX/Glass/", GPUBakerConstants::get_GGX_thin_glass_directional_albedo_texture_filename(flavor), GPUBakerConstants::GGX_THIN_GLASS_DIRECTIONAL_ALBEDO_TEXTURE_SIZE_IOR));
    {
        /* sheen LTC fit (Ai, Bi, Ri), 32x32 texels of 3 floats, compiled into the reference
         * (Device/includes/BSDFs/SheenLTCFittedParameters.h; CPURenderer.cpp:157): stored as a 96x32x1 table whose
         * rows are 32 RGB texels */
        char padded[32] = { 0 };
        std::strncpy(padded, "sheen_ltc", 31);
        int32_t dims[3] = { 96, 32, 1 };
        std::fwrite(padded, 1, 32, f);
        std::fwrite(dims, 4, 3, f);
        std::fwrite(reinterpret_cast<const float*>(ltc_parameters_table_approximation.data()), 4, 32 * 32 * 3, f);
        std::printf("sheen_ltc: 32x32 rgb\n");
    }
    std::fclose(f);
    return 0;
}
