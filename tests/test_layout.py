"""The byte offsets the Python binding pokes settings at (b2pt.SETTINGS_OFFSETS) against offsetof() on the C struct
(include/b2pt_render_data.h, itself checked member by member against the reference's HIPRTRenderData by oracle/layout_check.cpp)."""
import subprocess

import pytest

from conftest import ROOT

MEMBERS = {
    "nb_bounces": "render_settings.nb_bounces",
    "samples_per_frame": "render_settings.samples_per_frame",
    "sample_number": "render_settings.sample_number",
    "need_to_reset": "render_settings.need_to_reset",
    "do_russian_roulette": "render_settings.do_russian_roulette",
    "russian_roulette_min_depth": "render_settings.russian_roulette_min_depth",
    "freeze_random": "render_settings.freeze_random",
    "accumulate": "render_settings.accumulate",
    "enable_direct": "render_settings.enable_direct",
    "ris_number_of_light_candidates": "render_settings.ris_settings.number_of_light_candidates",
    "ris_number_of_bsdf_candidates": "render_settings.ris_settings.number_of_bsdf_candidates",
    "restir_di_do_temporal_reuse_pass": "render_settings.restir_di_settings.common_temporal_pass.do_temporal_reuse_pass",
    "restir_di_do_spatial_reuse_pass": "render_settings.restir_di_settings.common_spatial_pass.do_spatial_reuse_pass",
    "restir_di_number_of_passes": "render_settings.restir_di_settings.common_spatial_pass.number_of_passes",
    "restir_di_reuse_radius": "render_settings.restir_di_settings.common_spatial_pass.reuse_radius",
    "restir_di_reuse_neighbor_count": "render_settings.restir_di_settings.common_spatial_pass.reuse_neighbor_count",
    "restir_di_disocclusion_reuse_count": "render_settings.restir_di_settings.common_spatial_pass.disocclusion_reuse_count",
    "restir_di_use_adaptive_directional_spatial_reuse": "render_settings.restir_di_settings.common_spatial_pass.use_adaptive_directional_spatial_reuse",
    "restir_di_m_cap": "render_settings.restir_di_settings.m_cap",
    "restir_di_use_confidence_weights": "render_settings.restir_di_settings.use_confidence_weights",
    "restir_di_do_fused_spatiotemporal": "render_settings.restir_di_settings.do_fused_spatiotemporal",
    "current_camera_do_jittering": "current_camera.do_jittering",
    "ambient_light_type": "world_settings.ambient_light_type",
    "envmap_intensity": "world_settings.envmap_intensity",
    "white_furnace_mode": "bsdfs_data.white_furnace_mode",
}
SIZES = {"i": 4, "I": 4, "B": 1, "f": 4}


def test_settings_offsets_match_the_c_struct(b2pt, tmp_path):
    assert set(MEMBERS) == set(b2pt.SETTINGS_OFFSETS), "every setting the binding exposes is checked here"
    names = sorted(MEMBERS)
    lines = "\n".join(f'    std::printf("{n} %zu %zu\\n", offsetof(b2pt_render_data, {MEMBERS[n]}), sizeof(((b2pt_render_data*)0)->{MEMBERS[n]}));' for n in names)
    src = tmp_path / "offsets.cpp"
    src.write_text(f'#include <cstdio>\n#include <cstddef>\n#include "b2pt_render_data.h"\nint main()\n{{\n{lines}\n    return 0;\n}}\n')
    exe = tmp_path / "offsets"
    subprocess.run(["g++", "-w", f"-I{ROOT / 'include'}", str(src), "-o", str(exe)], check=True)
    out = subprocess.run([str(exe)], check=True, capture_output=True, text=True).stdout
    for line in out.splitlines():
        name, offset, size = line.split()
        expected_offset, kind = b2pt.SETTINGS_OFFSETS[name]
        assert int(offset) == expected_offset, f"{name}: the binding writes at {expected_offset}, the struct has it at {offset}"
        assert int(size) == SIZES[kind], f"{name}: the binding writes a '{kind}', the member is {size} bytes"
