#!/usr/bin/env python3
"""data/cornell_dispersion_16x9.b2scene: the Cornell box of data/cornell_pbr_16x9.b2scene with spectral dispersion on its glass sphere
(dispersion_scale 1, Abbe number 20, a little roughness so that both the delta and the microfacet glass paths run) and on a second,
rough transmissive material -- exercises sample_wavelength_uniformly / compute_dispersion_ior / get_dispersion_ray_color
(src/Device/includes/Dispersion.h:469-522, Principled.h:377-380,578-581,731-734, PathTracing.h:52). Deterministic.

    python tools/make_dispersion_scene.py data/cornell_pbr_16x9.b2scene data/cornell_dispersion_16x9.b2scene
"""
import sys

import numpy as np

import b2scene


def main(src, dst):
    s = b2scene.read_scene(src)
    mats = np.array(s["materials"], np.float32).copy()
    M = b2scene.MAT_INDEX
    mats[8, M["dispersion_scale"]] = 1.0
    mats[8, M["dispersion_abbe_number"]] = 20.0
    mats[7, M["specular_transmission"]] = 0.8
    mats[7, M["roughness"]] = 0.25
    mats[7, M["dispersion_scale"]] = 0.6
    mats[7, M["dispersion_abbe_number"]] = 35.0
    cam = s["camera"]
    camera = b2scene.camera_block(cam["inverse_view"], cam["inverse_projection"], cam["view_projection"], cam["position"], cam["vertical_fov"])
    b2scene.write_scene(dst, camera, s["positions"], s["normals"], s["texcoords"], s["indices"], s["material_indices"], s["has_normals"], mats)
    print(f"wrote {dst}")


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2])
