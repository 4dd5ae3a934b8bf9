"""Reader / writer for the B2SCENE1 scene container (layout: include/b2pt_scene_blob.h)."""
import re
import struct
from pathlib import Path

import numpy as np

_HDR = Path(__file__).resolve().parent.parent / "include" / "b2pt_scene_blob.h"


def material_fields():
    """Parse the X-macro list in the C header so the field order has a single source of truth."""
    text = _HDR.read_text()
    block = text.split("#define B2MAT_FIELDS(X)")[1].split("enum B2MatField")[0]
    return re.findall(r"X\((\w+)\)", block)


MAT_FIELDS = material_fields()
MAT_INDEX = {n: i for i, n in enumerate(MAT_FIELDS)}

# Defaults of the reference's CPUMaterial (MaterialCPU.h:217-329), by field name.
MAT_DEFAULTS = dict(
    emission_r=0, emission_g=0, emission_b=0, emission_strength=1, global_emissive_factor=1,
    base_color_r=1, base_color_g=1, base_color_b=1, roughness=0.3, oren_nayar_sigma=0.34906585039886591538,
    metallic=0, metallic_F90_falloff_exponent=5, metallic_F82_r=1, metallic_F82_g=1, metallic_F82_b=1,
    metallic_F90_r=1, metallic_F90_g=1, metallic_F90_b=1, anisotropy=0, anisotropy_rotation=0,
    second_roughness_weight=0, second_roughness=0.5, do_metallic_energy_compensation=1,
    specular=1, specular_tint=1, specular_color_r=1, specular_color_g=1, specular_color_b=1,
    specular_darkening=0, do_specular_energy_compensation=1,
    coat=0, coat_medium_absorption_r=1, coat_medium_absorption_g=1, coat_medium_absorption_b=1,
    coat_medium_thickness=5, coat_roughness=0, coat_roughening=1, coat_darkening=1, coat_anisotropy=0,
    coat_anisotropy_rotation=0, coat_ior=1.5, do_coat_energy_compensation=1,
    sheen=0, sheen_roughness=0.5, sheen_color_r=1, sheen_color_g=1, sheen_color_b=1,
    ior=1.4, specular_transmission=0, diffuse_transmission=0, absorption_at_distance=1,
    absorption_color_r=1, absorption_color_g=1, absorption_color_b=1,
    dispersion_scale=0, dispersion_abbe_number=20, thin_walled=0, do_glass_energy_compensation=1,
    thin_film=0, thin_film_ior=1.3, thin_film_thickness=500, thin_film_kappa_3=0,
    thin_film_hue_shift_degrees=0, thin_film_base_ior_override=1, thin_film_do_ior_override=0,
    alpha_opacity=1, dielectric_priority=0,
)
assert set(MAT_DEFAULTS) == set(MAT_FIELDS), set(MAT_FIELDS) ^ set(MAT_DEFAULTS)


def default_material(**overrides):
    rec = np.array([MAT_DEFAULTS[n] for n in MAT_FIELDS], dtype=np.float32)
    for k, v in overrides.items():
        if k in ("emission", "base_color", "metallic_F82", "metallic_F90", "specular_color",
                 "coat_medium_absorption", "sheen_color", "absorption_color"):
            for c, x in zip("rgb", v):
                rec[MAT_INDEX[f"{k}_{c}"]] = x
        else:
            rec[MAT_INDEX[k]] = v
    return rec


def camera_block(inverse_view, inverse_projection, view_projection, position, vertical_fov):
    return (np.asarray(inverse_view, np.float32).reshape(16).tobytes()
            + np.asarray(inverse_projection, np.float32).reshape(16).tobytes()
            + np.asarray(view_projection, np.float32).reshape(16).tobytes()
            + np.asarray(position, np.float32).reshape(3).tobytes()
            + struct.pack("<f", vertical_fov))


TEX_SLOTS = ["normal_map", "emission", "base_color", "roughness_metallic", "roughness", "metallic", "anisotropic", "specular", "coat", "sheen", "specular_transmission"]
TEX_NONE, TEX_CONSTANT_EMISSIVE = 65535, 65534


def write_scene(path, camera_bytes, positions, normals, texcoords, indices, material_indices, has_normals, materials, textures=None, texture_indices=None,
                base_color_texture_opaque=None):
    """textures: list of uint8 arrays (height, width, channels), row 0 first; texture_indices: (n_materials, 11) int32 (TEX_SLOTS order)."""
    positions = np.ascontiguousarray(positions, np.float32).reshape(-1, 3)
    normals = np.ascontiguousarray(normals, np.float32).reshape(-1, 3)
    texcoords = np.ascontiguousarray(texcoords, np.float32).reshape(-1, 2)
    indices = np.ascontiguousarray(indices, np.int32).reshape(-1, 3)
    material_indices = np.ascontiguousarray(material_indices, np.int32).reshape(-1)
    has_normals = np.ascontiguousarray(has_normals, np.uint8).reshape(-1)
    materials = np.ascontiguousarray(materials, np.float32).reshape(-1, len(MAT_FIELDS))
    nv, nt, nm = len(positions), len(indices), len(materials)
    assert len(normals) == nv and len(texcoords) == nv and len(has_normals) == nv and len(material_indices) == nt
    assert len(camera_bytes) == 52 * 4
    with open(path, "wb") as f:
        f.write(b"B2SCENE1")
        f.write(struct.pack("<4I", nv, nt, nm, len(MAT_FIELDS)))
        f.write(camera_bytes)
        for a in (positions, normals, texcoords, indices, material_indices, has_normals):
            f.write(a.tobytes())
        f.write(b"\0" * ((4 - (nv & 3)) & 3))
        f.write(materials.tobytes())
        if textures:
            texture_indices = np.ascontiguousarray(texture_indices, np.int32).reshape(nm, len(TEX_SLOTS))
            if base_color_texture_opaque is None:
                base_color_texture_opaque = np.ones(nm, np.uint8)
                for m in range(nm):
                    idx = texture_indices[m, TEX_SLOTS.index("base_color")]
                    if idx < len(textures) and textures[idx].shape[2] == 4:
                        base_color_texture_opaque[m] = 1 if (textures[idx][..., 3] == 255).all() else 0   # Image8Bit::is_fully_opaque
            f.write(b"B2TEXTR1")
            f.write(struct.pack("<I", len(textures)))
            for tex in textures:
                tex = np.ascontiguousarray(tex, np.uint8)
                hgt, wid, ch = tex.shape
                f.write(struct.pack("<3I", wid, hgt, ch))
                f.write(tex.tobytes())
                f.write(b"\0" * ((4 - (tex.size & 3)) & 3))
            f.write(texture_indices.tobytes())
            f.write(np.ascontiguousarray(base_color_texture_opaque, np.uint8).tobytes())


def read_scene(path):
    buf = Path(path).read_bytes()
    assert buf[:8] == b"B2SCENE1"
    nv, nt, nm, nf = struct.unpack_from("<4I", buf, 8)
    assert nf == len(MAT_FIELDS)
    off = 24
    cam = np.frombuffer(buf, np.float32, 52, off)
    off += 52 * 4

    def take(dtype, count, shape):
        nonlocal off
        a = np.frombuffer(buf, dtype, count, off).reshape(shape)
        off += a.nbytes
        return a

    out = dict(
        camera=dict(inverse_view=cam[0:16].reshape(4, 4), inverse_projection=cam[16:32].reshape(4, 4),
                    view_projection=cam[32:48].reshape(4, 4), position=cam[48:51], vertical_fov=float(cam[51])),
        positions=take(np.float32, 3 * nv, (nv, 3)),
        normals=take(np.float32, 3 * nv, (nv, 3)),
        texcoords=take(np.float32, 2 * nv, (nv, 2)),
        indices=take(np.int32, 3 * nt, (nt, 3)),
        material_indices=take(np.int32, nt, (nt,)),
        has_normals=take(np.uint8, nv, (nv,)),
    )
    off += (4 - (nv & 3)) & 3
    out["materials"] = take(np.float32, nm * nf, (nm, nf))
    if buf[off:off + 8] == b"B2TEXTR1":
        off += 8
        (n_tex,) = struct.unpack_from("<I", buf, off)
        off += 4
        textures = []
        for _ in range(n_tex):
            wid, hgt, ch = struct.unpack_from("<3I", buf, off)
            off += 12
            textures.append(take(np.uint8, wid * hgt * ch, (hgt, wid, ch)))
            off += (4 - ((wid * hgt * ch) & 3)) & 3
        out["textures"] = textures
        out["texture_indices"] = take(np.int32, nm * len(TEX_SLOTS), (nm, len(TEX_SLOTS)))
        out["base_color_texture_opaque"] = take(np.uint8, nm, (nm,))
    return out


# ----------------------------------------------------------------------------------------------
# Camera matrices. Restates what the reference host does with glm in Camera::to_hiprt
# (/root/reference/src/Scene/Camera.cpp:9-61) and SceneParser::parse_camera (SceneParser.cpp:227-283),
# including the quirk that the *projection* matrices are reinterpreted from glm's column-major memory
# without a transpose while the view matrix is transposed first.
# ----------------------------------------------------------------------------------------------
def quat_to_mat3(q):
    x, y, z, w = [np.float64(v) for v in q]
    return np.array([
        [1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w)],
        [2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w)],
        [2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)]])


def perspective_glm(fovy, aspect, z_near, z_far):
    """Math-convention (row, col) matrix of glm::perspective (RH, depth -1..1)."""
    t = np.tan(np.float32(fovy) / np.float32(2))
    m = np.zeros((4, 4), np.float64)
    m[0, 0] = 1.0 / (aspect * t)
    m[1, 1] = 1.0 / t
    m[2, 2] = -(z_far + z_near) / (z_far - z_near)
    m[3, 2] = -1.0
    m[2, 3] = -(2.0 * z_far * z_near) / (z_far - z_near)
    return m


def make_camera(translation, rotation_xyzw, horizontal_fov_assimp, aspect, z_near=0.1, z_far=100.0, clip_near=None, clip_far=None):
    """translation / rotation: camera-to-world; returns the 52-float camera block."""
    vertical_fov = np.float32(2.0) * np.arctan(np.tan(np.float32(horizontal_fov_assimp) * np.float32(0.5)) * np.float32(aspect)) + np.float32(0.425)
    # parse_camera builds the projection with the *file's* clip planes (SceneParser.cpp:254)
    cn = z_near if clip_near is None else clip_near
    cf = z_far if clip_far is None else clip_far
    P = perspective_glm(vertical_fov, aspect, cn, cf)
    R = quat_to_mat3(rotation_xyzw)
    V = np.eye(4)
    V[:3, :3] = R.T
    V[:3, 3] = -R.T @ np.asarray(translation, np.float64)
    inverse_view = np.linalg.inv(V)                  # row-major, right-multiplied by column points
    a, b, c, d = P[0, 0], P[1, 1], P[2, 2], P[2, 3]
    Pinv = np.array([[1 / a, 0, 0, 0], [0, 1 / b, 0, 0], [0, 0, 0, -1], [0, 0, 1 / d, c / d]])
    inverse_projection = Pinv.T                      # glm column-major memory read as row-major
    view_projection = P.T @ V                        # (V^T * P) in glm memory read as row-major
    return camera_block(inverse_view, inverse_projection, view_projection, translation, float(vertical_fov))
