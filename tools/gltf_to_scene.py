#!/usr/bin/env python3
"""glTF 2.0 -> B2SCENE1 converter for the untextured test scenes (cornell_pbr.gltf).

The reference ingests glTF through ASSIMP (`ReadFile(PreTransformVertices | Triangulate | GenBoundingBoxes)`,
/root/reference/src/Scene/SceneParser.cpp:25) and then flattens aiMeshes into arrays (:98-177) and reads material
keys (:391-440). ASSIMP-Fork needs CMake-generated headers, so it is not built here; this script restates the
part of that pipeline the test scenes exercise:
  * one output mesh per glTF primitive, ordered by material index (PreTransformVertices joins by material);
  * node TRS baked into positions; normals transformed by the inverse-transpose and re-normalised;
  * vertex indices made global by adding the running (max index + 1) offset;
  * material keys as mapped by ASSIMP's glTF2 importer (code/AssetLib/glTF2/glTF2Importer.cpp:254-374) and then
    consumed by `SceneParser::read_material_properties`.
Scene ingest is outside the hot path; oracle and product both consume the blob this
writes, so they see the same scene whatever ASSIMP would have produced.
"""
import argparse
import json
import struct
import sys
from pathlib import Path

import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parent))
import b2scene  # noqa: E402

_COMP = {5120: np.int8, 5121: np.uint8, 5122: np.int16, 5123: np.uint16, 5125: np.uint32, 5126: np.float32}
_NCOMP = {"SCALAR": 1, "VEC2": 2, "VEC3": 3, "VEC4": 4}


def _accessor(g, bins, idx):
    acc = g["accessors"][idx]
    bv = g["bufferViews"][acc["bufferView"]]
    dt = np.dtype(_COMP[acc["componentType"]])
    n = _NCOMP[acc["type"]]
    off = bv.get("byteOffset", 0) + acc.get("byteOffset", 0)
    stride = bv.get("byteStride", 0)
    buf = bins[bv["buffer"]]
    if stride and stride != dt.itemsize * n:
        rows = [np.frombuffer(buf, dt, n, off + i * stride) for i in range(acc["count"])]
        return np.stack(rows)
    return np.frombuffer(buf, dt, n * acc["count"], off).reshape(acc["count"], n)


def _node_matrix(node):
    if "matrix" in node:
        return np.array(node["matrix"], np.float64).reshape(4, 4).T
    T = np.eye(4)
    T[:3, 3] = node.get("translation", [0, 0, 0])
    R = np.eye(4)
    R[:3, :3] = b2scene.quat_to_mat3(node.get("rotation", [0, 0, 0, 1]))
    S = np.diag(list(node.get("scale", [1, 1, 1])) + [1.0])
    return T @ R @ S


def _material_record(m):
    """glTF material -> unpacked record, following glTF2Importer.cpp + SceneParser::read_material_properties."""
    pbr = m.get("pbrMetallicRoughness", {})
    ext = m.get("extensions", {})
    bc = pbr.get("baseColorFactor", [1, 1, 1, 1])
    kw = dict(base_color=bc[:3], alpha_opacity=bc[3],
              metallic=pbr.get("metallicFactor", 1.0), roughness=pbr.get("roughnessFactor", 1.0),
              emission=m.get("emissiveFactor", [0, 0, 0]))
    if "KHR_materials_emissive_strength" in ext:
        kw["emission_strength"] = ext["KHR_materials_emissive_strength"].get("emissiveStrength", 1.0)
    if "KHR_materials_specular" in ext:
        s = ext["KHR_materials_specular"]
        col = s.get("specularColorFactor", [1, 1, 1])
        fac = s.get("specularFactor", 1.0)
        # ASSIMP only emits the key when it differs from its defaults (glTF2Importer.cpp:288-291)
        if list(col) != [1, 1, 1] or fac != 0.0:
            kw.update(specular=fac, specular_tint=1.0, specular_color=[1, 1, 1])
    if "KHR_materials_sheen" in ext:
        s = ext["KHR_materials_sheen"]
        kw.update(sheen_color=s.get("sheenColorFactor", [0, 0, 0]), sheen_roughness=s.get("sheenRoughnessFactor", 0.0), sheen=1.0)
    if "KHR_materials_clearcoat" in ext:
        c = ext["KHR_materials_clearcoat"]
        kw.update(coat=c.get("clearcoatFactor", 0.0), coat_roughness=c.get("clearcoatRoughnessFactor", 0.0))
    if "KHR_materials_ior" in ext:
        kw["ior"] = ext["KHR_materials_ior"].get("ior", 1.5)
    if "KHR_materials_transmission" in ext:
        kw["specular_transmission"] = ext["KHR_materials_transmission"].get("transmissionFactor", 0.0)
    if "KHR_materials_volume" in ext:
        v = ext["KHR_materials_volume"]
        kw.update(absorption_color=v.get("attenuationColor", [1, 1, 1]))
        if "attenuationDistance" in v:
            kw["absorption_at_distance"] = v["attenuationDistance"]
    return b2scene.default_material(**kw)


def convert(gltf_path, out_path, width, height):
    gltf_path = Path(gltf_path)
    g = json.loads(gltf_path.read_text())
    bins = [(gltf_path.parent / b["uri"]).read_bytes() for b in g["buffers"]]

    # world matrices (flat scenes only need one level, but walk children anyway)
    world = {}

    def walk(i, parent):
        M = parent @ _node_matrix(g["nodes"][i])
        world[i] = M
        for c in g["nodes"][i].get("children", []):
            walk(c, M)

    for root in g["scenes"][g.get("scene", 0)]["nodes"]:
        walk(root, np.eye(4))

    prims = []  # (material index, world matrix, primitive)
    cam = None
    for i, node in enumerate(g["nodes"]):
        if i not in world:
            continue
        if "mesh" in node:
            for p in g["meshes"][node["mesh"]]["primitives"]:
                prims.append((p.get("material", 0), world[i], p))
        if "camera" in node and cam is None:
            cam = (g["cameras"][node["camera"]], world[i], node)
    prims.sort(key=lambda t: t[0])

    pos_l, nrm_l, uv_l, idx_l, mat_l, has_l = [], [], [], [], [], []
    offset = 0
    for mat_idx, M, p in prims:
        P = _accessor(g, bins, p["attributes"]["POSITION"]).astype(np.float64)
        Pw = (P @ M[:3, :3].T + M[:3, 3]).astype(np.float32)
        if "NORMAL" in p["attributes"]:
            N = _accessor(g, bins, p["attributes"]["NORMAL"]).astype(np.float64)
            Nw = N @ np.linalg.inv(M[:3, :3])          # (M^-T n) as row vectors
            Nw /= np.linalg.norm(Nw, axis=1, keepdims=True)
            Nw = Nw.astype(np.float32)
            has = np.ones(len(P), np.uint8)
        else:
            Nw = np.zeros((len(P), 3), np.float32)
            has = np.zeros(len(P), np.uint8)
        I = _accessor(g, bins, p["indices"]).astype(np.int64).reshape(-1, 3)
        pos_l.append(Pw)
        nrm_l.append(Nw)
        # the reference only keeps texcoords for meshes whose material has textures (SceneParser.cpp:136-140)
        uv_l.append(np.zeros((len(P), 2), np.float32))
        has_l.append(has)
        idx_l.append((I + offset).astype(np.int32))
        mat_l.append(np.full(len(I), mat_idx, np.int32))
        # reference quirk kept: offset advances by (max index + 1), not by the vertex count (SceneParser.cpp:155-199)
        # -- the vertex array still grows by mNumVertices, which only agree when every vertex is referenced.
        assert I.max() + 1 == len(P), "unreferenced trailing vertices: index offset and vertex count would disagree"
        offset += int(I.max()) + 1

    materials = np.stack([_material_record(m) for m in g["materials"]])

    aspect = width / height
    if cam is not None:
        c, M, node = cam
        persp = c["perspective"]
        # ASSIMP: mHorizontalFOV = 2 atan(tan(yfov/2) * aspectRatio) (glTF2Importer, ImportCameras)
        file_aspect = persp.get("aspectRatio", aspect)
        hfov = 2.0 * np.arctan(np.tan(persp["yfov"] * 0.5) * file_aspect)
        translation = M[:3, 3]
        rot = node.get("rotation", [0, 0, 0, 1])
        cam_bytes = b2scene.make_camera(translation, rot, hfov, aspect, clip_near=persp.get("znear", 0.1), clip_far=persp.get("zfar", 100.0))
    else:
        hfov = 40.0 / 180.0 * np.pi
        cam_bytes = b2scene.make_camera([0, 0, 0], [0, 0, 0, 1], hfov, aspect)

    b2scene.write_scene(out_path, cam_bytes, np.concatenate(pos_l), np.concatenate(nrm_l), np.concatenate(uv_l),
                        np.concatenate(idx_l), np.concatenate(mat_l), np.concatenate(has_l), materials)
    return dict(vertices=sum(len(p) for p in pos_l), triangles=sum(len(i) for i in idx_l), materials=len(materials))


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("gltf")
    ap.add_argument("out")
    ap.add_argument("--width", type=int, default=1920)
    ap.add_argument("--height", type=int, default=1080)
    a = ap.parse_args()
    print(convert(a.gltf, a.out, a.width, a.height))
