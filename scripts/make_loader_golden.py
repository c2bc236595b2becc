#!/usr/bin/env python3
"""Generate tests/golden/loader.json by running the REFERENCE's own loader code (scene.cpp +
tiny_obj_loader.h, compiled from /root/reference into oracle/_ref/libref_loader.so by `make ref`)
on the shipped coffee scene and on the small OBJ fixtures in tests/golden/obj/.  Runs only in the
build container (needs /root/reference); the JSON it writes is committed and is what the CPU test
suite checks our from-scratch Scene parser and OBJ reader against."""
import ctypes as C
import json
import os
import struct

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
REF_SCENES = "/root/reference/MinimalOptiX/scenes"


def main():
    L = C.CDLL(os.path.join(ROOT, "oracle", "_ref", "libref_loader.so"))
    L.ref_scene_open.restype = C.c_void_p
    L.ref_scene_open.argtypes = [C.c_char_p]
    L.ref_scene_close.argtypes = [C.c_void_p]
    L.ref_scene_counts.argtypes = [C.c_void_p] + [C.POINTER(C.c_int)] * 5
    L.ref_scene_mesh_name.restype = C.c_char_p
    L.ref_scene_mesh_name.argtypes = [C.c_void_p, C.c_int]
    L.ref_scene_texture.restype = C.c_char_p
    L.ref_scene_texture.argtypes = [C.c_void_p, C.c_int]
    L.ref_scene_material.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
    L.ref_scene_light.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
    L.ref_obj_load.argtypes = [C.c_char_p, C.POINTER(C.c_int)] + [C.POINTER(C.c_uint64)] * 3 + [C.c_uint64 * 3, C.c_uint64 * 16, C.c_uint64 * 16, C.c_int]
    L.ref_parse_double.argtypes = [C.c_char_p, C.POINTER(C.c_double)]

    out = {"_generated_by": "scripts/make_loader_golden.py from the reference's scene.cpp + tiny_obj_loader.h"}

    # --- coffee.scene through the reference's Scene class
    path = os.path.join(REF_SCENES, "coffee", "coffee.scene")
    h = L.ref_scene_open(path.encode())
    n = [C.c_int() for _ in range(5)]
    L.ref_scene_counts(h, *[C.byref(x) for x in n])
    meshes, materials, lights, width, height = [x.value for x in n]
    scene = {"meshes": meshes, "materials": materials, "lights": lights, "width": width, "height": height,
             "mesh_names": [], "textures": [], "material_bytes": [], "light_bytes": []}
    buf = C.create_string_buffer(72)
    for i in range(meshes):
        scene["mesh_names"].append(L.ref_scene_mesh_name(h, i).decode())
        scene["textures"].append(L.ref_scene_texture(h, i).decode())
        L.ref_scene_material(h, i, buf)
        scene["material_bytes"].append(buf.raw.hex())
    for i in range(lights):
        L.ref_scene_light(h, i, buf)
        raw = bytearray(buf.raw)
        raw[64:68] = b"\0\0\0\0"  # LightParams.radius is uninitialised for quads in the reference
        scene["light_bytes"].append(bytes(raw).hex())
    L.ref_scene_close(h)
    out["coffee_scene"] = scene

    # --- OBJ files through the reference's tinyobj
    def load(p):
        ns, nv, nn, nt = C.c_int(), C.c_uint64(), C.c_uint64(), C.c_uint64()
        ah, faces, ih = (C.c_uint64 * 3)(), (C.c_uint64 * 16)(), (C.c_uint64 * 16)()
        rc = L.ref_obj_load(p.encode(), C.byref(ns), C.byref(nv), C.byref(nn), C.byref(nt), ah, faces, ih, 16)
        assert rc == 0, p
        return {"shapes": ns.value, "vertices": nv.value, "normals": nn.value, "texcoords": nt.value,
                "attr_hash": [hex(x) for x in ah], "faces": list(faces)[:ns.value], "index_hash": [hex(x) for x in list(ih)[:ns.value]]}

    objs = {}
    cdir = os.path.join(REF_SCENES, "coffee")
    for name in sorted(os.listdir(cdir)):
        if name.endswith(".obj"):
            objs["coffee/" + name] = load(os.path.join(cdir, name))
    gdir = os.path.join(ROOT, "tests", "golden", "obj")
    for name in sorted(os.listdir(gdir)):
        if name.endswith(".obj"):
            objs["golden/" + name] = load(os.path.join(gdir, name))
    out["obj"] = objs

    # --- the float grammar
    texts = ["0", "-0", "1", "+1.5", "-2.5E+2", "1e-3", "0.1", "0.2", "0.3", "3.14159265358979", "123456789.123456789",
             "-0.000001234567", "1.", "7", "1e10", "1e-10", "0.30000001192092896", "16777217", "1.0000001", "0.811135",
             "-1.09417", "0.449693", "2.2250738585072014e-308", "1e308", "12345678901234567890", "0.000000000000000000001",
             "abc", ".5", "1e", "1e+", "--1", "1.5e3x", "1.5.5"]
    nums = []
    for t in texts:
        d = C.c_double(0.0)
        ok = L.ref_parse_double(t.encode(), C.byref(d))
        nums.append([t, int(ok), struct.pack("<d", d.value).hex()])
    out["parse_double"] = nums

    with open(os.path.join(ROOT, "tests", "golden", "loader.json"), "w") as f:
        json.dump(out, f, indent=1)
    print("wrote tests/golden/loader.json:", len(objs), "obj files,", meshes, "meshes,", lights, "lights")


if __name__ == "__main__":
    main()
