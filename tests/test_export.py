"""tools/export_mitsuba_scene.py: the procedural scenes as Mitsuba 0.6 scene directories, for a maintainer
whose box has the real `mitsuba` binary.  No Mitsuba loader exists here, so these tests hold what can be held: the XML is well-formed and names
only stock plugins and the reference's parameter names, and the OBJ / PFM files give back the arrays bit for bit."""
import os
import sys
import xml.etree.ElementTree as ET

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import export_mitsuba_scene as EX  # noqa: E402
from drmlt_mitsuba_b200 import abi, scenes  # noqa: E402

STOCK = {"obj", "diffuse", "dielectric", "conductor", "roughconductor", "roughdielectric", "plastic", "roughplastic", "twosided", "bitmap", "area",
         "perspective", "hdrfilm", "gaussian", "box", "independent", "$integrator"}


@pytest.mark.parametrize("name", ["textured", "plastic", "glossy", "C2"])
def test_exported_scene_directory_reproduces_the_arrays(name, tmp_path):
    import bench
    import ref_path_cases as RP
    params = None
    if name in bench.CONFIGS:
        scene_name, kw, params = bench.CONFIGS[name]
        data = scenes.SCENES[scene_name](**kw)
    else:
        data = RP.SCENES[name]()
    written = EX.export(data, str(tmp_path), params)
    P, N, I, mat, emi, flg, mats, emis, rt = data.arrays()
    UV = data._uv
    root = ET.parse(str(tmp_path / "scene.xml")).getroot()
    assert root.tag == "scene" and root.get("version") == "0.6.0"
    assert {e.get("type") for e in root.iter() if e.get("type")} <= STOCK
    shapes = root.findall("shape")
    assert len(shapes) == len(written) and sum(len(w[3]) for w in written) == len(I)
    # every mesh file gives back its triangles: positions, normals and texture coordinates bit for bit (17 significant digits)
    for shape, (fn, m, e, tris) in zip(shapes, written):
        assert shape.find("string[@name='filename']").get("value") == fn
        p, n, uv, f = EX.read_obj(str(tmp_path / fn))
        assert np.array_equal(p[f], P[I[tris]])
        smooth = bool(flg[tris[0]] & abi.DR_TRI_SMOOTH)
        assert (n is not None) == smooth and (shape.find("boolean[@name='faceNormals']") is None) == smooth
        if smooth:
            assert np.array_equal(n[f], N[I[tris]])
        has_uv = UV is not None and not (flg[tris[0]] & abi.DR_TRI_NO_TEXCOORDS)
        assert (uv is not None) == bool(has_uv)
        if has_uv:
            assert np.array_equal(uv[f], UV[I[tris]]) and shape.find("boolean[@name='flipTexCoords']").get("value") == "false"
        assert (shape.find("emitter") is not None) == (e >= 0)
        bsdf = shape.find("bsdf")
        want = data.materials[m]
        if want.flags & abi.DR_MAT_TWOSIDED:
            assert bsdf.get("type") == "twosided"
            bsdf = bsdf.find("bsdf")
        assert bsdf.get("type") == EX.BSDF_NAMES[want.type]
        refs = {r.get("name"): r.get("id") for r in bsdf.findall("ref")}
        assert len(refs) == bool((want.flags >> 8) & 0xfff) + bool(want.flags >> 20)
    # textures: PFM files that Mitsuba reads back as the same rows (it flips the file's bottom-up order, bitmap.cpp readPFM), with the lookup parameters
    texs = root.findall("texture")
    assert len(texs) == len(data.textures)
    for i, (t, arr) in enumerate(zip(data.textures, data._tex_keep)):
        assert np.array_equal(EX.read_pfm(str(tmp_path / ("texture_%d.pfm" % i))), arr)
        x = texs[i]
        assert x.get("id") == "texture%d" % i and x.find("string[@name='wrapModeU']").get("value") == EX.WRAP_NAMES[t.wrap_u]
        assert float(x.find("float[@name='uscale']").get("value")) == t.uv_scale[0] and float(x.find("float[@name='voffset']").get("value")) == t.uv_offset[1]
        assert (x.find("string[@name='filterType']").get("value") == "nearest") == bool(t.nearest)
    # the integrator: the reference's parameter names as $-substitutions with the configuration as defaults; the sensor and the film
    integ = root.find("integrator")
    assert integ.get("type") == "$integrator" and all(c.get("value") == "$" + c.get("name") for c in integ)
    defaults = {d.get("name"): d.get("value") for d in root.findall("default")}
    assert {c.get("name") for c in integ} <= set(defaults)
    for k, v in (params or {}).items():
        if k != "rfilter":
            assert defaults[k] == (("true" if v else "false") if isinstance(v, bool) else (repr(v) if isinstance(v, float) else str(v))), k
    film = root.find("sensor/film")
    assert (int(film.find("integer[@name='width']").get("value")), int(film.find("integer[@name='height']").get("value"))) == data.film
    assert film.find("rfilter").get("type") == ((params or {}).get("rfilter", "gaussian"))
    M = np.array([float(x) for x in root.find("sensor/transform/matrix").get("value").split()], np.float32)
    assert np.array_equal(M, np.array(data.camera.to_world[:], np.float32))
