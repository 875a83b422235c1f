#!/usr/bin/env python
"""export_mitsuba_scene.py -- a procedural scene of drmlt_mitsuba_b200.scenes as a Mitsuba 0.6 scene directory: scene.xml + one Wavefront
OBJ per (material, emitter) mesh group + one PFM per bitmap texture ("each generator writes the flattened
binary for the ABI and Mitsuba XML + meshes, so that the real reference can render the same scene on a box that has it").

    python tools/export_mitsuba_scene.py C5 /tmp/door             # BASELINE configuration C1..C5 (bench.py CONFIGS), or
    python tools/export_mitsuba_scene.py textured /tmp/textured   # any scene of tests/ref_path_cases.py SCENES
    mitsuba /tmp/door/scene.xml -D integrator=drmlt -D technique=mmlt -D type=orbital -D maxDepth=8 -D directSamples=-1 -D spp=64

The XML uses only stock Mitsuba 0.6 plugins (obj, diffuse / dielectric / conductor / roughconductor / roughdielectric / plastic /
roughplastic / twosided, bitmap, area, perspective, hdrfilm, gaussian, independent) and parameter names; the integrator's own parameters
are left to `-D` with the reference's defaults.  Numbers are written with 17 significant digits: the double-precision reference then holds
exactly the float32 values the GPU path holds.
This repository cannot run the result (no Mitsuba binary, no XML loader in oracle/_ref); tests/test_export.py checks that the XML is
well-formed, names only stock plugins and the reference's parameters, and that the OBJ / PFM files reproduce the arrays bit for bit.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

BSDF_NAMES = ["diffuse", "dielectric", "conductor", "roughconductor", "roughdielectric", "plastic", "roughplastic"]
WRAP_NAMES = ["repeat", "clamp", "mirror", "zero", "one"]


def _f(v):
    return "%.17g" % float(v)        # the float32 value's exact decimal expansion: Mitsuba's double build parses back the same number


def _rgb(v):
    return ", ".join(_f(x) for x in v)


def write_pfm(path, texels):
    """[h, w, 3] float32, row v = 0 first -> PFM (little endian; PFM stores the BOTTOM row first)."""
    h, w = texels.shape[:2]
    with open(path, "wb") as f:
        f.write(b"PF\n%d %d\n-1.0\n" % (w, h))
        f.write(np.ascontiguousarray(texels[::-1], "<f4").tobytes())


def read_pfm(path):
    with open(path, "rb") as f:
        assert f.readline().strip() == b"PF"
        w, h = (int(x) for x in f.readline().split())
        assert float(f.readline()) < 0
        return np.frombuffer(f.read(), "<f4").reshape(h, w, 3)[::-1]


def write_obj(path, P, N, UV, tris):
    """Positions / optional normals / optional texture coordinates of the vertices the triangles use, re-indexed from 1."""
    used = np.unique(tris.reshape(-1))
    remap = {int(v): i + 1 for i, v in enumerate(used)}
    with open(path, "w") as f:
        for v in used:
            f.write("v %s %s %s\n" % tuple(_f(x) for x in P[v]))
        if UV is not None:
            for v in used:
                f.write("vt %s %s\n" % (_f(UV[v][0]), _f(UV[v][1])))
        if N is not None:
            for v in used:
                f.write("vn %s %s %s\n" % tuple(_f(x) for x in N[v]))
        for t in tris:
            idx = [remap[int(v)] for v in t]
            if UV is not None and N is not None:
                f.write("f %s\n" % " ".join("%d/%d/%d" % (i, i, i) for i in idx))
            elif UV is not None:
                f.write("f %s\n" % " ".join("%d/%d" % (i, i) for i in idx))
            elif N is not None:
                f.write("f %s\n" % " ".join("%d//%d" % (i, i) for i in idx))
            else:
                f.write("f %d %d %d\n" % tuple(idx))


def read_obj(path):
    P, N, UV, F = [], [], [], []
    for line in open(path):
        k = line.split()
        if not k:
            continue
        if k[0] == "v":
            P.append([float(x) for x in k[1:4]])
        elif k[0] == "vn":
            N.append([float(x) for x in k[1:4]])
        elif k[0] == "vt":
            UV.append([float(x) for x in k[1:3]])
        elif k[0] == "f":
            F.append([int(x.split("/")[0]) - 1 for x in k[1:4]])
    return np.array(P, np.float32), (np.array(N, np.float32) if N else None), (np.array(UV, np.float32) if UV else None), np.array(F, np.int64)


def bsdf_xml(m, data, textures, ind):
    from drmlt_mitsuba_b200 import abi
    name = BSDF_NAMES[m.type]
    tr, tt = (m.flags >> 8) & 0xfff, m.flags >> 20
    plastic = m.type in (abi.DR_BSDF_PLASTIC, abi.DR_BSDF_ROUGHPLASTIC)
    refl_name = "reflectance" if m.type == abi.DR_BSDF_DIFFUSE else ("diffuseReflectance" if plastic else "specularReflectance")
    trans_name = "specularReflectance" if plastic else "specularTransmittance"
    lines = []

    def colour(pname, value, tex):
        if tex:
            lines.append('%s\t<ref name="%s" id="texture%d"/>' % (ind, pname, tex - 1))
        else:
            lines.append('%s\t<rgb name="%s" value="%s"/>' % (ind, pname, _rgb(value)))
    colour(refl_name, m.reflectance, tr)
    if m.type in (abi.DR_BSDF_DIELECTRIC, abi.DR_BSDF_ROUGHDIELECTRIC) or plastic:
        colour(trans_name, m.transmittance, tt)
        lines.append('%s\t<float name="intIOR" value="%s"/>' % (ind, _f(m.eta[0])))
        lines.append('%s\t<float name="extIOR" value="1"/>' % ind)
    if m.type in (abi.DR_BSDF_CONDUCTOR, abi.DR_BSDF_ROUGHCONDUCTOR):
        lines.append('%s\t<string name="material" value="none"/>' % ind)
        lines.append('%s\t<rgb name="eta" value="%s"/>' % (ind, _rgb(m.eta)))
        lines.append('%s\t<rgb name="k" value="%s"/>' % (ind, _rgb(m.k)))
        lines.append('%s\t<float name="extEta" value="1"/>' % ind)
    if m.type in (abi.DR_BSDF_ROUGHCONDUCTOR, abi.DR_BSDF_ROUGHDIELECTRIC, abi.DR_BSDF_ROUGHPLASTIC):
        lines.append('%s\t<string name="distribution" value="%s"/>' % (ind, "ggx" if m.flags & abi.DR_MAT_GGX else "beckmann"))
        lines.append('%s\t<float name="alpha" value="%s"/>' % (ind, _f(m.alpha)))
        lines.append('%s\t<boolean name="sampleVisible" value="%s"/>' % (ind, "true" if m.flags & abi.DR_MAT_SAMPLE_VISIBLE else "false"))
    if plastic:
        lines.append('%s\t<boolean name="nonlinear" value="%s"/>' % (ind, "true" if m.flags & abi.DR_MAT_NONLINEAR else "false"))
    inner = '%s<bsdf type="%s">\n%s\n%s</bsdf>' % (ind, name, "\n".join(lines), ind)
    if m.flags & abi.DR_MAT_TWOSIDED:
        inner = '%s<bsdf type="twosided">\n%s\n%s</bsdf>' % (ind, inner.replace("\n", "\n\t").replace(ind + "<bsdf", ind + "\t<bsdf", 1), ind)
    return inner


def export(data, outdir, params=None):
    """SceneData -> outdir/scene.xml, mesh_*.obj, texture_*.pfm; returns the list of (obj file, material index, emitter index, triangles)."""
    from drmlt_mitsuba_b200 import abi
    os.makedirs(outdir, exist_ok=True)
    P, N, I, mat, emi, flg, mats, emis, rt = data.arrays()
    UV = data._uv
    # the integrator, as the reference's own scene files declare it (README.md:93-130): every parameter a `$name` that `-D name=value`
    # overrides, with the configuration's values as defaults; parameters the chosen integrator does not query only raise a warning
    # (scenehandler.cpp:792-796)
    defaults = dict(integrator="drmlt", technique="mmlt", type="orbital", maxDepth=8, rrDepth=5, directSamples=-1, directSampling=False, pLarge=0.3,
                    sigma=1.0 / 64, scaleSecond=0.1, timidAfterLarge=False, fixEmitterPath=False, useMixture=False, acceptanceMap=False,
                    kelemenStyleMutation=True, kelemenStyleWeights=True, lightImage=True, twoStage=False, spp=64)
    params = dict(params or {})
    rfilter = params.pop("rfilter", "gaussian")            # the film's reconstruction filter plugin (acceptanceMap needs "box", drmlt_proc.cpp:75-79)
    defaults.update(params)
    kinds = dict(technique="string", type="string", maxDepth="integer", rrDepth="integer", directSamples="integer", directSampling="boolean", pLarge="float",
                 sigma="float", scaleSecond="float", timidAfterLarge="boolean", fixEmitterPath="boolean", useMixture="boolean", acceptanceMap="boolean",
                 kelemenStyleMutation="boolean", kelemenStyleWeights="boolean", lightImage="boolean", twoStage="boolean")

    def val(v):
        return ("true" if v else "false") if isinstance(v, bool) else (repr(v) if isinstance(v, float) else str(v))
    xml = ['<?xml version="1.0" encoding="utf-8"?>', '<scene version="0.6.0">']
    xml += ['\t<default name="%s" value="%s"/>' % (k, val(v)) for k, v in defaults.items()]
    xml += ['\t<integrator type="$integrator">'] + ['\t\t<%s name="%s" value="$%s"/>' % (kinds[k], k, k) for k in kinds] + ['\t</integrator>', ""]
    for i, (t, arr) in enumerate(zip(data.textures, data._tex_keep)):
        write_pfm(os.path.join(outdir, "texture_%d.pfm" % i), arr)
        xml += ['\t<texture type="bitmap" id="texture%d">' % i, '\t\t<string name="filename" value="texture_%d.pfm"/>' % i,
                '\t\t<string name="filterType" value="%s"/>' % ("nearest" if t.nearest else "ewa"),
                '\t\t<string name="wrapModeU" value="%s"/>' % WRAP_NAMES[t.wrap_u], '\t\t<string name="wrapModeV" value="%s"/>' % WRAP_NAMES[t.wrap_v],
                '\t\t<float name="gamma" value="1"/>',
                '\t\t<float name="uscale" value="%s"/>' % repr(float(t.uv_scale[0])), '\t\t<float name="vscale" value="%s"/>' % repr(float(t.uv_scale[1])),
                '\t\t<float name="uoffset" value="%s"/>' % repr(float(t.uv_offset[0])), '\t\t<float name="voffset" value="%s"/>' % repr(float(t.uv_offset[1])),
                '\t</texture>', ""]
    groups = {}
    for t in range(len(I)):
        key = (int(mat[t]), int(emi[t]), bool(flg[t] & abi.DR_TRI_SMOOTH), UV is not None and not (flg[t] & abi.DR_TRI_NO_TEXCOORDS))
        groups.setdefault(key, []).append(t)
    written = []
    for g, (key, tris) in enumerate(sorted(groups.items())):
        m, e, smooth, has_uv = key
        fn = "mesh_%03d.obj" % g
        write_obj(os.path.join(outdir, fn), P, N if smooth else None, UV if has_uv else None, I[tris])
        xml += ['\t<shape type="obj">', '\t\t<string name="filename" value="%s"/>' % fn]
        if not smooth:
            xml.append('\t\t<boolean name="faceNormals" value="true"/>')
        if has_uv:
            xml.append('\t\t<boolean name="flipTexCoords" value="false"/>')      # (obj.cpp:215 flips v by default)
        xml.append(bsdf_xml(data.materials[m], data, data.textures, "\t\t"))
        if e >= 0:
            em = data.emitters[e]
            xml += ['\t\t<emitter type="area">', '\t\t\t<rgb name="radiance" value="%s"/>' % _rgb(em.radiance),
                    '\t\t\t<float name="samplingWeight" value="%s"/>' % _f(em.sampling_weight), '\t\t</emitter>']
        xml += ['\t</shape>', ""]
        written.append((fn, m, e, np.asarray(tris)))
    cam = data.camera
    M = np.array(cam.to_world[:], np.float64).reshape(4, 4)
    xml += ['\t<sensor type="perspective">', '\t\t<transform name="toWorld">',
            '\t\t\t<matrix value="%s"/>' % " ".join(_f(x) for x in M.reshape(-1)), '\t\t</transform>',
            '\t\t<float name="fov" value="%s"/>' % _f(cam.xfov_deg), '\t\t<string name="fovAxis" value="x"/>',
            '\t\t<float name="nearClip" value="%s"/>' % _f(cam.near_clip), '\t\t<float name="farClip" value="%s"/>' % _f(cam.far_clip),
            '\t\t<sampler type="independent">', '\t\t\t<integer name="sampleCount" value="$spp"/>', '\t\t</sampler>',
            '\t\t<film type="hdrfilm">', '\t\t\t<integer name="width" value="%d"/>' % cam.film_width, '\t\t\t<integer name="height" value="%d"/>' % cam.film_height,
            '\t\t\t<boolean name="banner" value="false"/>', '\t\t\t<rfilter type="%s"/>' % rfilter, '\t\t</film>', '\t</sensor>', '</scene>', ""]
    with open(os.path.join(outdir, "scene.xml"), "w") as f:
        f.write("\n".join(xml))
    return written


def main():
    name, outdir = sys.argv[1], sys.argv[2]
    import bench
    from drmlt_mitsuba_b200 import scenes
    if name in bench.CONFIGS:
        scene_name, kw, params = bench.CONFIGS[name]
        data = scenes.SCENES[scene_name](**kw)
        hint = "   # (defaults = the configuration: " + " ".join("-D %s=%s" % (k, str(v).lower() if isinstance(v, bool) else v) for k, v in params.items()) + ")"
    else:
        import ref_path_cases as RP
        data = RP.SCENES[name]()
        hint = "   # (defaults: -D integrator=drmlt -D technique=mmlt -D type=orbital -D maxDepth=8 -D directSamples=-1)"
    written = export(data, outdir, params if name in bench.CONFIGS else None)
    print("wrote %s/scene.xml, %d meshes (%d triangles), %d textures" % (outdir, len(written), data.n_triangles, len(data.textures)))
    print("mitsuba %s/scene.xml -D spp=64%s" % (outdir, hint))


if __name__ == "__main__":
    main()
