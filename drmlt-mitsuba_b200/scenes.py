"""Procedural benchmark scenes C1..C5 (BASELINE.json `configs`).

All geometry is synthetic, generated from fixed seeds: triangles, constant-RGB materials, area
emitters on triangle meshes and a pinhole camera -- i.e. exactly what the Mitsuba-side shim
flattens out of a loaded `Scene` (shapes -> TriMesh arrays, BSDF/emitter parameters, sensor).
`SceneData.desc()` yields the `dr_scene_desc` of include/drmlt_b200.h.
"""
import ctypes as C
import math
import numpy as np

from . import abi


class SceneData:
    def __init__(self, name, film):
        self.name = name
        self.film = film
        self._P, self._N, self._I, self._mat, self._emi, self._flg = [], [], [], [], [], []
        self.materials, self.emitters = [], []
        self.rough_tables = []
        self._UV, self.textures, self._tex_keep = [], [], []
        self.has_uv = False
        self.n_vertices = 0
        self.n_triangles = 0
        self.camera = None
        self._keep = None

    # ---- construction helpers
    def add_material(self, type_, flags=0, reflectance=(0.5, 0.5, 0.5), transmittance=(1, 1, 1),
                     eta=(1.5, 0, 0), k=(0, 0, 0), alpha=0.1, rough_table=None, reflectance_tex=None, transmittance_tex=None):
        """reflectance_tex / transmittance_tex: index (add_texture) of a bitmap texture bound to that colour parameter; the constant is
        then the texture's average (only the sampling weights of plastic / roughplastic read it).
        rough_table (roughplastic): the DR_ROUGH_TABLE_DOUBLES doubles of include/drmlt_b200.h for this (distribution, eta, alpha)
        -- rough_tables.reduce(path to data/microfacet/<distribution>.dat, eta, alpha), or the reference's own RoughTransmittance."""
        m = abi.dr_material()
        if reflectance_tex is not None:
            flags |= abi.DR_MAT_TEX_REFLECTANCE(reflectance_tex)
            reflectance = self.textures[reflectance_tex]._avg
        if transmittance_tex is not None:
            flags |= abi.DR_MAT_TEX_TRANSMITTANCE(transmittance_tex)
            transmittance = self.textures[transmittance_tex]._avg
        m.type, m.flags = type_, flags
        m.reflectance[:] = reflectance
        m.transmittance[:] = transmittance
        m.eta[:] = eta
        m.k[:] = k
        m.alpha = alpha
        if type_ == abi.DR_BSDF_ROUGHPLASTIC:
            t = np.ascontiguousarray(rough_table, np.float64)
            if t.shape != (abi.DR_ROUGH_TABLE_DOUBLES,):
                raise ValueError("roughplastic needs a rough_table of %d doubles" % abi.DR_ROUGH_TABLE_DOUBLES)
            m.table = len(self.rough_tables)
            self.rough_tables.append(t)
        self.materials.append(m)
        return len(self.materials) - 1

    def add_texture(self, texels, wrap=abi.DR_WRAP_REPEAT, wrap_v=None, nearest=False, uv_scale=(1.0, 1.0), uv_offset=(0.0, 0.0)):
        """A bitmap texture (dr_texture): texels [h, w, 3] linear RGB float32, row v = 0 first."""
        arr = np.ascontiguousarray(texels, np.float32)
        t = abi.dr_texture()
        t.height, t.width = arr.shape[:2]
        t.texels = arr.ctypes.data_as(C.POINTER(C.c_float))
        t.wrap_u, t.wrap_v = wrap, wrap if wrap_v is None else wrap_v
        t.nearest = int(nearest)
        t.uv_scale[:] = uv_scale
        t.uv_offset[:] = uv_offset
        t._avg = tuple(float(x) for x in arr.reshape(-1, 3).astype(np.float64).mean(axis=0))
        self._tex_keep.append(arr)
        self.textures.append(t)
        return len(self.textures) - 1

    def add_mesh(self, P, I, material, N=None, radiance=None, sampling_weight=1.0, UV=None, uv_tangents=False):
        """P [nv,3] float, I [nt,3] int, optional vertex normals N and texture coordinates UV [nv,2].  `radiance` makes it an
        area emitter.  uv_tangents: the mesh carries UV tangents (DR_TRI_UV_TANGENTS; what Mitsuba does for meshes whose BSDF holds a
        filtered bitmap texture)."""
        P = np.asarray(P, np.float32).reshape(-1, 3)
        I = np.asarray(I, np.uint32).reshape(-1, 3)
        nt = I.shape[0]
        emitter = -1
        if radiance is not None:
            e = abi.dr_emitter()
            e.first_tri, e.n_tris = self.n_triangles, nt
            e.radiance[:] = radiance
            e.sampling_weight = sampling_weight
            self.emitters.append(e)
            emitter = len(self.emitters) - 1
        self._P.append(P)
        self._N.append(np.zeros_like(P) if N is None else np.asarray(N, np.float32).reshape(-1, 3))
        self._I.append(I + np.uint32(self.n_vertices))
        self._mat.append(np.full(nt, material, np.uint32))
        self._emi.append(np.full(nt, emitter, np.int32))
        self._UV.append(np.zeros((P.shape[0], 2), np.float32) if UV is None else np.asarray(UV, np.float32).reshape(-1, 2))
        self.has_uv |= UV is not None
        # (DR_TRI_NO_TEXCOORDS only matters in a scene where some other mesh has texture coordinates)
        self._flg.append(np.full(nt, (0 if N is None else abi.DR_TRI_SMOOTH) | (abi.DR_TRI_UV_TANGENTS if uv_tangents and UV is not None else 0) |
                                 (abi.DR_TRI_NO_TEXCOORDS if UV is None else 0), np.uint32))
        self.n_vertices += P.shape[0]
        self.n_triangles += nt

    def add_quad(self, a, b, c, d, material, nu=1, nv=1, uv=False, **kw):
        """Quad a,b,c,d (counter-clockwise seen from the front side), tessellated nu x nv.  uv=True: texture coordinates (0,0) at a,
        (1,0) at b, (1,1) at c, (0,1) at d."""
        a, b, c, d = [np.asarray(x, np.float64) for x in (a, b, c, d)]
        u = np.linspace(0, 1, nu + 1)[:, None, None]
        v = np.linspace(0, 1, nv + 1)[None, :, None]
        P = (1 - u) * (1 - v) * a + u * (1 - v) * b + u * v * c + (1 - u) * v * d
        idx = np.arange((nu + 1) * (nv + 1)).reshape(nu + 1, nv + 1)
        i00, i10, i11, i01 = idx[:-1, :-1], idx[1:, :-1], idx[1:, 1:], idx[:-1, 1:]
        I = np.concatenate([np.stack([i00, i10, i11], -1).reshape(-1, 3), np.stack([i00, i11, i01], -1).reshape(-1, 3)])
        if uv:
            kw["UV"] = np.stack(np.broadcast_arrays(u[..., 0], v[..., 0]), -1).reshape(-1, 2)
        self.add_mesh(P.reshape(-1, 3), I, material, **kw)

    def add_box(self, center, half, yrot_deg, material, tess=1, **kw):
        c = np.asarray(center, np.float64)
        hx, hy, hz = half
        th = math.radians(yrot_deg)
        R = np.array([[math.cos(th), 0, math.sin(th)], [0, 1, 0], [-math.sin(th), 0, math.cos(th)]])
        def V(x, y, z):
            return c + R @ np.array([x * hx, y * hy, z * hz])
        faces = [  # outward-facing
            (V(-1, -1, 1), V(1, -1, 1), V(1, 1, 1), V(-1, 1, 1)),      # +z
            (V(1, -1, -1), V(-1, -1, -1), V(-1, 1, -1), V(1, 1, -1)),  # -z
            (V(1, -1, 1), V(1, -1, -1), V(1, 1, -1), V(1, 1, 1)),      # +x
            (V(-1, -1, -1), V(-1, -1, 1), V(-1, 1, 1), V(-1, 1, -1)),  # -x
            (V(-1, 1, 1), V(1, 1, 1), V(1, 1, -1), V(-1, 1, -1)),      # +y
        ]
        for f in faces:
            self.add_quad(*f, material, nu=tess, nv=tess, **kw)

    def add_icosphere(self, center, radius, subdiv, material, smooth=True, **kw):
        P, I = _icosphere(subdiv)
        N = P.copy() if smooth else None
        self.add_mesh(P * radius + np.asarray(center, np.float64), I, material, N=N, **kw)

    def set_camera(self, origin, target, up, xfov_deg, near=1e-2, far=1e4):
        o, t, u = [np.asarray(x, np.float64) for x in (origin, target, up)]
        d = t - o
        d /= np.linalg.norm(d)
        left = np.cross(u, d)        # Mitsuba lookAt: left-handed camera frame (transform.cpp lookAt)
        left /= np.linalg.norm(left)
        newup = np.cross(d, left)
        M = np.eye(4)
        M[:3, 0], M[:3, 1], M[:3, 2], M[:3, 3] = left, newup, d, o
        cam = abi.dr_camera()
        cam.to_world[:] = M.astype(np.float32).reshape(-1).tolist()
        cam.xfov_deg, cam.near_clip, cam.far_clip = xfov_deg, near, far
        cam.film_width, cam.film_height = self.film
        self.camera = cam

    # ---- flattening
    def arrays(self):
        if self._keep is None:
            P = np.ascontiguousarray(np.concatenate(self._P), np.float32)
            N = np.ascontiguousarray(np.concatenate(self._N), np.float32)
            I = np.ascontiguousarray(np.concatenate(self._I), np.uint32)
            mat = np.ascontiguousarray(np.concatenate(self._mat), np.uint32)
            emi = np.ascontiguousarray(np.concatenate(self._emi), np.int32)
            flg = np.ascontiguousarray(np.concatenate(self._flg), np.uint32)
            mats = (abi.dr_material * len(self.materials))(*self.materials)
            emis = (abi.dr_emitter * max(1, len(self.emitters)))(*self.emitters)
            rt = np.ascontiguousarray(np.concatenate(self.rough_tables) if self.rough_tables else np.zeros(0), np.float64)
            self._uv = np.ascontiguousarray(np.concatenate(self._UV), np.float32) if self.has_uv else None
            self._texs = (abi.dr_texture * max(1, len(self.textures)))(*self.textures)
            self._keep = (P, N, I, mat, emi, flg, mats, emis, rt)
        return self._keep

    def desc(self):
        P, N, I, mat, emi, flg, mats, emis, rt = self.arrays()
        d = abi.dr_scene_desc()
        d.n_vertices, d.n_triangles = P.shape[0], I.shape[0]
        d.n_materials, d.n_emitters = len(self.materials), len(self.emitters)
        d.positions = P.ctypes.data_as(C.POINTER(C.c_float))
        d.normals = N.ctypes.data_as(C.POINTER(C.c_float))
        d.indices = I.ctypes.data_as(C.POINTER(C.c_uint32))
        d.tri_material = mat.ctypes.data_as(C.POINTER(C.c_uint32))
        d.tri_emitter = emi.ctypes.data_as(C.POINTER(C.c_int32))
        d.tri_flags = flg.ctypes.data_as(C.POINTER(C.c_uint32))
        d.materials = C.cast(mats, C.POINTER(abi.dr_material))
        d.emitters = C.cast(emis, C.POINTER(abi.dr_emitter))
        if len(self.rough_tables):
            d.rough_tables = rt.ctypes.data_as(C.POINTER(C.c_double))
            d.n_rough_tables = len(self.rough_tables)
        if self._uv is not None:
            d.texcoords = self._uv.ctypes.data_as(C.POINTER(C.c_float))
        if len(self.textures):
            d.textures = C.cast(self._texs, C.POINTER(abi.dr_texture))
            d.n_textures = len(self.textures)
        d.camera = self.camera
        return d


def _icosphere(subdiv):
    t = (1.0 + math.sqrt(5.0)) / 2.0
    V = np.array([[-1, t, 0], [1, t, 0], [-1, -t, 0], [1, -t, 0], [0, -1, t], [0, 1, t], [0, -1, -t], [0, 1, -t],
                  [t, 0, -1], [t, 0, 1], [-t, 0, -1], [-t, 0, 1]], np.float64)
    V /= np.linalg.norm(V, axis=1, keepdims=True)
    F = np.array([[0, 11, 5], [0, 5, 1], [0, 1, 7], [0, 7, 10], [0, 10, 11], [1, 5, 9], [5, 11, 4], [11, 10, 2],
                  [10, 7, 6], [7, 1, 8], [3, 9, 4], [3, 4, 2], [3, 2, 6], [3, 6, 8], [3, 8, 9], [4, 9, 5],
                  [2, 4, 11], [6, 2, 10], [8, 6, 7], [9, 8, 1]], np.int64)
    for _ in range(subdiv):
        e = np.concatenate([F[:, [0, 1]], F[:, [1, 2]], F[:, [2, 0]]])
        es = np.sort(e, axis=1)
        uniq, inv = np.unique(es, axis=0, return_inverse=True)
        inv = inv.reshape(-1)
        mid = V[uniq[:, 0]] + V[uniq[:, 1]]
        mid /= np.linalg.norm(mid, axis=1, keepdims=True)
        base = V.shape[0]
        V = np.concatenate([V, mid])
        n = F.shape[0]
        a, b, c = base + inv[:n], base + inv[n:2 * n], base + inv[2 * n:]
        F = np.concatenate([np.stack([F[:, 0], a, c], 1), np.stack([F[:, 1], b, a], 1),
                            np.stack([F[:, 2], c, b], 1), np.stack([a, b, c], 1)])
    return V, F


def _room(s, white, red, green, tess, light_half=0.25, radiance=(15.0, 15.0, 15.0), light_mat=None):
    """Cornell-style room [-1,1]^3 open towards +z, normals pointing inwards."""
    s.add_quad((-1, -1, 1), (1, -1, 1), (1, -1, -1), (-1, -1, -1), white, tess, tess)     # floor  (+y)
    s.add_quad((-1, 1, -1), (1, 1, -1), (1, 1, 1), (-1, 1, 1), white, tess, tess)         # ceiling (-y)
    s.add_quad((-1, -1, -1), (1, -1, -1), (1, 1, -1), (-1, 1, -1), white, tess, tess)     # back   (+z)
    s.add_quad((-1, -1, 1), (-1, -1, -1), (-1, 1, -1), (-1, 1, 1), red, tess, tess)       # left   (+x)
    s.add_quad((1, -1, -1), (1, -1, 1), (1, 1, 1), (1, 1, -1), green, tess, tess)         # right  (-x)
    h = light_half
    if light_mat is None:
        light_mat = white
    s.add_quad((-h, 0.995, -h), (h, 0.995, -h), (h, 0.995, h), (-h, 0.995, h), light_mat, radiance=radiance)   # faces -y


def cornell_box(film=(256, 256), tess=8, plastic=False, rough_tables=None):
    """C1 / C2: Cornell box, area light, one-sided diffuse walls, ~1k triangles.
    plastic=True: the two boxes are `plastic` (one linear, one nonlinear + twosided).
    rough_tables = (table of (beckmann, 1.49, 0.1), table of (ggx, 1.9, 0.3)): the boxes are `roughplastic` (Beckmann / GGX with
    visible-normal sampling, nonlinear, twosided) -- the tables come from rough_tables.reduce or the reference (include/drmlt_b200.h)."""
    s = SceneData("cornell-roughplastic" if rough_tables is not None else "cornell-plastic" if plastic else "cornell", film)
    white = s.add_material(abi.DR_BSDF_DIFFUSE, reflectance=(0.73, 0.73, 0.73))
    red = s.add_material(abi.DR_BSDF_DIFFUSE, reflectance=(0.63, 0.065, 0.05))
    green = s.add_material(abi.DR_BSDF_DIFFUSE, reflectance=(0.14, 0.45, 0.091))
    _room(s, white, red, green, tess)
    box1 = box2 = white
    if plastic:       # reflectance = diffuseReflectance, transmittance = specularReflectance (include/drmlt_b200.h)
        box1 = s.add_material(abi.DR_BSDF_PLASTIC, reflectance=(0.1, 0.27, 0.36), transmittance=(1, 1, 1), eta=(1.49, 0, 0))
        box2 = s.add_material(abi.DR_BSDF_PLASTIC, flags=abi.DR_MAT_NONLINEAR | abi.DR_MAT_TWOSIDED, reflectance=(0.6, 0.5, 0.2),
                              transmittance=(0.9, 0.9, 1.0), eta=(1.9, 0, 0))
    if rough_tables is not None:
        box1 = s.add_material(abi.DR_BSDF_ROUGHPLASTIC, reflectance=(0.1, 0.27, 0.36), transmittance=(1, 1, 1), eta=(1.49, 0, 0), alpha=0.1,
                              rough_table=rough_tables[0])
        box2 = s.add_material(abi.DR_BSDF_ROUGHPLASTIC, flags=abi.DR_MAT_GGX | abi.DR_MAT_SAMPLE_VISIBLE | abi.DR_MAT_NONLINEAR | abi.DR_MAT_TWOSIDED,
                              reflectance=(0.6, 0.5, 0.2), transmittance=(0.9, 0.9, 1.0), eta=(1.9, 0, 0), alpha=0.3, rough_table=rough_tables[1])
    s.add_box((0.33, -0.7, 0.35), (0.3, 0.3, 0.3), -17.0, box1, tess=4)
    s.add_box((-0.33, -0.4, -0.3), (0.3, 0.6, 0.3), 17.0, box2, tess=4)
    s.set_camera((0, 0, 3.9), (0, 0, 0), (0, 1, 0), 39.0)
    return s


def procedural_texels(w, h, seed, lo=0.05, hi=0.9, cell=4):
    """A deterministic colour bitmap [h, w, 3] float32: blocky random colours (cell x cell texels) plus per-texel noise, so that
    bilinear, nearest and the wrap modes all give visibly different answers."""
    rng = np.random.RandomState(seed)
    coarse = rng.uniform(lo, hi, ((h + cell - 1) // cell, (w + cell - 1) // cell, 3))
    img = np.repeat(np.repeat(coarse, cell, axis=0), cell, axis=1)[:h, :w]
    img = np.clip(img + rng.uniform(-0.04, 0.04, img.shape), 0.01, 0.95)
    # rounded to half precision: the reference keeps its MIP levels as halfs (bitmap.cpp:175-177), so these texels are exactly
    # what its texture holds
    return np.ascontiguousarray(img.astype(np.float16), np.float32)


def cornell_box_textured(film=(256, 256), tess=8, uv_tangents=True):
    """The Cornell box with bitmap textures: floor = diffuse with a bilinear, repeating texture (uscale = vscale = 2.5,
    offset), back wall = diffuse, mirror-wrapped, on a mesh with UV tangents, left wall = rough conductor with a textured
    specularReflectance (clamp), one box = plastic with textured diffuseReflectance AND specularReflectance, the other = diffuse with a nearest-filtered
    texture (zero / one wrap), a glass pane with both dielectric colours textured; everything else as cornell_box.  uv_tangents=True flags every mesh DR_TRI_UV_TANGENTS -- what the
    reference does for any mesh with texture coordinates (trimesh.cpp:400-402); False keeps the edge-based shading frames.  The ceiling
    and the right wall carry no texture coordinates (a mixed scene: DR_TRI_NO_TEXCOORDS, barycentric uv, edge-based frames)."""
    s = SceneData("cornell-textured", film)
    T = bool(uv_tangents)
    t_floor = s.add_texture(procedural_texels(32, 24, 11), wrap=abi.DR_WRAP_REPEAT, uv_scale=(2.5, 2.5), uv_offset=(0.125, -0.3))
    t_back = s.add_texture(procedural_texels(17, 29, 12), wrap=abi.DR_WRAP_MIRROR, uv_scale=(1.7, 1.3), uv_offset=(-0.2, 0.1))
    t_left = s.add_texture(procedural_texels(16, 16, 13, lo=0.4, hi=0.95), wrap=abi.DR_WRAP_CLAMP, uv_scale=(1.5, 1.5), uv_offset=(-0.25, -0.25))
    t_box1 = s.add_texture(procedural_texels(8, 8, 14, lo=0.1, hi=0.7, cell=2), wrap=abi.DR_WRAP_REPEAT)
    t_box1s = s.add_texture(procedural_texels(6, 10, 16, lo=0.5, hi=0.95, cell=2), wrap=abi.DR_WRAP_MIRROR, uv_scale=(2.0, 1.0))
    t_box2 = s.add_texture(procedural_texels(12, 6, 15, cell=3), wrap=abi.DR_WRAP_ZERO, wrap_v=abi.DR_WRAP_ONE, nearest=True,
                           uv_scale=(1.25, 1.25), uv_offset=(-0.125, -0.125))
    white = s.add_material(abi.DR_BSDF_DIFFUSE, reflectance=(0.73, 0.73, 0.73))
    green = s.add_material(abi.DR_BSDF_DIFFUSE, reflectance=(0.14, 0.45, 0.091))
    floor = s.add_material(abi.DR_BSDF_DIFFUSE, reflectance_tex=t_floor)
    back = s.add_material(abi.DR_BSDF_DIFFUSE, reflectance_tex=t_back)
    left = s.add_material(abi.DR_BSDF_ROUGHCONDUCTOR, flags=abi.DR_MAT_GGX, reflectance_tex=t_left, eta=(0.2, 0.92, 1.1), k=(3.9, 2.45, 2.14), alpha=0.25)
    # plastic with BOTH colour parameters textured (diffuseReflectance -> `reflectance`, specularReflectance -> `transmittance`)
    box1 = s.add_material(abi.DR_BSDF_PLASTIC, reflectance_tex=t_box1, transmittance_tex=t_box1s, eta=(1.49, 0, 0))
    box2 = s.add_material(abi.DR_BSDF_DIFFUSE, flags=abi.DR_MAT_TWOSIDED, reflectance_tex=t_box2)
    # a glass pane whose specularReflectance AND specularTransmittance are textured (dielectric: `reflectance`, then `transmittance`)
    pane = s.add_material(abi.DR_BSDF_DIELECTRIC, reflectance_tex=t_box1s, transmittance_tex=t_left, eta=(1.5, 0, 0))
    s.add_quad((-0.95, -0.3, 0.75), (-0.35, -0.3, 0.75), (-0.35, 0.55, 0.75), (-0.95, 0.55, 0.75), pane, 2, 2, uv=True, uv_tangents=T)
    s.add_quad((-1, -1, 1), (1, -1, 1), (1, -1, -1), (-1, -1, -1), floor, tess, tess, uv=True, uv_tangents=T)
    s.add_quad((-1, 1, -1), (1, 1, -1), (1, 1, 1), (-1, 1, 1), white, tess, tess)         # (no texture coordinates: DR_TRI_NO_TEXCOORDS)
    s.add_quad((-1, -1, -1), (1, -1, -1), (1, 1, -1), (-1, 1, -1), back, tess, tess, uv=True, uv_tangents=T)
    s.add_quad((-1, -1, 1), (-1, -1, -1), (-1, 1, -1), (-1, 1, 1), left, tess, tess, uv=True, uv_tangents=T)
    s.add_quad((1, -1, -1), (1, -1, 1), (1, 1, 1), (1, 1, -1), green, tess, tess)
    h = 0.25
    # the light: every vertex at the same uv -- a degenerate parameterisation, for which the reference picks arbitrary tangents
    # perpendicular to the normal (trimesh.cpp:750-754)
    s.add_mesh([(-h, 0.995, -h), (h, 0.995, -h), (h, 0.995, h), (-h, 0.995, h)], [(0, 1, 2), (0, 2, 3)], white, radiance=(15.0, 15.0, 15.0),
               UV=[(0.5, 0.5)] * 4, uv_tangents=T)
    s.add_box((0.33, -0.7, 0.35), (0.3, 0.3, 0.3), -17.0, box1, tess=4, uv=True, uv_tangents=T)
    s.add_box((-0.33, -0.4, -0.3), (0.3, 0.6, 0.3), 17.0, box2, tess=4, uv=True, uv_tangents=T)
    s.set_camera((0, 0, 3.9), (0, 0, 0), (0, 1, 0), 39.0)
    return s


def glossy_scene(film=(512, 512), subdiv=5, rough_glass=None):
    """C3: room + three GGX rough-conductor spheres + one dielectric sphere, ~100k triangles.
    rough_glass = (alpha, flags): the glass sphere and a frosted pane in front of the back wall become `roughdielectric`
    (SURVEY 8f rank 4)."""
    s = SceneData("glossy", film)
    white = s.add_material(abi.DR_BSDF_DIFFUSE, reflectance=(0.73, 0.73, 0.73))
    red = s.add_material(abi.DR_BSDF_DIFFUSE, reflectance=(0.63, 0.065, 0.05))
    green = s.add_material(abi.DR_BSDF_DIFFUSE, reflectance=(0.14, 0.45, 0.091))
    _room(s, white, red, green, 24)
    cu_eta, cu_k = (0.2004, 0.9240, 1.1022), (3.9129, 2.4528, 2.1421)
    for i, alpha in enumerate((0.05, 0.1, 0.3)):
        m = s.add_material(abi.DR_BSDF_ROUGHCONDUCTOR, flags=abi.DR_MAT_GGX | abi.DR_MAT_SAMPLE_VISIBLE,
                           reflectance=(1, 1, 1), eta=cu_eta, k=cu_k, alpha=alpha)
        s.add_icosphere((-0.6 + 0.6 * i, -0.7, -0.3 + 0.1 * i), 0.3, subdiv, m)
    if rough_glass is None:
        glass = s.add_material(abi.DR_BSDF_DIELECTRIC, reflectance=(1, 1, 1), transmittance=(1, 1, 1), eta=(1.5, 0, 0))
    else:
        s.name = "glossy-roughglass"
        glass = s.add_material(abi.DR_BSDF_ROUGHDIELECTRIC, flags=rough_glass[1], reflectance=(1, 1, 1), transmittance=(1, 1, 1),
                               eta=(1.5, 0, 0), alpha=rough_glass[0])
        # a thin frosted pane (two faces, outward normals) between the camera and the back wall
        z0, z1 = -0.55, -0.5
        s.add_quad((-0.9, -0.2, z1), (0.9, -0.2, z1), (0.9, 0.8, z1), (-0.9, 0.8, z1), glass)
        s.add_quad((0.9, -0.2, z0), (-0.9, -0.2, z0), (-0.9, 0.8, z0), (0.9, 0.8, z0), glass)
    s.add_icosphere((0.1, -0.65, 0.45), 0.35, subdiv, glass)
    s.set_camera((0, 0, 3.9), (0, 0, 0), (0, 1, 0), 39.0)
    return s


def caustic_scene(film=(512, 512), grid=220, seed=7):
    """C4: small emitter above a wavy dielectric slab over a diffuse floor, ~100k triangles."""
    rng = np.random.RandomState(seed)
    s = SceneData("caustic", film)
    white = s.add_material(abi.DR_BSDF_DIFFUSE, reflectance=(0.73, 0.73, 0.73))
    grey = s.add_material(abi.DR_BSDF_DIFFUSE, reflectance=(0.4, 0.4, 0.4))
    glass = s.add_material(abi.DR_BSDF_DIELECTRIC, eta=(1.5, 0, 0))
    _room(s, white, grey, grey, 8, light_half=0.01, radiance=(9000.0, 9000.0, 9000.0))
    # slab: x,z in [-0.8,0.8], bottom y=-0.3, wavy top around y=-0.2
    n = grid
    x = np.linspace(-0.8, 0.8, n + 1)
    z = np.linspace(-0.8, 0.8, n + 1)
    X, Z = np.meshgrid(x, z, indexing="ij")
    ph = rng.uniform(0, 2 * math.pi, 4)
    Y = -0.2 + 0.02 * np.sin(9 * X + ph[0]) * np.sin(7 * Z + ph[1]) + 0.01 * np.sin(17 * X + 13 * Z + ph[2])
    top = np.stack([X, Y, Z], -1).reshape(-1, 3)
    idx = np.arange((n + 1) * (n + 1)).reshape(n + 1, n + 1)
    i00, i10, i11, i01 = idx[:-1, :-1], idx[1:, :-1], idx[1:, 1:], idx[:-1, 1:]
    # +y facing: (x,z) grid with normal up => order i00, i01, i11 ...
    I = np.concatenate([np.stack([i00, i01, i11], -1).reshape(-1, 3), np.stack([i00, i11, i10], -1).reshape(-1, 3)])
    # smooth normals from the analytic height field gradient (finite differences)
    gx = np.gradient(Y, x, axis=0)
    gz = np.gradient(Y, z, axis=1)
    Nn = np.stack([-gx, np.ones_like(gx), -gz], -1).reshape(-1, 3)
    Nn /= np.linalg.norm(Nn, axis=1, keepdims=True)
    s.add_mesh(top, I, glass, N=Nn)
    yb = -0.3
    s.add_quad((-0.8, yb, -0.8), (0.8, yb, -0.8), (0.8, yb, 0.8), (-0.8, yb, 0.8), glass)   # bottom faces -y
    # side strips following the wavy boundary
    def strip(px, pz, py, flip):
        m = len(px)
        P = np.concatenate([np.stack([px, py, pz], -1), np.stack([px, np.full(m, yb), pz], -1)])
        a, b = np.arange(m - 1), np.arange(1, m)
        tri = np.concatenate([np.stack([a, a + m, b + m], -1), np.stack([a, b + m, b], -1)])
        if flip:
            tri = tri[:, ::-1]
        s.add_mesh(P, tri, glass)
    strip(x, np.full(n + 1, z[0]), Y[:, 0], True)
    strip(x, np.full(n + 1, z[-1]), Y[:, -1], False)
    strip(np.full(n + 1, x[0]), z, Y[0, :], False)
    strip(np.full(n + 1, x[-1]), z, Y[-1, :], True)
    s.set_camera((0, 0.6, 3.6), (0, -0.5, 0), (0, 1, 0), 39.0)
    return s


def door_scene(film=(1280, 720), floor_grid=580, n_spheres=64, sphere_subdiv=4, seed=42):
    """C5: occluded-light 'door' scene, ~1M triangles.  The emitter sits in an adjacent room; light
    reaches the main room only through a door that is ajar."""
    rng = np.random.RandomState(seed)
    s = SceneData("door", film)
    white = s.add_material(abi.DR_BSDF_DIFFUSE, reflectance=(0.7, 0.7, 0.7))
    wood = s.add_material(abi.DR_BSDF_DIFFUSE, flags=abi.DR_MAT_TWOSIDED, reflectance=(0.45, 0.3, 0.15))
    blue = s.add_material(abi.DR_BSDF_DIFFUSE, reflectance=(0.2, 0.3, 0.6))
    metal = s.add_material(abi.DR_BSDF_ROUGHCONDUCTOR, flags=abi.DR_MAT_GGX | abi.DR_MAT_SAMPLE_VISIBLE,
                           reflectance=(1, 1, 1), eta=(0.2004, 0.9240, 1.1022), k=(3.9129, 2.4528, 2.1421), alpha=0.2)
    # main room: x in [-2,2], y in [-1,1], z in [-2,2]; adjacent room: x in [2,4]
    W = 12
    s.add_quad((-2, 1, -2), (2, 1, -2), (2, 1, 2), (-2, 1, 2), white, W, W)           # ceiling
    s.add_quad((-2, -1, -2), (2, -1, -2), (2, 1, -2), (-2, 1, -2), white, W, W)       # back (+z)
    s.add_quad((2, -1, 2), (-2, -1, 2), (-2, 1, 2), (2, 1, 2), white, W, W)           # front (-z), behind the camera
    s.add_quad((-2, -1, 2), (-2, -1, -2), (-2, 1, -2), (-2, 1, 2), blue, W, W)        # left (+x)
    # right wall x=2 with a door opening z in [-0.4,0.4], y in [-1,0.6]; two-sided so it also bounds room 2
    wall = s.add_material(abi.DR_BSDF_DIFFUSE, flags=abi.DR_MAT_TWOSIDED, reflectance=(0.7, 0.7, 0.7))
    s.add_quad((2, -1, -2), (2, -1, -0.4), (2, 1, -0.4), (2, 1, -2), wall, W, W)
    s.add_quad((2, -1, 0.4), (2, -1, 2), (2, 1, 2), (2, 1, 0.4), wall, W, W)
    s.add_quad((2, 0.6, -0.4), (2, 0.6, 0.4), (2, 1, 0.4), (2, 1, -0.4), wall, 4, 4)
    # door leaf hinged at z=-0.4, ajar by 10 degrees
    a = math.radians(10.0)
    hz, hx = -0.4, 2.0
    ex, ez = hx + 0.8 * math.sin(a), hz + 0.8 * math.cos(a)
    s.add_quad((hx, -1, hz), (ex, -1, ez), (ex, 0.6, ez), (hx, 0.6, hz), wood, 8, 8)
    # adjacent room
    s.add_quad((2, -1, 2), (4, -1, 2), (4, -1, -2), (2, -1, -2), white, W, W)         # floor
    s.add_quad((2, 1, -2), (4, 1, -2), (4, 1, 2), (2, 1, 2), white, W, W)             # ceiling
    s.add_quad((2, -1, -2), (4, -1, -2), (4, 1, -2), (2, 1, -2), white, W, W)         # back
    s.add_quad((4, -1, 2), (2, -1, 2), (2, 1, 2), (4, 1, 2), white, W, W)             # front
    s.add_quad((4, -1, -2), (4, -1, 2), (4, 1, 2), (4, 1, -2), white, W, W)           # far right (-x)
    s.add_quad((2.7, 0.99, -0.3), (3.3, 0.99, -0.3), (3.3, 0.99, 0.3), (2.7, 0.99, 0.3), white,
               radiance=(60.0, 55.0, 45.0))                                            # emitter facing down
    # displaced-noise floor of the main room
    n = floor_grid
    x = np.linspace(-2, 2, n + 1)
    z = np.linspace(-2, 2, n + 1)
    X, Z = np.meshgrid(x, z, indexing="ij")
    Y = np.full_like(X, -1.0)
    for _ in range(6):
        fx, fz = rng.uniform(2, 14, 2)
        px, pz = rng.uniform(0, 2 * math.pi, 2)
        Y += 0.012 * np.sin(fx * X + px) * np.sin(fz * Z + pz)
    P = np.stack([X, Y, Z], -1).reshape(-1, 3)
    idx = np.arange((n + 1) * (n + 1)).reshape(n + 1, n + 1)
    i00, i10, i11, i01 = idx[:-1, :-1], idx[1:, :-1], idx[1:, 1:], idx[:-1, 1:]
    I = np.concatenate([np.stack([i00, i01, i11], -1).reshape(-1, 3), np.stack([i00, i11, i10], -1).reshape(-1, 3)])
    gx = np.gradient(Y, x, axis=0)
    gz = np.gradient(Y, z, axis=1)
    Nn = np.stack([-gx, np.ones_like(gx), -gz], -1).reshape(-1, 3)
    Nn /= np.linalg.norm(Nn, axis=1, keepdims=True)
    s.add_mesh(P, I, white, N=Nn)
    # icospheres scattered over the floor ("instanced as copies")
    k = int(round(math.sqrt(n_spheres)))
    for i in range(n_spheres):
        gx_, gz_ = i % k, i // k
        cx = -1.6 + 3.2 * (gx_ + 0.5) / k + rng.uniform(-0.08, 0.08)
        cz = -1.6 + 3.2 * (gz_ + 0.5) / k + rng.uniform(-0.08, 0.08)
        r = rng.uniform(0.09, 0.16)
        s.add_icosphere((cx, -0.97 + r, cz), r, sphere_subdiv, metal if i % 3 == 0 else white)
    s.set_camera((-1.7, 0.2, 1.9), (1.4, -0.5, -0.4), (0, 1, 0), 60.0)
    return s


SCENES = {"cornell": cornell_box, "glossy": glossy_scene, "caustic": caustic_scene, "door": door_scene}
