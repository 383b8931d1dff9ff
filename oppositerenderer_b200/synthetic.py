"""Seeded procedural scenes of the sizes BASELINE.json names (the Sponza / Conference downloads are not
available offline): a closed room of ~3000 x 1200 x 1500 units (centimetre scale, so that
IScene::getSceneInitialPPMRadiusEstimate gives the regime the reference runs Sponza in), a colonnade of
tessellated columns, draped displaced-plane curtains, tessellated glass / mirror spheres, procedural RGBA8
textures (value noise, albedo in [0.2, 0.8]) with a few normal maps, one large quad emitter and one point light.
One material per mesh, as scene/Scene.cpp:496-503 does for Collada meshes.
"""
import numpy as np

from . import capi
from .renderer import Camera, Light, SceneDescription, make_material

ROOM = (3000.0, 1200.0, 1500.0)


def _grid(fn, nu, nv, flip=False, uv_scale=(1.0, 1.0)):
    """Tessellate p = fn(u, v), u,v in [0,1], into (nu x nv) quads -> dict(vertices, normals, uvs, tangents, bitangents, indices)."""
    u = np.linspace(0.0, 1.0, nu + 1, dtype=np.float64)
    v = np.linspace(0.0, 1.0, nv + 1, dtype=np.float64)
    U, V = np.meshgrid(u, v, indexing="xy")
    P = fn(U, V)                                   # (nv+1, nu+1, 3)
    e = 1e-4
    Tu = (fn(np.clip(U + e, 0, 1), V) - fn(np.clip(U - e, 0, 1), V))
    Tv = (fn(U, np.clip(V + e, 0, 1)) - fn(U, np.clip(V - e, 0, 1)))
    N = np.cross(Tu, Tv)
    if flip:
        N = -N
    def unit(a):
        n = np.linalg.norm(a, axis=-1, keepdims=True)
        return a / np.maximum(n, 1e-20)
    N, Tu, Tv = unit(N), unit(Tu), unit(Tv)
    idx = (np.arange(nv)[:, None] * (nu + 1) + np.arange(nu)[None, :]).reshape(-1)
    a, b, c, d = idx, idx + 1, idx + nu + 2, idx + nu + 1
    tri = np.stack([np.stack([a, b, c], 1), np.stack([a, c, d], 1)], 1).reshape(-1, 3)
    if flip:
        tri = tri[:, ::-1]
    uv = np.stack([U * uv_scale[0], V * uv_scale[1]], -1)
    f32 = lambda x: np.ascontiguousarray(x.reshape(-1, x.shape[-1]), np.float32)
    return dict(vertices=f32(P), normals=f32(N), uvs=f32(uv), tangents=f32(Tu), bitangents=f32(Tv), indices=np.ascontiguousarray(tri, np.int32))


def _value_noise(rng, size, octaves=5):
    img = np.zeros((size, size), np.float32)
    amp, total = 1.0, 0.0
    for o in range(octaves):
        n = 4 << o
        g = rng.random((n, n), dtype=np.float32)
        g = np.concatenate([g, g[:1]], 0)
        g = np.concatenate([g, g[:, :1]], 1)
        x = np.linspace(0, n, size, endpoint=False, dtype=np.float32)
        xi = x.astype(np.int32)
        xf = x - xi
        xf = xf * xf * (3 - 2 * xf)
        a = g[xi][:, xi]; b = g[xi][:, xi + 1]; c = g[xi + 1][:, xi]; d = g[xi + 1][:, xi + 1]
        img += amp * ((a * (1 - xf)[None, :] + b * xf[None, :]) * (1 - xf)[:, None] + (c * (1 - xf)[None, :] + d * xf[None, :]) * xf[:, None])
        total += amp
        amp *= 0.5
    return img / total


def _albedo_texture(rng, size, tint):
    n = _value_noise(rng, size)
    rgb = 0.2 + 0.6 * np.clip(n[..., None] * 0.7 + 0.3 * np.asarray(tint, np.float32), 0, 1)
    out = np.empty((size, size, 4), np.uint8)
    out[..., :3] = (rgb * 255 + 0.5).astype(np.uint8)
    out[..., 3] = 255
    return out


def _normal_texture(rng, size, strength=2.0):
    h = _value_noise(rng, size)
    dx = (np.roll(h, -1, 1) - np.roll(h, 1, 1)) * strength * size / 64.0
    dy = (np.roll(h, -1, 0) - np.roll(h, 1, 0)) * strength * size / 64.0
    n = np.stack([-dx, -dy, np.ones_like(h)], -1)
    n /= np.linalg.norm(n, axis=-1, keepdims=True)
    out = np.empty((size, size, 4), np.uint8)
    out[..., :3] = ((n * 0.5 + 0.5) * 255 + 0.5).astype(np.uint8)
    out[..., 3] = 255
    return out


def make_scene(kind="sponza", seed=1, detail=1.0, texture_size=1024, textured=True):
    """kind: 'sponza' (~1M triangles at detail=1, textured; BASELINE configs 4/5) or 'conference' (~300K triangles,
    glass/mirror caustics; config 3).  detail scales the tessellation (triangle count ~ detail^2)."""
    rng = np.random.default_rng(seed)
    X, Y, Z = ROOM
    s = float(detail)
    q = lambda n: max(2, int(round(n * s)))
    meshes, materials, textures = [], [], []
    normal_mapped = set()      # materials with a normal map: their meshes must carry tangents / bitangents (Texture.cu:92-97)

    def tex_material(tint, normal=False):
        if not textured:
            materials.append(make_material(capi.MAT_DIFFUSE, Kd=tuple(0.2 + 0.6 * np.clip(np.asarray(tint), 0, 1))))
            return len(materials) - 1
        textures.append(_albedo_texture(rng, texture_size, tint))
        d = len(textures) - 1
        nidx = -1
        if normal:
            textures.append(_normal_texture(rng, texture_size))
            nidx = len(textures) - 1
        materials.append(make_material(capi.MAT_TEXTURE, diffuseTexture=d, normalTexture=nidx))
        if normal:
            normal_mapped.add(len(materials) - 1)
        return len(materials) - 1

    def add(mesh, mat, keep_tangents=False):
        # assimp's CalcTangentSpace (scene/Scene.cpp:96) gives every mesh with normals and uvs a tangent frame; here only the
        # normal-mapped meshes keep theirs (nothing else reads them)
        if not keep_tangents and mat not in normal_mapped:
            mesh["tangents"] = mesh["bitangents"] = None
        mesh["material"] = mat
        meshes.append(mesh)

    sponza = kind == "sponza"
    # ---- room shell (normals point inward)
    floor_mat = tex_material((0.8, 0.7, 0.6), normal=True)
    wall_mat = tex_material((0.9, 0.85, 0.7))
    ceil_mat = tex_material((0.9, 0.9, 0.9))
    add(_grid(lambda u, v: np.stack([u * X, 2.0 * np.sin(u * 40) * np.sin(v * 25), v * Z], -1), q(384), q(192), flip=True, uv_scale=(12, 6)), floor_mat, True)
    add(_grid(lambda u, v: np.stack([u * X, np.full_like(u, Y), v * Z], -1), q(128), q(64), uv_scale=(6, 3)), ceil_mat)
    add(_grid(lambda u, v: np.stack([u * X, v * Y, np.full_like(u, Z)], -1), q(128), q(64), flip=True, uv_scale=(6, 3)), wall_mat)
    add(_grid(lambda u, v: np.stack([u * X, v * Y, np.zeros_like(u)], -1), q(128), q(64), uv_scale=(6, 3)), wall_mat)
    add(_grid(lambda u, v: np.stack([np.zeros_like(u), v * Y, u * Z], -1), q(64), q(64), flip=True, uv_scale=(3, 3)), tex_material((0.8, 0.3, 0.3)))
    add(_grid(lambda u, v: np.stack([np.full_like(u, X), v * Y, u * Z], -1), q(64), q(64), uv_scale=(3, 3)), tex_material((0.3, 0.8, 0.3)))

    # ---- colonnade: two rows of fluted columns
    ncol = 12 if sponza else 6
    col_mats = [tex_material((0.85, 0.8, 0.7), normal=(i == 0)) for i in range(3 if sponza else 1)]
    for row, zc in enumerate((Z * 0.28, Z * 0.72)):
        for i in range(ncol):
            xc = X * (i + 0.5) / ncol
            rad = 38.0 + 6.0 * rng.random()
            def col(u, v, xc=xc, zc=zc, rad=rad):
                ang = u * 2 * np.pi
                r = rad * (1.0 + 0.06 * np.cos(ang * 16)) * (1.0 - 0.15 * v + 0.35 * np.exp(-((v - 0.97) * 25) ** 2) + 0.3 * np.exp(-(v * 25) ** 2))
                return np.stack([xc + r * np.cos(ang), v * (Y * 0.82), zc + r * np.sin(ang)], -1)
            add(_grid(col, q(96) if sponza else q(64), q(112) if sponza else q(64), flip=True, uv_scale=(2, 6)), col_mats[(i + row) % len(col_mats)])
    # ---- beams over the columns
    beam_mat = tex_material((0.7, 0.7, 0.75))
    for zc in (Z * 0.28, Z * 0.72):
        def beam(u, v, zc=zc):
            ang = v * 2 * np.pi
            return np.stack([u * X, Y * 0.86 + 45 * np.sin(ang) + 6 * np.sin(u * 60), zc + 55 * np.cos(ang)], -1)
        add(_grid(beam, q(256), q(48), uv_scale=(16, 1)), beam_mat)
    # ---- curtains (sponza) or furniture blocks (conference)
    if sponza:
        cur_mats = [tex_material(t) for t in ((0.8, 0.2, 0.2), (0.2, 0.3, 0.8), (0.2, 0.7, 0.3), (0.8, 0.7, 0.2))]
        for i in range(8):
            x0 = X * (i + 0.15) / 8
            zc = Z * (0.5 + (0.12 if i % 2 else -0.12))
            ph = rng.random() * 6
            def cur(u, v, x0=x0, zc=zc, ph=ph):
                fold = 28.0 * np.sin(u * 34 + ph) * (0.3 + 0.7 * (1 - v))
                return np.stack([x0 + u * X * 0.09, Y * 0.15 + v * Y * 0.7, zc + fold + 10 * np.sin(v * 9 + u * 3)], -1)
            add(_grid(cur, q(128), q(128), uv_scale=(2, 3)), cur_mats[i % 4])
    else:
        tab_mat = tex_material((0.6, 0.4, 0.25))
        for i in range(24):
            cx, cz = X * (0.1 + 0.8 * rng.random()), Z * (0.35 + 0.3 * rng.random())
            w, d, h = 120 + 60 * rng.random(), 80 + 40 * rng.random(), 70 + 10 * rng.random()
            def top(u, v, cx=cx, cz=cz, w=w, d=d, h=h):
                return np.stack([cx + (u - 0.5) * w, h + 1.5 * np.sin(u * 20) * np.sin(v * 20), cz + (v - 0.5) * d], -1)
            add(_grid(top, q(40), q(40), flip=True), tab_mat)
            def side(u, v, cx=cx, cz=cz, w=w, d=d, h=h):
                ang = u * 2 * np.pi
                return np.stack([cx + 0.5 * w * np.clip(1.5 * np.cos(ang), -1, 1), v * h, cz + 0.5 * d * np.clip(1.5 * np.sin(ang), -1, 1)], -1)
            add(_grid(side, q(64), q(16), flip=True), tab_mat)
    # ---- specular objects: tessellated glass and mirror spheres (caustics)
    materials.append(make_material(capi.MAT_GLASS, Kr=(1, 1, 1), Kt=(1, 1, 1), ior=1.5))
    glass = len(materials) - 1
    materials.append(make_material(capi.MAT_MIRROR, Kr=(0.9, 0.9, 0.95)))
    mirror = len(materials) - 1
    nsph = 6 if sponza else 64
    for i in range(nsph):
        rad = (60.0 + 50.0 * rng.random()) if sponza else (18.0 + 12.0 * rng.random())
        c = np.array([X * (0.08 + 0.84 * rng.random()), rad + (0 if sponza else 72) + 2.0, Z * (0.38 + 0.24 * rng.random())])
        def sph(u, v, c=c, rad=rad):
            th, phi = v * np.pi, u * 2 * np.pi
            return np.stack([c[0] + rad * np.sin(th) * np.cos(phi), c[1] + rad * np.cos(th), c[2] + rad * np.sin(th) * np.sin(phi)], -1)
        add(_grid(sph, q(96) if sponza else q(32), q(64) if sponza else q(20), flip=False), glass if i % 3 else mirror)
    if not sponza:   # mirror panels
        for i in range(4):
            x0 = X * (0.15 + 0.2 * i)
            add(_grid(lambda u, v, x0=x0: np.stack([x0 + u * 300, 200 + v * 500, np.full_like(u, Z - 4.0)], -1), q(8), q(8), flip=True), mirror)

    # ---- lights: a large quad emitter under the ceiling + one point light (scene/Scene.cpp keeps both the light record
    # and the emitter geometry)
    lights = []
    quads = []
    if sponza:
        quads.append(((X * 0.3, Y - 3.0, Z * 0.35), (0.0, 0.0, Z * 0.3), (X * 0.4, 0.0, 0.0), (2.4e7, 2.2e7, 2.0e7)))
    else:
        quads.append(((X * 0.15, Y - 3.0, Z * 0.4), (0.0, 0.0, Z * 0.2), (X * 0.25, 0.0, 0.0), (1.2e7, 1.1e7, 1.0e7)))
        quads.append(((X * 0.60, Y - 3.0, Z * 0.4), (0.0, 0.0, Z * 0.2), (X * 0.25, 0.0, 0.0), (1.2e7, 1.1e7, 1.0e7)))
    for anchor, v1, v2, power in quads:
        L = Light(power, anchor, v1, v2)
        lights.append(L)
        materials.append(make_material(capi.MAT_DIFFUSE_EMITTER, Kd=(1, 1, 1), power=power, inverseArea=L.pod.inverseArea))
        a, e1, e2 = (np.array(t, np.float32) for t in (anchor, v1, v2))
        verts = np.stack([a, a + e1, a + e1 + e2, a + e2]).astype(np.float32)
        meshes.append(dict(vertices=verts, normals=None, uvs=None, tangents=None, bitangents=None,
                           indices=np.array([[0, 1, 2], [0, 2, 3]], np.int32), material=len(materials) - 1))
    if sponza:
        lights.append(Light((6.0e6, 6.0e6, 7.0e6), (X * 0.85, Y * 0.6, Z * 0.5)))

    cam = Camera(eye=(X * 0.04, Y * 0.42, Z * 0.5), lookat=(X * 0.9, Y * 0.35, Z * 0.5), up=(0, 1, 0), hfov=60.0, vfov=60.0,
                 aperture=0.0, aspectRatioMode=Camera.KeepHorizontal)
    pad = 1.0
    scene = SceneDescription("synthetic_" + kind, (-pad, -pad - 4, -pad), (X + pad, Y + pad, Z + pad), materials, lights, meshes=meshes,
                             textures=textures, default_camera=cam)
    return scene
