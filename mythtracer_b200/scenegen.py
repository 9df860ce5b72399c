"""Seeded procedural scenes (OBJ + MTL + PPM textures) standing in for the reference's living-room model.

The reference hard-codes "../Models/Living Room USSU Design.obj" (reference VerStarting/main_local.cc:35),
which is not redistributable/offline, so BASELINE.json's configs are served by synthetic interiors of the
same scale (room ~400 x 120 x 320 units, camera ~(300,57,160), aov 110; cf. main_net_master.cc:255-294).

The files obey the quirks of the reference loader (reference VerStarting/objreader.cc), because the very
same files are fed to the unmodified reference (oracle/_ref) and to the B200 path:
  * every `f` line ends with a space -- the last token is dropped otherwise (objreader.cc:111-115);
  * exactly one triangle per `f` line, so PerPixelDebugInfo::line_no identifies a triangle
    (objreader.cc:180, mythtracer.cc:34);
  * lines stay below 126 characters (char line[128], objreader.cc:234);
  * every face sits under a `usemtl` of a material that exists: a shadow ray that hits a triangle with
    mtl == nullptr dereferences it (mythtracer.cc:121);
  * only map_Ka textures are honoured (objreader.cc:501); they are written as binary PPM, the one format
    oracle/sdl_stub and the product loader both decode.
Coordinates are written with 6 decimals.
"""
from __future__ import annotations

import math
import os
from dataclasses import dataclass, field

import numpy as np

__all__ = ["SceneFiles", "CONFIGS", "generate_scene", "generate_config", "write_ppm", "make_texture"]


@dataclass
class SceneFiles:
    obj_path: str
    mtl_path: str
    n_triangles: int
    camera: tuple  # origin x,y,z, pitch, yaw, roll, aov  (reference camera.h:31-33)
    lights: list   # each 12 floats: position, ambient, diffuse, specular (reference light.h:8-14)
    textures: list = field(default_factory=list)


# ----------------------------------------------------------------------------------------------------
# mesh helpers: every helper returns (verts[N,3], normals[N,3], uvs[N,2], faces[M,3]) with local indices
# ----------------------------------------------------------------------------------------------------

def _grid_patch(origin, du, dv, nu, nv, normal, uv_scale=1.0, jitter=0.0, rng=None):
    """A planar patch origin + s*du + t*dv tessellated nu x nv quads -> 2 triangles each."""
    origin = np.asarray(origin, float)
    du = np.asarray(du, float)
    dv = np.asarray(dv, float)
    s = np.linspace(0.0, 1.0, nu + 1)
    t = np.linspace(0.0, 1.0, nv + 1)
    S, T = np.meshgrid(s, t, indexing="xy")
    if jitter > 0.0 and rng is not None:
        # move interior grid points inside the plane so that edges do not all lie on octree planes
        JS = (rng.random(S.shape) - 0.5) * jitter / nu
        JT = (rng.random(T.shape) - 0.5) * jitter / nv
        JS[:, 0] = JS[:, -1] = 0.0
        JS[0, :] = JS[-1, :] = 0.0
        JT[:, 0] = JT[:, -1] = 0.0
        JT[0, :] = JT[-1, :] = 0.0
        S = S + JS
        T = T + JT
    P = origin[None, None, :] + S[..., None] * du[None, None, :] + T[..., None] * dv[None, None, :]
    verts = P.reshape(-1, 3)
    normals = np.tile(np.asarray(normal, float), (verts.shape[0], 1))
    ulen = np.linalg.norm(du)
    vlen = np.linalg.norm(dv)
    uvs = np.stack([S.reshape(-1) * ulen * uv_scale, T.reshape(-1) * vlen * uv_scale], axis=1)
    idx = np.arange((nu + 1) * (nv + 1)).reshape(nv + 1, nu + 1)
    a = idx[:-1, :-1].reshape(-1)
    b = idx[:-1, 1:].reshape(-1)
    c = idx[1:, 1:].reshape(-1)
    d = idx[1:, :-1].reshape(-1)
    faces = np.concatenate([np.stack([a, b, c], 1), np.stack([c, d, a], 1)], axis=0)
    return verts, normals, uvs, faces


def _uv_sphere(center, radius, nlat, nlon, squash=(1.0, 1.0, 1.0)):
    center = np.asarray(center, float)
    lat = np.linspace(0.0, math.pi, nlat + 1)[1:-1]
    lon = np.linspace(0.0, 2.0 * math.pi, nlon, endpoint=False)
    LA, LO = np.meshgrid(lat, lon, indexing="ij")
    n = np.stack([np.sin(LA) * np.cos(LO), np.cos(LA), np.sin(LA) * np.sin(LO)], axis=-1).reshape(-1, 3)
    ring = n.shape[0]
    n = np.concatenate([n, [[0.0, 1.0, 0.0]], [[0.0, -1.0, 0.0]]], axis=0)
    sq = np.asarray(squash, float)
    verts = center[None, :] + radius * n * sq[None, :]
    nn = n / sq[None, :]
    nn = nn / np.linalg.norm(nn, axis=1, keepdims=True)
    uvs = np.concatenate([
        np.stack([LO.reshape(-1) / (2 * math.pi) * 4.0, 1.0 - LA.reshape(-1) / math.pi * 2.0], 1),
        [[0.5, 1.0]], [[0.5, -1.0]]], axis=0)
    top, bot = ring, ring + 1
    idx = np.arange(ring).reshape(nlat - 1, nlon)
    nxt = np.roll(idx, -1, axis=1)
    faces = []
    # body quads
    a = idx[:-1].reshape(-1)
    b = nxt[:-1].reshape(-1)
    c = nxt[1:].reshape(-1)
    d = idx[1:].reshape(-1)
    faces.append(np.stack([a, b, c], 1))
    faces.append(np.stack([c, d, a], 1))
    # caps
    faces.append(np.stack([np.full(nlon, top), nxt[0], idx[0]], 1))
    faces.append(np.stack([np.full(nlon, bot), idx[-1], nxt[-1]], 1))
    return verts, nn, uvs, np.concatenate(faces, 0)


def _torus(center, R, r, nu, nv, axis=1):
    center = np.asarray(center, float)
    u = np.linspace(0.0, 2 * math.pi, nu, endpoint=False)
    v = np.linspace(0.0, 2 * math.pi, nv, endpoint=False)
    U, V = np.meshgrid(u, v, indexing="ij")
    cx = (R + r * np.cos(V)) * np.cos(U)
    cz = (R + r * np.cos(V)) * np.sin(U)
    cy = r * np.sin(V)
    nx = np.cos(V) * np.cos(U)
    nz = np.cos(V) * np.sin(U)
    ny = np.sin(V)
    P = np.stack([cx, cy, cz], -1).reshape(-1, 3)
    N = np.stack([nx, ny, nz], -1).reshape(-1, 3)
    if axis == 0:
        P = P[:, [1, 0, 2]]
        N = N[:, [1, 0, 2]]
    elif axis == 2:
        P = P[:, [0, 2, 1]]
        N = N[:, [0, 2, 1]]
    verts = center[None, :] + P
    uvs = np.stack([U.reshape(-1) / (2 * math.pi) * 6.0, V.reshape(-1) / (2 * math.pi) * 2.0], 1)
    idx = np.arange(nu * nv).reshape(nu, nv)
    iu = np.roll(idx, -1, axis=0)
    iv = np.roll(idx, -1, axis=1)
    iuv = np.roll(iu, -1, axis=1)
    a, b, c, d = idx.reshape(-1), iu.reshape(-1), iuv.reshape(-1), iv.reshape(-1)
    faces = np.concatenate([np.stack([a, b, c], 1), np.stack([c, d, a], 1)], 0)
    return verts, N, uvs, faces


def _box(lo, hi, n):
    """Axis-aligned box with each face tessellated n x n."""
    lo = np.asarray(lo, float)
    hi = np.asarray(hi, float)
    d = hi - lo
    parts = [
        _grid_patch(lo, [d[0], 0, 0], [0, 0, d[2]], n, n, [0, -1, 0]),
        _grid_patch([lo[0], hi[1], lo[2]], [d[0], 0, 0], [0, 0, d[2]], n, n, [0, 1, 0]),
        _grid_patch(lo, [d[0], 0, 0], [0, d[1], 0], n, n, [0, 0, -1]),
        _grid_patch([lo[0], lo[1], hi[2]], [d[0], 0, 0], [0, d[1], 0], n, n, [0, 0, 1]),
        _grid_patch(lo, [0, 0, d[2]], [0, d[1], 0], n, n, [-1, 0, 0]),
        _grid_patch([hi[0], lo[1], lo[2]], [0, 0, d[2]], [0, d[1], 0], n, n, [1, 0, 0]),
    ]
    return _merge(parts)


def _merge(parts):
    vs, ns, us, fs = [], [], [], []
    off = 0
    for v, n, u, f in parts:
        vs.append(v)
        ns.append(n)
        us.append(u)
        fs.append(f + off)
        off += v.shape[0]
    return np.concatenate(vs), np.concatenate(ns), np.concatenate(us), np.concatenate(fs)


def _tri_soup(rng, n, centers, spread, size):
    """n independent random small triangles around the given centres (flat normals)."""
    c = centers[rng.integers(0, centers.shape[0], n)] + rng.normal(0.0, spread, (n, 3))
    a = c + rng.normal(0.0, size, (n, 3))
    b = c + rng.normal(0.0, size, (n, 3))
    d = c + rng.normal(0.0, size, (n, 3))
    verts = np.stack([a, b, d], 1).reshape(-1, 3)
    nrm = np.cross(b - a, d - a)
    ln = np.linalg.norm(nrm, axis=1, keepdims=True)
    ln[ln == 0.0] = 1.0
    nrm = np.repeat(nrm / ln, 3, axis=0)
    uvs = np.tile(np.array([[0.0, 0.0], [1.0, 0.0], [0.0, 1.0]]), (n, 1))
    faces = np.arange(3 * n).reshape(n, 3)
    return verts, nrm, uvs, faces


def _sticks(rng, n, lo, hi, width):
    """Long thin triangles spanning the room: they straddle the top octree planes (stress case)."""
    lo = np.asarray(lo, float)
    hi = np.asarray(hi, float)
    a = lo + rng.random((n, 3)) * (hi - lo)
    b = lo + rng.random((n, 3)) * (hi - lo)
    w = rng.normal(0.0, width, (n, 3))
    c = a + w
    verts = np.stack([a, b, c], 1).reshape(-1, 3)
    nrm = np.cross(b - a, c - a)
    ln = np.linalg.norm(nrm, axis=1, keepdims=True)
    ln[ln == 0.0] = 1.0
    nrm = np.repeat(nrm / ln, 3, axis=0)
    uvs = np.tile(np.array([[0.0, 0.0], [1.0, 0.0], [0.0, 1.0]]), (n, 1))
    faces = np.arange(3 * n).reshape(n, 3)
    return verts, nrm, uvs, faces


# ----------------------------------------------------------------------------------------------------
# textures
# ----------------------------------------------------------------------------------------------------

def make_texture(kind: str, size: int, seed: int) -> np.ndarray:
    rng = np.random.default_rng(seed)
    y, x = np.mgrid[0:size, 0:size]
    if kind == "checker":
        cells = ((x // (size // 8)) + (y // (size // 8))) % 2
        base = np.where(cells[..., None] == 0, np.array([235, 225, 200]), np.array([70, 45, 30]))
        noise = rng.integers(-12, 13, (size, size, 1))
        img = np.clip(base + noise, 0, 255)
    elif kind == "wood":
        r = np.sqrt((x - size * 0.3) ** 2 + (y * 0.25 - size * 0.1) ** 2)
        rings = 0.5 + 0.5 * np.sin(r * 0.35 + rng.normal(0, 0.4, (size, size)))
        img = np.stack([120 + 90 * rings, 70 + 60 * rings, 30 + 30 * rings], -1)
    else:  # "noise"
        coarse = rng.random((size // 16 + 1, size // 16 + 1, 3))
        img = np.kron(coarse, np.ones((16, 16, 1)))[:size, :size] * 200 + 40
        img = img + rng.integers(-10, 11, (size, size, 3))
    return np.clip(img, 0, 255).astype(np.uint8)


def write_ppm(path: str, img: np.ndarray) -> None:
    h, w, _ = img.shape
    with open(path, "wb") as f:
        f.write(b"P6\n%d %d\n255\n" % (w, h))
        f.write(np.ascontiguousarray(img, dtype=np.uint8).tobytes())


# ----------------------------------------------------------------------------------------------------
# materials (MTL keys the reference reads: Ka Kd Ks Ns Ni Tr Tf Refl map_Ka; objreader.cc:487-503)
# ----------------------------------------------------------------------------------------------------

def _mtl(name, ka, kd, ks, ns, refl=0.0, tr=0.0, tf=(1.0, 1.0, 1.0), ni=1.0, tex=None):
    return dict(name=name, ka=ka, kd=kd, ks=ks, ns=ns, refl=refl, tr=tr, tf=tf, ni=ni, tex=tex)


def _write_mtl(path, materials):
    with open(path, "w") as f:
        f.write("# synthetic materials (mythtracer_b200.scenegen)\n")
        for m in materials:
            f.write("newmtl %s\n" % m["name"])
            f.write("Ka %.6f %.6f %.6f\n" % tuple(m["ka"]))
            f.write("Kd %.6f %.6f %.6f\n" % tuple(m["kd"]))
            f.write("Ks %.6f %.6f %.6f\n" % tuple(m["ks"]))
            f.write("Ns %.6f\n" % m["ns"])
            f.write("Ni %.6f\n" % m["ni"])
            f.write("Tr %.6f\n" % m["tr"])
            f.write("Tf %.6f %.6f %.6f\n" % tuple(m["tf"]))
            f.write("Refl %.6f\n" % m["refl"])
            if m["tex"]:
                f.write("map_Ka %s\n" % m["tex"])
            f.write("\n")


def _format_rows(fmt, arr):
    # fast "%.6f" formatting of big arrays
    return "".join([fmt % tuple(r) for r in arr.tolist()])


def _write_obj(path, mtl_name, groups):
    """groups: list of (material_name, (verts, normals, uvs, faces), use_uv)."""
    with open(path, "w") as f:
        f.write("# synthetic scene (mythtracer_b200.scenegen)\n")
        f.write("mtllib %s\n" % mtl_name)
        voff = 0
        toff = 0
        n_tris = 0
        for mat, (v, n, u, faces), use_uv in groups:
            f.write("o part_%s_%d\n" % (mat, voff))
            f.write(_format_rows("v %.6f %.6f %.6f\n", v))
            f.write(_format_rows("vn %.6f %.6f %.6f\n", n))
            if use_uv:
                f.write(_format_rows("vt %.6f %.6f\n", u))
            f.write("usemtl %s\n" % mat)
            fi = faces + voff + 1
            if use_uv:
                ft = faces + toff + 1   # vt entries exist only for textured parts: separate running index
                rows = np.stack([fi[:, 0], ft[:, 0], fi[:, 0], fi[:, 1], ft[:, 1], fi[:, 1], fi[:, 2], ft[:, 2], fi[:, 2]], 1)
                f.write(_format_rows("f %d/%d/%d %d/%d/%d %d/%d/%d \n", rows))
                toff += v.shape[0]
            else:
                rows = np.repeat(fi, 2, axis=1)
                f.write(_format_rows("f %d//%d %d//%d %d//%d \n", rows))
            voff += v.shape[0]
            n_tris += faces.shape[0]
    return n_tris


# ----------------------------------------------------------------------------------------------------
# the room
# ----------------------------------------------------------------------------------------------------

ROOM = (400.0, 120.0, 320.0)


def generate_scene(out_dir: str, target_tris: int, seed: int, textured: bool = False, stress: bool = False,
                   name: str = "scene") -> SceneFiles:
    """Writes <out_dir>/<name>.obj/.mtl (+ textures) with roughly `target_tris` triangles."""
    os.makedirs(out_dir, exist_ok=True)
    rng = np.random.default_rng(seed)
    W, H, D = ROOM
    # irrational-ish offsets keep geometry off the octree's binary lattice
    ox, oy, oz = 0.0137, 0.0071, 0.0093

    tex_files = []
    if textured:
        for i, kind in enumerate(["checker", "wood", "noise", "checker"]):
            fname = "%s_tex%d.ppm" % (name, i)
            write_ppm(os.path.join(out_dir, fname), make_texture(kind, 512 if target_tris > 20000 else 64, seed * 100 + i))
            tex_files.append(fname)

    def tex(i):
        return tex_files[i % len(tex_files)] if textured else None

    materials = [
        _mtl("floor", (0.55, 0.5, 0.45), (0.6, 0.55, 0.5), (0.25, 0.25, 0.25), 40.0, refl=0.18, tex=tex(0)),
        _mtl("wall", (0.62, 0.64, 0.6), (0.55, 0.55, 0.52), (0.05, 0.05, 0.05), 8.0, tex=tex(2)),
        _mtl("ceiling", (0.7, 0.7, 0.72), (0.4, 0.4, 0.4), (0.0, 0.0, 0.0), 1.0),
        _mtl("mirror", (0.05, 0.05, 0.06), (0.1, 0.1, 0.1), (0.9, 0.9, 0.9), 180.0, refl=0.85),
        _mtl("glass", (0.05, 0.07, 0.06), (0.08, 0.1, 0.09), (0.8, 0.8, 0.8), 120.0, refl=0.12, tr=0.8,
             tf=(0.85, 0.97, 0.9), ni=1.5),
        _mtl("amber", (0.12, 0.07, 0.02), (0.2, 0.12, 0.03), (0.6, 0.6, 0.5), 90.0, refl=0.05, tr=0.6,
             tf=(1.0, 0.75, 0.35), ni=1.33),
        _mtl("chrome", (0.15, 0.15, 0.17), (0.25, 0.25, 0.28), (0.8, 0.8, 0.85), 200.0, refl=0.6),
        _mtl("red", (0.6, 0.12, 0.1), (0.7, 0.15, 0.12), (0.5, 0.5, 0.5), 60.0, refl=0.08),
        _mtl("green", (0.12, 0.5, 0.18), (0.15, 0.6, 0.2), (0.3, 0.3, 0.3), 25.0),
        _mtl("blue", (0.1, 0.18, 0.6), (0.12, 0.2, 0.7), (0.6, 0.6, 0.6), 110.0, refl=0.1),
        _mtl("fabric", (0.5, 0.42, 0.3), (0.55, 0.45, 0.33), (0.02, 0.02, 0.02), 4.0, tex=tex(1)),
        _mtl("clutter", (0.45, 0.4, 0.5), (0.5, 0.45, 0.55), (0.2, 0.2, 0.2), 30.0),
    ]

    budget = float(target_tris)
    if stress:
        n_clutter = int(0.45 * budget)
        n_sticks = min(20000, int(0.02 * budget))
        budget -= n_clutter + n_sticks
    wall_budget = 0.2 * budget
    obj_budget = 0.7 * budget
    box_budget = 0.1 * budget

    groups = []
    use_uv = textured

    # --- shell: floor, ceiling, 4 walls (6 patches share wall_budget by area) ---
    patches = [
        ("floor", (ox, oy, oz), (W, 0, 0), (0, 0, D), (0, 1, 0), W * D),
        ("ceiling", (ox, H + oy, oz), (W, 0, 0), (0, 0, D), (0, -1, 0), W * D),
        ("wall", (ox, oy, oz), (W, 0, 0), (0, H, 0), (0, 0, 1), W * H),
        ("wall", (ox, oy, D + oz), (W, 0, 0), (0, H, 0), (0, 0, -1), W * H),
        ("wall", (ox, oy, oz), (0, 0, D), (0, H, 0), (1, 0, 0), D * H),
        ("wall", (W + ox, oy, oz), (0, 0, D), (0, H, 0), (-1, 0, 0), D * H),
    ]
    area_total = sum(p[5] for p in patches)
    for mat, o, du, dv, nrm, area in patches:
        quads = max(1.0, wall_budget * area / area_total / 2.0)
        lu = np.linalg.norm(du)
        lv = np.linalg.norm(dv)
        nu = max(1, int(round(math.sqrt(quads * lu / lv))))
        nv = max(1, int(round(quads / nu)))
        mesh = _grid_patch(o, du, dv, nu, nv, nrm, uv_scale=1.0 / 80.0, jitter=0.6, rng=rng)
        groups.append((mat, mesh, use_uv and mat in ("floor", "wall")))

    # --- a wall mirror and a glass pane (few, large triangles: long-range secondary rays) ---
    groups.append(("mirror", _grid_patch((120.0 + ox, 25.0, 1.5 + oz), (160, 0, 0), (0, 70, 0), 2, 2, (0, 0, 1)), False))
    groups.append(("glass", _grid_patch((150.0 + ox, 5.0, 118.7 + oz), (0, 0, 90), (0, 75, 0), 3, 2, (1, 0, 0)), False))

    # --- curved objects: spheres and tori share obj_budget ---
    objs = [
        ("sphere", "mirror", (110.3, 38.0, 95.7), 30.0, (1, 1, 1)),
        ("sphere", "glass", (215.6, 33.0, 150.2), 26.0, (1, 1, 1)),
        ("sphere", "red", (300.4, 24.0, 70.9), 22.0, (1, 1, 1)),
        ("sphere", "blue", (60.2, 20.5, 200.3), 19.0, (1.3, 0.8, 1.0)),
        ("sphere", "amber", (255.8, 58.0, 235.1), 17.0, (1, 1, 1)),
        ("sphere", "chrome", (180.9, 84.0, 60.4), 15.0, (1, 1, 1)),
        ("sphere", "fabric", (340.1, 30.0, 250.6), 28.0, (1.2, 0.9, 1.1)),
        ("torus", "chrome", (200.2, 12.5, 230.8), 34.0, 9.0, 1),
        ("torus", "green", (90.7, 60.0, 285.4), 26.0, 7.0, 2),
        ("torus", "red", (330.5, 75.0, 140.3), 20.0, 6.0, 0),
        ("sphere", "green", (150.6, 15.0, 40.2), 13.0, (1, 1, 1)),
        ("sphere", "glass", (275.3, 16.0, 185.7), 14.0, (1, 1.1, 1)),
    ]
    weights = np.array([o[3] ** 2 if o[0] == "sphere" else o[3] * o[4] * 2.0 for o in objs], float)
    weights = weights / weights.sum()
    for o, wgt in zip(objs, weights):
        tris = max(16.0, obj_budget * wgt)
        if o[0] == "sphere":
            nlon = max(5, int(round(math.sqrt(tris))))
            nlat = max(3, int(round(tris / (2.0 * nlon))) + 1)
            mesh = _uv_sphere(o[2], o[3], nlat, nlon, o[4])
        else:
            nu = max(5, int(round(math.sqrt(tris / 2.0 * o[3] / o[4]))))
            nv = max(4, int(round(tris / 2.0 / nu)))
            mesh = _torus(o[2], o[3], o[4], nu, nv, axis=o[5])
        groups.append((o[1], mesh, use_uv and o[1] == "fabric"))

    # --- furniture boxes ---
    boxes = [
        ("fabric", (40.0, 0.4, 30.0), (130.0, 28.0, 75.0)),
        ("green", (310.0, 0.4, 180.0), (370.0, 45.0, 300.0)),
        ("blue", (230.0, 0.4, 20.0), (280.0, 18.0, 60.0)),
        ("red", (20.0, 0.4, 230.0), (45.0, 90.0, 300.0)),
    ]
    per_box = box_budget / len(boxes)
    for mat, lo, hi in boxes:
        n = max(1, int(round(math.sqrt(per_box / 12.0))))
        groups.append((mat, _box(np.array(lo) + (ox, oy, oz), np.array(hi) + (ox, oy, oz), n), use_uv and mat == "fabric"))

    if stress:
        centers = np.array([[70.0, 70.0, 60.0], [330.0, 90.0, 60.0], [200.0, 60.0, 290.0], [120.0, 95.0, 170.0],
                            [290.0, 20.0, 120.0], [45.0, 30.0, 120.0]]) + (ox, oy, oz)
        groups.append(("clutter", _tri_soup(rng, n_clutter, centers, 7.0, 0.22), False))
        groups.append(("clutter", _sticks(rng, n_sticks, (5.0, 5.0, 5.0), (W - 5.0, H - 5.0, D - 5.0), 0.35), False))

    obj_path = os.path.join(out_dir, name + ".obj")
    mtl_path = os.path.join(out_dir, name + ".mtl")
    _write_mtl(mtl_path, materials)
    n_tris = _write_obj(obj_path, name + ".mtl", groups)

    camera = (301.37, 57.21, 161.13, 4.0, 243.0, 0.0, 110.0)
    rig = [
        (231.82174, 81.69966, 27.78259, 0.3, 0.3, 0.3, 1.0, 1.0, 1.0, 1.0, 1.0, 1.0),
        (200.0, 95.0, 160.0, 0.0, 0.0, 0.0, 0.3, 0.3, 0.3, 0.3, 0.3, 0.3),
        (120.0, 100.0, 250.0, 0.0, 0.0, 0.0, 0.3, 0.3, 0.3, 0.3, 0.3, 0.3),
        (330.0, 100.0, 80.0, 0.0, 0.0, 0.0, 0.3, 0.3, 0.3, 0.3, 0.3, 0.3),
    ]
    return SceneFiles(obj_path, mtl_path, n_tris, camera, rig, [os.path.join(out_dir, t) for t in tex_files])


# BASELINE.json configs (SURVEY.md section 8 shorthand C1..C5)
CONFIGS = {
    "C1": dict(target_tris=2000, seed=1, textured=False, stress=False, width=320, height=240, n_lights=1, depth=2),
    "C2": dict(target_tris=100000, seed=2, textured=True, stress=False, width=1280, height=720, n_lights=2, depth=3),
    "C3": dict(target_tris=500000, seed=3, textured=False, stress=False, width=1920, height=1080, n_lights=2, depth=5),
    "C4": dict(target_tris=2000000, seed=4, textured=False, stress=True, width=1920, height=1080, n_lights=2, depth=5),
    "C5": dict(target_tris=500000, seed=3, textured=False, stress=False, width=3840, height=2160, n_lights=4, depth=8),
}


def generate_config(name: str, out_dir: str, scale: float = 1.0):
    """Generates (or reuses) the scene of a BASELINE config; returns (SceneFiles, config dict)."""
    cfg = dict(CONFIGS[name])
    tris = max(200, int(cfg["target_tris"] * scale))
    scene_name = "%s_%d_s%d" % (name.lower() if name != "C5" else "c3", tris, cfg["seed"])
    stamp = os.path.join(out_dir, scene_name + ".done")
    files = None
    if os.path.exists(stamp):
        with open(stamp) as f:
            n_tris = int(f.read().strip())
        tex_files = []
        if cfg["textured"]:
            tex_files = [os.path.join(out_dir, "%s_tex%d.ppm" % (scene_name, i)) for i in range(4)]
        cam = (301.37, 57.21, 161.13, 4.0, 243.0, 0.0, 110.0)
        files = SceneFiles(os.path.join(out_dir, scene_name + ".obj"), os.path.join(out_dir, scene_name + ".mtl"),
                           n_tris, cam, _default_rig(), tex_files)
    else:
        files = generate_scene(out_dir, tris, cfg["seed"], cfg["textured"], cfg["stress"], name=scene_name)
        with open(stamp, "w") as f:
            f.write(str(files.n_triangles))
    files.lights = files.lights[: cfg["n_lights"]]
    return files, cfg


def _default_rig():
    return [
        (231.82174, 81.69966, 27.78259, 0.3, 0.3, 0.3, 1.0, 1.0, 1.0, 1.0, 1.0, 1.0),
        (200.0, 95.0, 160.0, 0.0, 0.0, 0.0, 0.3, 0.3, 0.3, 0.3, 0.3, 0.3),
        (120.0, 100.0, 250.0, 0.0, 0.0, 0.0, 0.3, 0.3, 0.3, 0.3, 0.3, 0.3),
        (330.0, 100.0, 80.0, 0.0, 0.0, 0.0, 0.3, 0.3, 0.3, 0.3, 0.3, 0.3),
    ]
