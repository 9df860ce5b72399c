"""TEST INFRASTRUCTURE ONLY: ctypes bindings of the parity oracle (oracle/mt_oracle.cc) and of the unmodified
reference build (oracle/_ref/libmythtracer_ref.so), plus an independent Python OBJ/MTL/PPM reader.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / `--impl reference` legs import this
module.  The product package (mythtracer_b200/) never does.
"""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ORACLE_SO = os.path.join(HERE, "_build", "libmt_oracle.so")
REF_SO = os.path.join(HERE, "_ref", "libmythtracer_ref.so")

TRI_DTYPE = np.dtype([("vertex", "f8", (9,)), ("normal", "f8", (9,)), ("uvw", "f8", (9,)),
                      ("material", "i4"), ("line_no", "i4")], align=True)
MTL_DTYPE = np.dtype([("ambient", "f8", (3,)), ("diffuse", "f8", (3,)), ("specular", "f8", (3,)),
                      ("specular_exp", "f8"), ("reflectance", "f8"), ("transparency", "f8"),
                      ("transmission_filter", "f8", (3,)), ("refraction_index", "f8"),
                      ("texture", "i4"), ("pad_", "i4")], align=True)
LIGHT_DTYPE = np.dtype([("position", "f8", (3,)), ("ambient", "f8", (3,)), ("diffuse", "f8", (3,)),
                        ("specular", "f8", (3,))], align=True)
CAMERA_DTYPE = np.dtype([("origin", "f8", (3,)), ("pitch", "f8"), ("yaw", "f8"), ("roll", "f8"), ("aov", "f8")],
                        align=True)
STATS_FIELDS = ["rays", "primary", "shadow", "reflect", "refract", "n_slab", "n_visit", "n_triaabb", "n_mt",
                "n_hit", "n_shade"]
assert TRI_DTYPE.itemsize == 224 and MTL_DTYPE.itemsize == 136 and LIGHT_DTYPE.itemsize == 96


class TextureStruct(ctypes.Structure):
    _fields_ = [("width", ctypes.c_int32), ("height", ctypes.c_int32), ("rgba", ctypes.c_void_p)]


class TapsStruct(ctypes.Structure):
    _fields_ = [("sig_hits", ctypes.c_void_p), ("sig_shadow", ctypes.c_void_p), ("n_rays", ctypes.c_void_p)]


def build(force: bool = False) -> None:
    """Compiles the restatement and (when /root/reference is present) the reference library."""
    if force or not os.path.exists(ORACLE_SO) or os.path.getmtime(ORACLE_SO) < os.path.getmtime(
            os.path.join(HERE, "mt_oracle.cc")):
        subprocess.check_call(["make", "-C", HERE, "_build/libmt_oracle.so"], stdout=subprocess.DEVNULL)
    if os.path.isdir(os.environ.get("MTB_REFERENCE_DIR", "/root/reference/VerStarting")) or os.path.exists(REF_SO):
        subprocess.check_call([os.path.join(HERE, "build_ref.sh")], stdout=subprocess.DEVNULL)


def _ptr(a):
    return None if a is None else a.ctypes.data_as(ctypes.c_void_p)


# ----------------------------------------------------------------------------------------------------
# independent OBJ / MTL / PPM reader (semantics of reference objreader.cc for the files scenegen writes and
# for the quirk tests: 128-byte fgets buffer, 0-based line numbers, %i face tokens, lost last token)
# ----------------------------------------------------------------------------------------------------

def read_ppm(path: str) -> np.ndarray:
    with open(path, "rb") as f:
        data = f.read()
    assert data[:2] == b"P6", path
    vals = []
    pos = 2
    while len(vals) < 3:
        while data[pos:pos + 1].isspace():
            pos += 1
        if data[pos:pos + 1] == b"#":
            while data[pos:pos + 1] != b"\n":
                pos += 1
            continue
        end = pos
        while not data[end:end + 1].isspace():
            end += 1
        vals.append(int(data[pos:end]))
        pos = end
    pos += 1
    w, h, maxval = vals
    assert maxval == 255
    rgb = np.frombuffer(data, np.uint8, w * h * 3, pos).reshape(h, w, 3)
    rgba = np.full((h, w, 4), 255, np.uint8)
    rgba[..., :3] = rgb
    return rgba


def _split_lines_128(path):
    """fgets(line, 128) semantics: longer lines are returned in 127-byte pieces, each counted as a line."""
    with open(path, "rb") as f:
        raw = f.read()
    out = []
    for ln in raw.split(b"\n"):
        ln_nl = ln + b"\n"
        while len(ln_nl) > 127:
            out.append(ln_nl[:127])
            ln_nl = ln_nl[127:]
        out.append(ln_nl)
    if out and out[-1] == b"\n" and raw.endswith(b"\n"):
        out.pop()
    return [x.decode("latin-1") for x in out]


def read_mtl(path: str):
    names, mats, tex_names = [], [], []
    cur = None
    base = os.path.dirname(path)
    for line in _split_lines_128(path):
        line = line.replace("\r", "").replace("\n", "")
        tok = line.split()
        if not tok or tok[0].startswith("#"):
            continue
        key = tok[0]
        if key == "newmtl":
            cur = np.zeros((), MTL_DTYPE)
            cur["texture"] = -1
            names.append(tok[1])
            mats.append(cur)
        elif key in ("Ka", "Kd", "Ks", "Tf"):
            field = {"Ka": "ambient", "Kd": "diffuse", "Ks": "specular", "Tf": "transmission_filter"}[key]
            cur[field] = [float(tok[1]), float(tok[2]), float(tok[3])]
        elif key in ("Ns", "Ni", "Tr", "Refl"):
            field = {"Ns": "specular_exp", "Ni": "refraction_index", "Tr": "transparency", "Refl": "reflectance"}[key]
            cur[field] = float(tok[1])
        elif key == "map_Ka":
            fname = line.split("map_Ka", 1)[1].strip()
            full = os.path.join(base, fname) if base else fname
            if full not in tex_names:
                tex_names.append(full)
            cur["texture"] = tex_names.index(full)
    # later definitions of the same name replace earlier ones (materials[mtl_name] = ..., objreader.cc:280)
    table = {}
    for n, m in zip(names, mats):
        table[n] = m
    uniq = list(table.keys())
    arr = np.zeros(len(uniq), MTL_DTYPE)
    for i, n in enumerate(uniq):
        arr[i] = table[n]
    return uniq, arr, tex_names


def read_obj(path: str):
    """Returns (triangles[TRI_DTYPE], materials[MTL_DTYPE], textures[list of rgba arrays])."""
    verts, norms, uvs = [], [], []
    tris = []
    mtl_names, mtl_arr, tex_files = [], np.zeros(0, MTL_DTYPE), []
    cur_mtl = -1
    base = os.path.dirname(path)
    for line_no, line in enumerate(_split_lines_128(path)):
        line = line.replace("\r", "").replace("\n", "")
        tok = line.split()
        if not tok or tok[0].startswith("#"):
            continue
        key = tok[0]
        if key == "v":
            verts.append((float(tok[1]), float(tok[2]), float(tok[3])))
        elif key == "vn":
            norms.append((float(tok[1]), float(tok[2]), float(tok[3])))
        elif key == "vt":
            uvs.append((float(tok[1]), float(tok[2]), float(tok[3]) if len(tok) > 3 else 0.0))
        elif key == "mtllib":
            fname = line.split("mtllib", 1)[1].strip()
            mtl_names, mtl_arr, tex_files = read_mtl(os.path.join(base, fname) if base else fname)
        elif key == "usemtl":
            cur_mtl = mtl_names.index(tok[1]) if tok[1] in mtl_names else -1
        elif key == "f":
            toks = tok[1:]
            if not line.endswith((" ", "\t")):
                toks = toks[:-1]  # objreader.cc:111-115: the last token is lost without trailing space
            vi, ti, ni = [], [], []
            for t in toks:
                parts = t.split("/")
                vi.append(int(parts[0], 0) - 1)
                ti.append(int(parts[1], 0) - 1 if len(parts) > 1 and parts[1] else -1)
                ni.append(int(parts[2], 0) - 1 if len(parts) > 2 and parts[2] else -1)
            assert len(vi) in (3, 4), "unsupported face count"
            if len(vi) == 4:
                vi.append(vi[0]); ti.append(ti[0]); ni.append(ni[0])
            for i in range(3, len(vi) + 1, 2):
                tris.append((vi[i - 3:i], ti[i - 3:i], ni[i - 3:i], cur_mtl, line_no))
    V = np.array(verts, np.float64).reshape(-1, 3)
    N = np.array(norms, np.float64).reshape(-1, 3)
    T = np.array(uvs, np.float64).reshape(-1, 3)
    out = np.zeros(len(tris), TRI_DTYPE)
    if tris:
        vi = np.array([t[0] for t in tris])
        ti = np.array([t[1] for t in tris])
        ni = np.array([t[2] for t in tris])
        out["vertex"] = V[vi].reshape(-1, 9)
        has_n = (ni != -1).all(axis=1)
        if has_n.any():
            out["normal"][has_n] = N[ni[has_n]].reshape(-1, 9)
        has_t = (ti != -1).all(axis=1)
        if has_t.any():
            out["uvw"][has_t] = T[ti[has_t]].reshape(-1, 9)
        out["material"] = [t[3] for t in tris]
        out["line_no"] = [t[4] for t in tris]
    textures = [read_ppm(p) for p in tex_files]
    return out, mtl_arr, textures


def make_lights(rows) -> np.ndarray:
    arr = np.zeros(len(rows), LIGHT_DTYPE)
    for i, r in enumerate(rows):
        arr[i]["position"] = r[0:3]
        arr[i]["ambient"] = r[3:6]
        arr[i]["diffuse"] = r[6:9]
        arr[i]["specular"] = r[9:12]
    return arr


def make_camera(cam) -> np.ndarray:
    c = np.zeros((), CAMERA_DTYPE)
    c["origin"] = cam[0:3]
    c["pitch"], c["yaw"], c["roll"], c["aov"] = cam[3], cam[4], cam[5], cam[6]
    return c


# ----------------------------------------------------------------------------------------------------
# the restatement
# ----------------------------------------------------------------------------------------------------

class Oracle:
    """CPU restatement (oracle/mt_oracle.cc) over explicit scene arrays."""

    _lib = None

    @classmethod
    def lib(cls):
        if cls._lib is None:
            build()
            lib = ctypes.CDLL(ORACLE_SO)
            lib.mto_create.restype = ctypes.c_void_p
            lib.mto_create.argtypes = [ctypes.c_void_p, ctypes.c_int64, ctypes.c_void_p, ctypes.c_int32,
                                       ctypes.c_void_p, ctypes.c_int32]
            lib.mto_destroy.argtypes = [ctypes.c_void_p]
            lib.mto_set_lights.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int32]
            lib.mto_tree_info.argtypes = [ctypes.c_void_p, ctypes.c_void_p]
            lib.mto_scene_aabb.argtypes = [ctypes.c_void_p, ctypes.c_void_p]
            lib.mto_triangle_nodes.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]
            lib.mto_render.argtypes = [ctypes.c_void_p, ctypes.c_void_p] + [ctypes.c_int] * 7 + [ctypes.c_void_p] * 5
            lib.mto_render_color.argtypes = [ctypes.c_void_p, ctypes.c_void_p] + [ctypes.c_int] * 7 + [ctypes.c_void_p]
            lib.mto_intersect.argtypes = [ctypes.c_void_p, ctypes.c_int64] + [ctypes.c_void_p] * 6
            lib.mto_intersect_brute.argtypes = [ctypes.c_void_p, ctypes.c_int64] + [ctypes.c_void_p] * 4
            lib.mto_camera_ray.argtypes = [ctypes.c_void_p] + [ctypes.c_int] * 4 + [ctypes.c_void_p]
            lib.mto_camera_sensor.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_void_p]
            lib.mto_texture_sample.argtypes = [ctypes.c_void_p, ctypes.c_int32, ctypes.c_double, ctypes.c_double,
                                               ctypes.c_void_p]
            lib.mto_triangle_normal.argtypes = [ctypes.c_void_p, ctypes.c_int64, ctypes.c_void_p, ctypes.c_void_p]
            lib.mto_triangle_uvw.argtypes = [ctypes.c_void_p, ctypes.c_int64, ctypes.c_void_p, ctypes.c_void_p]
            lib.mto_quantize.argtypes = [ctypes.c_void_p, ctypes.c_void_p]
            lib.mto_math3d.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]
            lib.mto_mix64.restype = ctypes.c_uint64
            lib.mto_mix64.argtypes = [ctypes.c_uint64] * 3
            lib.mto_set_threads.argtypes = [ctypes.c_int32]
            lib.mto_get_threads.restype = ctypes.c_int32
            cls._lib = lib
        return cls._lib

    def __init__(self, tris, mtls, textures=()):
        lib = self.lib()
        self.tris = np.ascontiguousarray(tris, TRI_DTYPE)
        self.mtls = np.ascontiguousarray(mtls, MTL_DTYPE)
        self._tex_keep = [np.ascontiguousarray(t, np.uint8) for t in textures]
        tex_arr = (TextureStruct * max(1, len(self._tex_keep)))()
        for i, t in enumerate(self._tex_keep):
            tex_arr[i].width, tex_arr[i].height = t.shape[1], t.shape[0]
            tex_arr[i].rgba = t.ctypes.data
        self.handle = lib.mto_create(_ptr(self.tris), len(self.tris), _ptr(self.mtls), len(self.mtls),
                                     ctypes.cast(tex_arr, ctypes.c_void_p), len(self._tex_keep))
        self.n_lights = 0

    @classmethod
    def from_obj(cls, path):
        tris, mtls, texs = read_obj(path)
        return cls(tris, mtls, texs)

    def close(self):
        if self.handle:
            self.lib().mto_destroy(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_lights(self, rows):
        arr = make_lights(rows)
        self.lib().mto_set_lights(self.handle, _ptr(arr), len(arr))
        self.n_lights = len(arr)

    def set_threads(self, n):
        self.lib().mto_set_threads(n)

    def threads(self):
        return self.lib().mto_get_threads()

    def tree_info(self):
        out = np.zeros(5, np.int64)
        self.lib().mto_tree_info(self.handle, _ptr(out))
        return dict(nodes=int(out[0]), depth=int(out[1]), biggest_list=int(out[2]), root_list=int(out[3]),
                    interior_tris=int(out[4]))

    def aabb(self):
        out = np.zeros(6)
        self.lib().mto_scene_aabb(self.handle, _ptr(out))
        return out

    def triangle_nodes(self):
        box = np.zeros((len(self.tris), 6))
        depth = np.zeros(len(self.tris), np.int32)
        self.lib().mto_triangle_nodes(self.handle, _ptr(box), _ptr(depth))
        return box, depth

    def render(self, cam, image_w, image_h, chunk=None, depth=5, taps=False, debug=True):
        cx, cy, cw, ch = chunk if chunk is not None else (0, 0, image_w, image_h)
        c = make_camera(cam)
        rgb = np.zeros((ch, cw, 3), np.uint8)
        line_no = np.zeros((ch, cw), np.int32) if debug else None
        points = np.zeros((ch, cw, 3), np.float64) if debug else None
        stats = np.zeros(len(STATS_FIELDS), np.uint64)
        res = dict()
        tap_struct = None
        if taps:
            res["sig_hits"] = np.zeros((ch, cw), np.uint64)
            res["sig_shadow"] = np.zeros((ch, cw), np.uint64)
            res["n_rays"] = np.zeros((ch, cw), np.uint32)
            tap_struct = TapsStruct(res["sig_hits"].ctypes.data, res["sig_shadow"].ctypes.data,
                                    res["n_rays"].ctypes.data)
        rc = self.lib().mto_render(self.handle, _ptr(c), image_w, image_h, cx, cy, cw, ch, depth, _ptr(rgb),
                                   _ptr(line_no), _ptr(points),
                                   ctypes.cast(ctypes.pointer(tap_struct), ctypes.c_void_p) if taps else None,
                                   _ptr(stats))
        assert rc == 0
        res.update(rgb=rgb, line_no=line_no, points=points, stats=dict(zip(STATS_FIELDS, [int(x) for x in stats])))
        return res

    def render_color(self, cam, image_w, image_h, chunk=None, depth=5):
        cx, cy, cw, ch = chunk if chunk is not None else (0, 0, image_w, image_h)
        c = make_camera(cam)
        out = np.zeros((ch, cw, 3), np.float64)
        rc = self.lib().mto_render_color(self.handle, _ptr(c), image_w, image_h, cx, cy, cw, ch, depth, _ptr(out))
        assert rc == 0
        return out

    def intersect(self, origins, dirs, brute=False):
        o = np.ascontiguousarray(origins, np.float64).reshape(-1, 3)
        d = np.ascontiguousarray(dirs, np.float64).reshape(-1, 3)
        n = o.shape[0]
        tri = np.full(n, -1, np.int32)
        t = np.zeros(n)
        p = np.zeros((n, 3))
        stats = np.zeros(len(STATS_FIELDS), np.uint64)
        if brute:
            self.lib().mto_intersect_brute(self.handle, n, _ptr(o), _ptr(d), _ptr(tri), _ptr(t))
        else:
            self.lib().mto_intersect(self.handle, n, _ptr(o), _ptr(d), _ptr(tri), _ptr(t), _ptr(p), _ptr(stats))
        return dict(tri=tri, t=t, point=p, stats=dict(zip(STATS_FIELDS, [int(x) for x in stats])))

    @classmethod
    def camera_sensor(cls, cam, w, h):
        out = np.zeros(9)
        cls.lib().mto_camera_sensor(_ptr(make_camera(cam)), w, h, _ptr(out))
        return out.reshape(3, 3)

    def camera_ray(self, cam, w, h, x, y):
        out = np.zeros(3)
        self.lib().mto_camera_ray(_ptr(make_camera(cam)), w, h, x, y, _ptr(out))
        return out

    def texture_sample(self, tex, u, v):
        out = np.zeros(3)
        self.lib().mto_texture_sample(self.handle, tex, u, v, _ptr(out))
        return out

    def triangle_normal(self, tri, point):
        out = np.zeros(3)
        p = np.ascontiguousarray(point, np.float64)
        self.lib().mto_triangle_normal(self.handle, tri, _ptr(p), _ptr(out))
        return out

    def triangle_uvw(self, tri, point):
        out = np.zeros(3)
        p = np.ascontiguousarray(point, np.float64)
        self.lib().mto_triangle_uvw(self.handle, tri, _ptr(p), _ptr(out))
        return out


def math3d(a, b) -> np.ndarray:
    a = np.ascontiguousarray(a, np.float64)
    b = np.ascontiguousarray(b, np.float64)
    out = np.zeros(9)
    Oracle.lib().mto_math3d(_ptr(a), _ptr(b), _ptr(out))
    return out


def quantize(color) -> np.ndarray:
    c = np.ascontiguousarray(color, np.float64)
    out = np.zeros(3, np.uint8)
    Oracle.lib().mto_quantize(_ptr(c), _ptr(out))
    return out


# ----------------------------------------------------------------------------------------------------
# the unmodified reference
# ----------------------------------------------------------------------------------------------------

class Reference:
    """The unmodified reference renderer (oracle/_ref); one scene at a time (process-global state)."""

    _lib = None

    @classmethod
    def available(cls) -> bool:
        try:
            build()
        except Exception:
            pass
        return os.path.exists(REF_SO)

    @classmethod
    def lib(cls):
        if cls._lib is None:
            build()
            lib = ctypes.CDLL(REF_SO)
            lib.ref_load_obj.argtypes = [ctypes.c_char_p]
            lib.ref_set_lights.argtypes = [ctypes.c_void_p, ctypes.c_int]
            lib.ref_set_depth.argtypes = [ctypes.c_int]
            lib.ref_scene_aabb.argtypes = [ctypes.c_void_p]
            lib.ref_render.argtypes = [ctypes.c_void_p] + [ctypes.c_int] * 6 + [ctypes.c_void_p] * 4
            lib.ref_intersect.argtypes = [ctypes.c_int64] + [ctypes.c_void_p] * 5
            cls._lib = lib
        return cls._lib

    def __init__(self, obj_path):
        if self.lib().ref_load_obj(obj_path.encode()) != 0:
            raise RuntimeError("reference LoadObj failed: %s" % obj_path)

    def set_lights(self, rows):
        arr = np.ascontiguousarray(np.array(rows, np.float64).reshape(-1, 12))
        self.lib().ref_set_lights(_ptr(arr), arr.shape[0])

    def threads(self):
        return self.lib().ref_num_threads()

    def set_threads(self, n: int):
        """Team size of the reference's OpenMP loop (a launcher may have exported OMP_NUM_THREADS=1)."""
        lib = self.lib()
        if hasattr(lib, "ref_set_threads"):
            lib.ref_set_threads.argtypes = [ctypes.c_int]
            lib.ref_set_threads(int(n))

    def aabb(self):
        out = np.zeros(6)
        self.lib().ref_scene_aabb(_ptr(out))
        return out

    def render(self, cam, image_w, image_h, chunk=None, depth=5, debug=True):
        cx, cy, cw, ch = chunk if chunk is not None else (0, 0, image_w, image_h)
        self.lib().ref_set_depth(depth)
        c = np.ascontiguousarray(np.array(cam, np.float64))
        rgb = np.zeros((ch, cw, 3), np.uint8)
        line_no = np.zeros((ch, cw), np.int32) if debug else None
        points = np.zeros((ch, cw, 3), np.float64) if debug else None
        sec = ctypes.c_double(0.0)
        rc = self.lib().ref_render(_ptr(c), image_w, image_h, cx, cy, cw, ch, _ptr(rgb), _ptr(line_no), _ptr(points),
                                   ctypes.cast(ctypes.byref(sec), ctypes.c_void_p))
        assert rc == 0
        return dict(rgb=rgb, line_no=line_no, points=points, seconds=sec.value)

    def intersect(self, origins, dirs):
        o = np.ascontiguousarray(origins, np.float64).reshape(-1, 3)
        d = np.ascontiguousarray(dirs, np.float64).reshape(-1, 3)
        n = o.shape[0]
        line_no = np.full(n, -1, np.int32)
        t = np.zeros(n)
        p = np.zeros((n, 3))
        rc = self.lib().ref_intersect(n, _ptr(o), _ptr(d), _ptr(line_no), _ptr(t), _ptr(p))
        assert rc == 0
        return dict(line_no=line_no, t=t, point=p)
