"""Host-side logic of the product, no GPU needed: the C ABI library loads and exports what the header
declares, the loader and the octree builder agree with independent implementations, camera / WorkChunk
(de)serialisation follow the reference's wire forms, and rendering without a device fails loudly."""
import os
import re

import numpy as np
import pytest

from tests import scenes

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol(product_lib):
    from mythtracer_b200 import api
    header = open(os.path.join(ROOT, "include", "mythtracer_b200.h")).read()
    declared = set(re.findall(r"\b(mtb_[a-z_0-9]+)\s*\(", header))
    assert declared, "no declarations parsed"
    assert declared == set(api.EXPORTED_SYMBOLS)
    for name in declared:
        assert getattr(product_lib, name) is not None
    assert b"sm_100a" in product_lib.mtb_version()


def test_no_cpu_fallback(product_lib):
    """A context without a device can load and inspect scenes but must refuse to render."""
    from mythtracer_b200 import MythTracer, MythTracerError
    files, cfg = scenes.config_scene("C1", "/tmp/mtb_scenes")
    mt = MythTracer(host_only=True)
    assert mt.LoadObj(files.obj_path)
    with pytest.raises(MythTracerError, match="no CPU fallback"):
        mt.render_chunk(files.camera, 16, 16, 0, 0, 16, 16)
    with pytest.raises(MythTracerError, match="no CPU fallback"):
        mt.intersect_rays([[0, 0, 0]], [[0, 0, 1]])
    assert mt.RayTrace(16, 16, files.camera) is None


def test_product_does_not_import_the_oracle():
    """The product path must never route through oracle/ (checker only)."""
    pkg = os.path.join(ROOT, "mythtracer_b200")
    for dirpath, _, names in os.walk(pkg):
        for n in names:
            if n.endswith((".py", ".cu", ".cc", ".h")):
                text = open(os.path.join(dirpath, n), errors="ignore").read()
                assert "oracle_py" not in text and "mt_oracle" not in text and "libmythtracer_ref" not in text, n


@pytest.mark.parametrize("name,scale", [("C1", 1.0), ("C2", 0.1)])
def test_loader_matches_independent_parser(product_lib, oracle_mod, scene_dir, name, scale):
    from mythtracer_b200 import MythTracer
    files, cfg = scenes.config_scene(name, scene_dir, scale)
    mt = MythTracer(host_only=True)
    assert mt.LoadObj(files.obj_path)
    tris, mtls = mt.scene_arrays()
    ptris, pmtls, ptex = oracle_mod.read_obj(files.obj_path)
    assert tris.tobytes() == ptris.tobytes()
    assert mtls.tobytes() == pmtls.tobytes()
    assert mt.scene_info()["n_textures"] == len(ptex)


def test_loader_quirks(product_lib, oracle_mod, scene_dir):
    """objreader.cc quirks: lost last token, quads, %i indices, 128-byte line pieces, unknown usemtl."""
    from mythtracer_b200 import MythTracer
    base = "mtllib q.mtl\nv 0 0 0\nv 1 0 0\nv 1 1 0\nv 0 1 0\nvn 0 0 1\nvt 0.5 0.25\nusemtl matte\n"

    def load(text):
        path = scenes.write_obj(os.path.join(scene_dir, "q.obj"), base + text, scenes.BASIC_MTL)
        mt = MythTracer(host_only=True)
        ok = mt.LoadObj(path)
        return ok, mt

    ok, mt = load("f 1 2 3\n")                       # no trailing space: third vertex dropped -> rejected
    assert not ok and "unsupported face count (2)" in mt.last_error()
    ok, mt = load("f 1 2 3 4\n")                     # quad without trailing space -> ONE triangle
    assert ok and mt.scene_info()["n_triangles"] == 1
    ok, mt = load("f 1 2 3 4 \n")                    # proper quad -> (0,1,2) + (2,3,0), same line number
    tris, _ = mt.scene_arrays()
    assert ok and len(tris) == 2 and tris["line_no"].tolist() == [8, 8]
    assert tris[1]["vertex"].tolist() == [1, 1, 0, 0, 1, 0, 0, 0, 0]
    ok, mt = load("f 0x1//1 02//1 3//1 \n")          # %i accepts hex and octal
    tris, _ = mt.scene_arrays()
    assert ok and tris[0]["vertex"].tolist() == [0, 0, 0, 1, 0, 0, 1, 1, 0] and tris[0]["normal"][2] == 1.0
    ok, mt = load("f 1/1 2/1 3/1 \n")                # v/vt: uvw copied, no normals
    tris, _ = mt.scene_arrays()
    assert ok and tris[0]["uvw"][:2].tolist() == [0.5, 0.25] and not tris[0]["normal"].any()
    ok, mt = load("usemtl nope\nf 1 2 3 \n")         # unknown material -> mtl == nullptr, parsing goes on
    tris, _ = mt.scene_arrays()
    assert ok and tris[0]["material"] == -1
    ok, mt = load("# " + "x" * 200 + "\nf 1 2 3 \n")  # a 203-byte comment line is two fgets pieces (lines 8 and 9)
    tris, _ = mt.scene_arrays()
    assert ok and tris[0]["line_no"] == 10
    ptris, _, _ = oracle_mod.read_obj(os.path.join(scene_dir, "q.obj"))
    assert ptris.tobytes() == tris.tobytes()
    ok, mt = load("f 1 2 9 \n")                      # out-of-range index (undefined upstream) is refused
    assert not ok and "out of range" in mt.last_error()


def test_loader_result_does_not_depend_on_the_chunking(product_lib, scene_dir, tmp_path, monkeypatch):
    """The OBJ reader parses the file in chunks on several threads (obj_loader.cc); triangles, their order, materials
    and 0-based line numbers (pieces of 127 bytes count as lines, objreader.cc:233-235) must be those of a sequential
    read, whatever the chunking - including over-long lines, a usemtl in front of its mtllib, CR LF line ends, and the
    error a sequential reader would meet first."""
    from mythtracer_b200 import MythTracer
    from tests import scenes
    files, cfg = scenes.config_scene("C2", scene_dir, 0.1)
    quirky = tmp_path / "quirky.obj"
    body = ["usemtl matte", "mtllib quirky.mtl", "# " + "x" * 300, "v 0 0 0", "v 1 0 0\r", "v 0 1 0", "vn 0 0 1", "usemtl matte",
            "f 1//1 2//1 3//1 ", "o thing", "v 0 0 1" + " " * 200, "usemtl nowhere", "f 1 2 4 ", "usemtl mirror", "f 1 2 3 4 "]
    body += ["v %d 0.5 0.25" % i for i in range(400)] + ["f 5 6 7 ", "f 2 3 400 "]
    quirky.write_text("\n".join(body) + "\n")
    (tmp_path / "quirky.mtl").write_text(scenes.BASIC_MTL)
    broken = tmp_path / "broken.obj"
    broken.write_text("\n".join(["v 0 0 0", "v 1 0 0", "v 0 1 0", "f 1 2 3 "] * 50 + ["f 1 2 900 "] + ["v 1 1 1"] * 900 + ["v oops"]) + "\n")
    results = {}
    for threads in ("1", "3", "16"):
        monkeypatch.setenv("MTB_LOADER_THREADS", threads)
        for path in (files.obj_path, str(quirky)):
            mt = MythTracer(host_only=True)
            assert mt.LoadObj(path), mt.last_error()
            tris, mtls = mt.scene_arrays()
            key = os.path.basename(path)
            if key in results:
                assert np.array_equal(results[key][0], tris) and np.array_equal(results[key][1], mtls), (key, threads)
            results[key] = (tris, mtls)
        mt = MythTracer(host_only=True)
        assert not mt.LoadObj(str(broken))
        assert "index out of range" in mt.last_error(), mt.last_error()  # the earlier of the two errors
    q = results["quirky.obj"][0]
    assert len(q) == 6 and q["material"].tolist()[:2] == [0, -1]  # first usemtl precedes its mtllib: unknown at that point -> none
    assert q["line_no"][0] > 8  # the 300-character comment counts as three lines


@pytest.mark.parametrize("name,scale", [("C1", 1.0), ("C2", 0.2), ("C4", 0.02)])
def test_octree_builder_matches_oracle(product_lib, oracle_mod, scene_dir, name, scale):
    """Same boxes, same list membership, same depth as the reference rules (octtree.cc:46-135)."""
    from mythtracer_b200 import MythTracer
    files, cfg = scenes.config_scene(name, scene_dir, scale)
    mt = MythTracer(host_only=True)
    assert mt.LoadObj(files.obj_path)
    tris, mtls = mt.scene_arrays()
    orc = oracle_mod.Oracle(tris, mtls, [])
    info, tree = mt.scene_info(), orc.tree_info()
    assert info["n_nodes"] == tree["nodes"] and info["tree_depth"] == tree["depth"]
    assert info["root_list"] == tree["root_list"] and info["biggest_list"] == tree["biggest_list"]
    assert info["interior_triangles"] == tree["interior_tris"]
    b1, d1 = mt.triangle_nodes()
    b2, d2 = orc.triangle_nodes()
    assert np.array_equal(b1, b2) and np.array_equal(d1, d2)
    assert np.array_equal(orc.aabb(), np.array(info["aabb_min"] + info["aabb_max"]))


def test_camera_sensor_and_serialisation(product_lib, oracle_mod):
    from mythtracer_b200 import Camera
    rng = np.random.default_rng(3)
    for _ in range(50):
        cam = Camera(tuple(rng.uniform(-300, 300, 3)), *rng.uniform(-180, 180, 3), rng.uniform(20, 140))
        w, h = int(rng.integers(16, 4000)), int(rng.integers(16, 2200))
        t = (*cam.origin, cam.pitch, cam.yaw, cam.roll, cam.aov)
        assert np.array_equal(cam.GetSensor(w, h), oracle_mod.Oracle.camera_sensor(t, w, h))
        blob = cam.Serialize()
        assert len(blob) == Camera.kSerializedSize == 56
        back = Camera.Deserialize(blob)
        assert back.Serialize() == blob
    assert Camera.Deserialize(b"\0" * 55) is None


def test_workchunk_wire_forms():
    """WorkChunk::SerializeInput / DeserializeInput / SerializeOutput / DeserializeOutput (mythtracer.cc:314-429)."""
    from mythtracer_b200 import WorkChunk
    c = WorkChunk(1920, 1080, 128, 256, 128, 128)
    blob = c.SerializeInput()
    assert blob == np.array([1920, 1080, 128, 256, 128, 128], "<u4").tobytes()
    d = WorkChunk()
    assert d.DeserializeInput(blob) and (d.chunk_x, d.chunk_y, d.chunk_width) == (128, 256, 128)
    for bad in ([0, 1080, 0, 0, 1, 1], [1920, 1080, 1900, 0, 128, 128], [100001, 10, 0, 0, 1, 1], [64, 64, 0, 0, 0, 8]):
        assert not WorkChunk().DeserializeInput(np.array(bad, "<u4").tobytes())
    assert not WorkChunk().DeserializeInput(blob[:-1])
    c.output_bitmap = (np.arange(128 * 128 * 3) % 251).astype(np.uint8).reshape(128, 128, 3)
    out = c.SerializeOutput()
    assert out[:4] == np.array([128 * 128 * 3], "<u4").tobytes() and len(out) == 4 + 128 * 128 * 3
    e = WorkChunk(chunk_width=128, chunk_height=128)
    assert e.DeserializeOutput(out) and np.array_equal(e.output_bitmap, c.output_bitmap)
    assert not WorkChunk(chunk_width=64, chunk_height=128).DeserializeOutput(out)
    assert not e.DeserializeOutput(out[:3])


def test_strip_partition_math():
    from mythtracer_b200 import tiles
    for h in (1, 7, 8, 9, 270, 1080, 2160):
        for world in (1, 2, 3, 4, 8):
            rows = sorted(r for k in range(world) for r in tiles.owned_rows(h, k, world))
            assert rows == list(range(h))
            assert tiles.padded_height(h, world) % (8 * world) == 0 and tiles.padded_height(h, world) >= h
            for k in range(world):
                assert all(tiles.strip_owner(s, world) == k for s in tiles.owned_strips(h, k, world))


def test_octree_depth_limit_and_degenerate_inputs(product_lib, oracle_mod):
    """The reference has no depth cap (octtree.cc:52-55): 16 coincident primitives recurse until its stack
    overflows.  The builder refuses such scenes with MTB_ERR_LIMIT instead; an empty scene and scenes below the
    split threshold are fine."""
    from mythtracer_b200 import MythTracer, MythTracerError
    from mythtracer_b200.api import MTL_DTYPE, TRI_DTYPE
    mt = MythTracer(host_only=True)
    tris = np.zeros(16, TRI_DTYPE)
    tris["vertex"] = [1.3, 1.7, 1.1] * 3   # 16 coincident zero-size triangles: they fit into ever smaller children
    tris["material"] = -1
    with pytest.raises(MythTracerError, match="deeper than"):
        mt.upload(tris, np.zeros(0, MTL_DTYPE))
    mt.upload(tris[:15], np.zeros(0, MTL_DTYPE))                        # below SPLIT_BOUNDARY: one node
    info = mt.scene_info()
    assert info["n_nodes"] == 1 and info["tree_depth"] == 0 and info["root_list"] == 15
    mt.upload(np.zeros(0, TRI_DTYPE), np.zeros(0, MTL_DTYPE))           # empty scene: root box {0,0,0}-{0,0,0}
    info = mt.scene_info()
    assert info["n_triangles"] == 0 and info["n_nodes"] == 1 and info["aabb_max"] == [0.0, 0.0, 0.0]
    bad = np.zeros(1, TRI_DTYPE)
    bad["material"] = 3
    with pytest.raises(MythTracerError, match="material"):
        mt.upload(bad, np.zeros(1, MTL_DTYPE))


def test_scene_bvh_is_a_conservative_partition(product_lib, scene_dir):
    """The certified fast traversal relies on these properties of the scene BVH: every triangle
    is referenced from at least one leaf; the stored box of a child (FP32, padded and rounded outwards) contains
    everything below it - for a triangle with ONE reference its exact FP64 box, for a triangle that was split into
    several references (SplitTriangle, scene_build.cc) the boxes of its references together cover the triangle.
    Checked on generated scenes (C1: big wall triangles are split; the C4 stress scene: room-spanning sticks), on
    scenes smaller than a leaf and on an empty scene."""
    from mythtracer_b200 import MythTracer
    from mythtracer_b200.api import MTL_DTYPE, TRI_DTYPE
    from tests import scenes
    import sys
    sys.setrecursionlimit(10000)

    from tests.bvh_check import check_scene_bvh as check

    files, cfg = scenes.config_scene("C1", scene_dir)
    mt = MythTracer(host_only=True)
    assert mt.LoadObj(files.obj_path)
    tris, mtls = mt.scene_arrays()
    check(mt, tris, expect_splits=True)
    for n in (1, 2, 3, 5):
        small = MythTracer(host_only=True)
        small.upload(tris[:n], mtls)
        check(small, tris[:n])
    empty = MythTracer(host_only=True)
    empty.upload(np.zeros(0, TRI_DTYPE), np.zeros(0, MTL_DTYPE))
    check(empty, np.zeros(0, TRI_DTYPE))
    files4, _ = scenes.config_scene("C4", scene_dir, 0.02)
    mt4 = MythTracer(host_only=True)
    assert mt4.LoadObj(files4.obj_path)
    check(mt4, mt4.scene_arrays()[0], expect_splits=True)


def _png_filter_rows(rows, bpp, filters):
    """Filtered scanlines (filter byte + bytes) of one PNG (sub-)image; rows: uint8 [h, bytes per row]."""
    raw = bytearray()
    prev = np.zeros(rows.shape[1], np.int32)
    for y in range(rows.shape[0]):
        cur = rows[y].astype(np.int32)
        f = filters[y % len(filters)]
        a = np.concatenate([np.zeros(bpp, np.int32), cur[:-bpp]])
        c = np.concatenate([np.zeros(bpp, np.int32), prev[:-bpp]])
        if f == 0:
            pred = np.zeros_like(cur)
        elif f == 1:
            pred = a
        elif f == 2:
            pred = prev
        elif f == 3:
            pred = (a + prev) // 2
        else:
            p = a + prev - c
            pa, pb, pc = np.abs(p - a), np.abs(p - prev), np.abs(p - c)
            pred = np.where((pa <= pb) & (pa <= pc), a, np.where(pb <= pc, prev, c))
        raw.append(f)
        raw += ((cur - pred) & 255).astype(np.uint8).tobytes()
        prev = cur
    return raw


def _png_bytes(img, color_type, depth=8, filters=(0, 1, 2, 3, 4), palette=None, level=6, width=None, fixed=False, adam7=False):
    """Minimal PNG writer (zlib from the standard library) with a chosen scanline filter per row; adam7: the seven
    interlace passes (8- and 16-bit samples only)."""
    import struct, zlib
    h, w = img.shape[:2]
    rows = img.reshape(h, -1).astype(np.uint8)
    bpp = max(1, rows.shape[1] // w) if depth >= 8 else 1
    if adam7:
        assert depth >= 8
        px = rows.reshape(h, w, bpp)
        raw = bytearray()
        for x0, y0, dx, dy in ((0, 0, 8, 8), (4, 0, 8, 8), (0, 4, 4, 8), (2, 0, 4, 4), (0, 2, 2, 4), (1, 0, 2, 2), (0, 1, 1, 2)):
            sub = px[y0::dy, x0::dx]
            if sub.shape[0] and sub.shape[1]:
                raw += _png_filter_rows(np.ascontiguousarray(sub).reshape(sub.shape[0], -1), bpp, filters)
    else:
        raw = _png_filter_rows(rows, bpp, filters)
    w = width or w

    def chunk(tag, body):
        return struct.pack(">I", len(body)) + tag + body + struct.pack(">I", zlib.crc32(tag + body) & 0xffffffff)
    out = b"\x89PNG\r\n\x1a\n" + chunk(b"IHDR", struct.pack(">IIBBBBB", w, h, depth, color_type, 0, 0, 1 if adam7 else 0))
    if palette is not None:
        out += chunk(b"PLTE", palette.astype(np.uint8).tobytes())
    co = zlib.compressobj(level, zlib.DEFLATED, 15, 8, zlib.Z_FIXED if fixed else zlib.Z_DEFAULT_STRATEGY)
    comp = co.compress(bytes(raw)) + co.flush()
    out += chunk(b"IDAT", comp[:len(comp) // 2]) + chunk(b"IDAT", comp[len(comp) // 2:]) + chunk(b"IEND", b"")
    return out


def test_texture_decoders_give_the_same_texels(product_lib, tmp_path):
    """PPM, PNG (RGB / RGBA / grey / palette, every scanline filter, stored / fixed / dynamic deflate
    blocks), BMP (24 / 32 bits, both row orders) and TGA (raw / run-length, both origins) files of one image must
    all load to the same RGBA32 texels through MtlFileReader's map_Ka path."""
    import struct
    from mythtracer_b200 import MythTracer, MythTracerError
    rng = np.random.default_rng(3)
    h, w = 37, 53
    img = rng.integers(0, 256, (h, w, 3), dtype=np.uint8)
    img[5:20, 7:30] = (200, 10, 60)            # flat areas: long deflate matches and TGA runs
    img[25:, :] = img[25:26, :]
    files = {}
    files["a.ppm"] = b"P6\n# comment\n%d %d\n255\n" % (w, h) + img.tobytes()
    files["b.png"] = _png_bytes(img, 2)
    rgba = np.dstack([img, rng.integers(0, 256, (h, w, 1), dtype=np.uint8)])
    files["c.png"] = _png_bytes(rgba, 6, level=0)                       # stored blocks
    files["d.png"] = _png_bytes(img, 2, filters=(4,), fixed=True)       # fixed-Huffman blocks, Paeth only
    rgb16 = np.dstack([img[..., c // 2] if c % 2 == 0 else rng.integers(0, 256, (h, w), dtype=np.uint8) for c in range(6)])
    files["e.png"] = _png_bytes(rgb16, 2, depth=16)                     # 16 bits: the high byte is the texel
    bgr = img[..., ::-1]
    pad = (-w * 3) % 4
    rows_up = b"".join(bgr[y].tobytes() + b"\0" * pad for y in range(h - 1, -1, -1))
    hdr = lambda bits, height, size: b"BM" + struct.pack("<IHHI", 54 + size, 0, 0, 54) + struct.pack("<IiiHHIIiiII", 40, w, height, 1, bits, 0, size, 2835, 2835, 0, 0)
    files["f.bmp"] = hdr(24, h, len(rows_up)) + rows_up
    bgra = np.dstack([bgr, np.full((h, w, 1), 255, np.uint8)])
    files["g.bmp"] = hdr(32, -h, w * h * 4) + bgra.tobytes()            # top-down, 32 bits
    tga_hdr = lambda typ, bits, desc: struct.pack("<BBBHHBHHHHBB", 0, 0, typ, 0, 0, 0, 0, 0, w, h, bits, desc)
    files["h.tga"] = tga_hdr(2, 24, 0x20) + bgr.tobytes()               # raw, top-left origin
    rle = bytearray()
    flat = bgr[::-1].reshape(-1, 3)                                     # bottom-left origin
    i = 0
    while i < len(flat):
        run = 1
        while i + run < len(flat) and run < 128 and np.array_equal(flat[i + run], flat[i]):
            run += 1
        if run > 1:
            rle.append(128 | (run - 1))
            rle += flat[i].tobytes()
        else:
            lit = 1
            while i + lit < len(flat) and lit < 128 and not np.array_equal(flat[i + lit], flat[i + lit - 1]):
                lit += 1
            run = lit
            rle.append(lit - 1)
            rle += flat[i:i + lit].tobytes()
        i += run
    files["i.tga"] = tga_hdr(10, 24, 0x00) + bytes(rle)
    grey = img[..., 0]
    files["j.png"] = _png_bytes(grey, 0)
    pal = rng.integers(0, 256, (16, 3), dtype=np.uint8)
    idx = rng.integers(0, 16, (h, w), dtype=np.uint8)
    packed = np.zeros((h, (w + 1) // 2), np.uint8)
    packed[:, :w // 2] = (idx[:, 0:w - 1:2] << 4) | idx[:, 1::2]
    if w % 2:
        packed[:, -1] = idx[:, -1] << 4
    files["k.png"] = _png_bytes(packed, 3, depth=4, palette=pal, filters=(0, 2), width=w)
    for name, data in files.items():
        (tmp_path / name).write_bytes(data)
    mtl = "".join("newmtl m%s\nKa 1 1 1\nmap_Ka %s\n" % (n[0], n) for n in sorted(files))
    (tmp_path / "all.mtl").write_text(mtl)
    mt = MythTracer(host_only=True)
    assert mt.LoadMtl(str(tmp_path / "all.mtl")), mt.last_error()
    tex = {mt.texture_name(i): mt.texture(i) for i in range(mt.scene_info()["n_textures"])}
    assert sorted(tex) == sorted(files)
    for name in "abdefghi":
        got = tex[[n for n in files if n[0] == name][0]]
        assert got.shape == (h, w, 4) and np.array_equal(got[..., :3], img), name
    assert np.array_equal(tex["c.png"], rgba)
    assert np.array_equal(tex["j.png"][..., :3], np.dstack([grey] * 3))
    assert np.array_equal(tex["k.png"][..., :3], pal[idx])
    # truncated files fail the whole load, as an undecodable texture does upstream (objreader.cc:467-469)
    (tmp_path / "bad.png").write_bytes(files["b.png"][:200])
    (tmp_path / "bad.mtl").write_text("newmtl x\nmap_Ka bad.png\n")
    assert not MythTracer(host_only=True).LoadMtl(str(tmp_path / "bad.mtl"))


def test_jpeg_and_interlaced_png_textures(product_lib, tmp_path):
    """SURVEY 8 f4, the formats round 1 refused.  JPEG: baseline, progressive, restart intervals, optimised tables,
    4:4:4 / 4:2:2 / 4:2:0 (incl. a chroma plane too narrow for the fancy filter), greyscale -- every texel must equal
    what libjpeg-turbo decodes (tests/golden/jpeg/expected.npz, made by make_jpeg_fixtures.py with Pillow).  PNG: the
    Adam7 passes of RGB / RGBA / grey / 16-bit images, sizes that leave some passes empty."""
    import shutil
    from mythtracer_b200 import MythTracer
    jdir = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "jpeg")
    exp = np.load(os.path.join(jdir, "expected.npz"))
    names = sorted(exp.files)
    assert len(names) >= 11
    for n in names:
        shutil.copy(os.path.join(jdir, n), tmp_path / n)
    rng = np.random.default_rng(5)
    pngs = {}
    for k, (h, w) in enumerate([(37, 53), (1, 1), (2, 3), (5, 1), (9, 4), (8, 8)]):
        img = rng.integers(0, 256, (h, w, 3), dtype=np.uint8)
        pngs["i%d_rgb.png" % k] = (_png_bytes(img, 2, adam7=True), img)
        rgba = rng.integers(0, 256, (h, w, 4), dtype=np.uint8)
        pngs["i%d_rgba.png" % k] = (_png_bytes(rgba, 6, adam7=True, filters=(4, 3, 1)), rgba)
    grey = rng.integers(0, 256, (21, 30), dtype=np.uint8)
    pngs["i_grey.png"] = (_png_bytes(grey, 0, adam7=True), np.dstack([grey] * 3))
    img = rng.integers(0, 256, (19, 23, 3), dtype=np.uint8)
    rgb16 = np.dstack([img[..., c // 2] if c % 2 == 0 else rng.integers(0, 256, (19, 23), dtype=np.uint8) for c in range(6)])
    pngs["i_rgb16.png"] = (_png_bytes(rgb16, 2, depth=16, adam7=True), img)
    for n, (data, _) in pngs.items():
        (tmp_path / n).write_bytes(data)
    every = names + sorted(pngs)
    (tmp_path / "all.mtl").write_text("".join("newmtl m%d\nKa 1 1 1\nmap_Ka %s\n" % (i, n) for i, n in enumerate(every)))
    mt = MythTracer(host_only=True)
    assert mt.LoadMtl(str(tmp_path / "all.mtl")), mt.last_error()
    tex = {mt.texture_name(i): mt.texture(i) for i in range(mt.scene_info()["n_textures"])}
    assert sorted(tex) == sorted(every)
    for n in names:
        assert tex[n].shape[:2] == exp[n].shape[:2], n
        assert np.array_equal(tex[n][..., :3], exp[n]), "%s: %d channels differ from libjpeg-turbo" % (
            n, int((tex[n][..., :3] != exp[n]).sum()))
        assert (tex[n][..., 3] == 255).all()
    for n, (_, want) in pngs.items():
        assert np.array_equal(tex[n][..., :want.shape[2]], want), n
    # CMYK-like four-component frames and truncated streams fail the load
    bad = bytearray((tmp_path / "q90_444.jpg").read_bytes())
    (tmp_path / "cut.jpg").write_bytes(bytes(bad[:300]))
    (tmp_path / "cut.mtl").write_text("newmtl x\nmap_Ka cut.jpg\n")
    got = MythTracer(host_only=True)
    if got.LoadMtl(str(tmp_path / "cut.mtl")):  # a cut stream decodes as far as it goes (zero bits), like libjpeg
        assert got.texture(0).shape[:2] == exp["q90_444.jpg"].shape[:2]

