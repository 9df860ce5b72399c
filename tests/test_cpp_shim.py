"""The header-compatible C++ classes of include/mythtracer/ (the reference's API surface for this path):
a caller written like the reference's main_local.cc / octtree_test.cc is compiled against them and run."""
import os
import shutil
import subprocess

import numpy as np
import pytest

from tests import scenes

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def shim_binary(product_lib, tmp_path_factory):
    if shutil.which("g++") is None:
        pytest.skip("g++ not available")
    out = str(tmp_path_factory.mktemp("shim") / "shim_check")
    lib_dir = os.path.join(ROOT, "mythtracer_b200")
    subprocess.check_call(["g++", "-std=c++17", "-O1", "-Wall", "-Wextra", "-Werror", "-I", os.path.join(ROOT, "include"),
                           os.path.join(ROOT, "tests", "cpp", "shim_check.cc"), "-L", lib_dir, "-lmythtracer_b200",
                           "-Wl,-rpath," + lib_dir, "-o", out])
    return out


def test_reference_style_caller_compiles_and_loads(shim_binary, scene_dir):
    files, cfg = scenes.config_scene("C2", scene_dir, 0.1)
    res = subprocess.run([shim_binary, "host", files.obj_path], capture_output=True, text=True)
    assert res.returncode == 0, res.stderr
    assert "host ok" in res.stdout
    assert "triangles %d materials 12 textures 3" % files.n_triangles in res.stdout


@pytest.mark.gpu
def test_reference_style_caller_renders(shim_binary, scene_dir, tmp_path):
    from mythtracer_b200 import Light, MythTracer
    files, cfg = scenes.config_scene("C2", scene_dir, 0.1)
    raw = str(tmp_path / "dump_00000.raw")
    W, H = 200, 120
    res = subprocess.run([shim_binary, "render", files.obj_path, raw, str(W), str(H)], capture_output=True, text=True)
    assert res.returncode == 0, res.stdout[-2000:] + res.stderr[-2000:]
    assert "render ok" in res.stdout
    got = np.fromfile(raw, np.uint8).reshape(H, W, 3)
    mt = MythTracer(max_depth=3)
    assert mt.LoadObj(files.obj_path)
    mt.GetScene().lights = [Light.from_tuple(l) for l in scenes.LIGHT_RIG[:2]]
    ref = mt.RayTrace(W, H, (301.37, 57.21, 161.13, 4.0, 243.0, 0.0, 110.0))
    assert np.array_equal(got, ref)


def test_triangle_virtuals_match_the_reference_arithmetic(shim_binary, oracle_mod, tmp_path):
    """Triangle::IntersectRay of the header shim (called through Primitive*, primitive.h:20-24) against the oracle -
    itself pinned bit for bit to the unmodified reference - on one-triangle scenes: there OctTree::IntersectRay is the
    root slab test on the triangle's own box followed by exactly this call (octtree.cc:26-40,177-196)."""
    from mythtracer_b200.api import MTL_DTYPE, TRI_DTYPE
    rng = np.random.default_rng(21)
    n = 400
    verts = rng.uniform(-5.0, 5.0, (n, 3, 3))
    verts[::7] = np.round(verts[::7])                       # lattice triangles: exact ties in the slab test
    target = verts.mean(1) + rng.normal(0, 0.8, (n, 3))     # aim near the triangle: hits, edge grazes and misses
    origin = target + rng.normal(0, 6.0, (n, 3))
    d = target - origin
    d /= np.sqrt((d * d).sum(1))[:, None]
    d[::11, 0] = 0.0                                        # zero components: the +-inf inverse path
    d[5::13] *= rng.uniform(0.2, 5.0, (len(d[5::13]), 1))   # un-normalised directions
    rec = np.concatenate([verts.reshape(n, 9), origin, d], 1)
    rec.astype("<f8").tofile(tmp_path / "in.bin")
    res = subprocess.run([shim_binary, "tri", str(tmp_path / "in.bin"), str(tmp_path / "out.bin")], capture_output=True, text=True)
    assert res.returncode == 0, res.stderr
    got = np.fromfile(tmp_path / "out.bin", "<f8").reshape(n, 5)
    hits = 0
    for i in range(n):
        arr = np.zeros(1, TRI_DTYPE)
        arr["vertex"] = verts[i].reshape(9)
        arr["material"] = -1
        ref = oracle_mod.Oracle(arr, np.zeros(0, MTL_DTYPE), []).intersect(origin[i:i + 1], d[i:i + 1])
        if ref["tri"][0] < 0:
            assert got[i, 0] == 0.0, i
        else:
            hits += 1
            assert got[i, 0] == 1.0, i
            assert got[i, 1] == ref["t"][0] and np.array_equal(got[i, 2:], ref["point"][0]), i
    assert 50 < hits < n - 50
