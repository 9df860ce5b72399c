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
