"""Pins the oracle (CPU restatement) against fixtures the UNMODIFIED reference produced
(tests/golden/make_golden.py) and against the known answers in the reference's own tests."""
import os

import numpy as np
import pytest

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
CASES = ["room_small", "room_textured", "lattice"]


def load_case(name):
    z = np.load(os.path.join(GOLDEN, name + ".npz"))
    return z, os.path.join(GOLDEN, name + ".obj")


@pytest.mark.parametrize("name", CASES)
def test_render_matches_reference_fixture(oracle_mod, name):
    z, obj = load_case(name)
    orc = oracle_mod.Oracle.from_obj(obj)
    orc.set_lights(z["lights"].tolist())
    out = orc.render(z["camera"].tolist(), int(z["width"]), int(z["height"]), depth=int(z["depth"]))
    assert np.array_equal(out["line_no"], z["line_no"])
    # hit points: bit-exact when libm's sin/cos of this host equal those of the host that made the fixture;
    # a 1-ulp difference in the camera basis moves points by ~1e-13, so compare with a tight tolerance
    np.testing.assert_allclose(out["points"], z["points"], rtol=0, atol=1e-9, equal_nan=True)
    assert np.array_equal(out["rgb"], z["rgb"])
    assert np.array_equal(orc.aabb(), z["aabb"])


@pytest.mark.parametrize("name", CASES)
def test_intersect_matches_reference_fixture(oracle_mod, name):
    z, obj = load_case(name)
    orc = oracle_mod.Oracle.from_obj(obj)
    out = orc.intersect(z["ray_o"], z["ray_d"])
    line_no = np.where(out["tri"] >= 0, orc.tris["line_no"][np.maximum(out["tri"], 0)], -1)
    assert np.array_equal(line_no, z["ray_line_no"])
    hit = z["ray_line_no"] >= 0
    assert hit.sum() > 1000
    assert np.array_equal(out["t"][hit], z["ray_t"][hit])          # no libm involved: bit-exact
    assert np.array_equal(out["point"][hit], z["ray_point"][hit])


def test_octtree_test_known_answers(oracle_mod):
    """reference octtree_test.cc:14-72 (with CacheAABB, as the OBJ loader does: objreader.cc:183)."""
    tris = np.zeros(2, oracle_mod.TRI_DTYPE)
    tris[0]["vertex"] = [1, 1, 0, 1, 0, 0, 0, 0, 0]
    tris[1]["vertex"] = [1, 1, 1, 1, 0, 1, 0, 0, 1]
    tris["material"] = -1
    orc = oracle_mod.Oracle(tris, np.zeros(0, oracle_mod.MTL_DTYPE))
    out = orc.intersect([[0.9, 0.9, -10.0], [0.9, 0.9, 10.0], [5.0, 5.0, 5.0]],
                        [[0.0, 0.0, 1.0], [0.0, 0.0, -1.0], [0.0, 0.0, 1.0]])
    assert out["tri"].tolist() == [0, 1, -1]
    assert out["t"][:2].tolist() == [10.0, 9.0]


def test_math3d_test_known_answers(oracle_mod):
    """reference math3d_test.cc:68-89 (tolerance 1e-7 as test_helper.cc:5-12)."""
    out = oracle_mod.math3d([1, 2, 3], [5, 4, 3])
    assert abs(out[0] - 3.7416573867739413) < 1e-7
    assert out[2] == 22.0
    assert out[3:6].tolist() == [-6.0, 12.0, -6.0]
    np.testing.assert_allclose(out[6:9], [0.2672612419124, 0.5345224838248, 0.8017837257372], atol=1e-7)
    assert abs(oracle_mod.math3d([1, 1, 1], [2, 2, 2])[1] - 1.7320508075688772) < 1e-7


def test_quantize_contract(oracle_mod):
    """V3DtoRGB (mythtracer.cc:235-241): clamp, truncate, NaN -> 0."""
    q = oracle_mod.quantize
    assert q([1.5, -0.2, 0.5]).tolist() == [255, 0, 127]
    assert q([1.0, 0.0, 0.999999]).tolist() == [255, 0, 254]
    assert q([float("nan"), 0.25, 1.0000001]).tolist() == [0, 63, 255]


def test_brute_force_equals_octree(oracle_mod):
    """on these scenes the octree is a pure accelerator."""
    z, obj = load_case("room_small")
    orc = oracle_mod.Oracle.from_obj(obj)
    a = orc.intersect(z["ray_o"], z["ray_d"])
    b = orc.intersect(z["ray_o"], z["ray_d"], brute=True)
    assert np.array_equal(a["tri"], b["tri"])
    hit = a["tri"] >= 0
    assert np.array_equal(a["t"][hit], b["t"][hit])
