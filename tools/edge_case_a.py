"""Stated edge case (a) of DESIGN.md section 4 on the device (not part of the test suite: it was constructed after the
round's GPU time was used up).  Prints what the default traversal and MTB_FLAG_EXACT_OCTREE return for the ray of
tests/test_certification_math.py::test_stated_edge_case_sibling_entry_tie_is_a_property_of_the_reference; the
reference returns triangle 0 at t = 4.904..., brute force triangle 1 at t = 3.0."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from mythtracer_b200 import MythTracer, MTB_FLAG_EXACT_OCTREE
from mythtracer_b200.api import MTL_DTYPE, TRI_DTYPE

tris = []


def tri(a, b, c):
    tris.append([*a, *b, *c])


tri((1, 3, 1.0), (3, 1, 1.0), (2.2, 2.2, 3.9))
tri((4, 4, 2.0), (4, 4, 3.25), (6, 2, 2.5))
tri((0, 0, 0), (0.3, 0, 0), (0, 0.3, 0))
tri((8, 8, 8), (7.7, 8, 8), (8, 7.7, 8))
rng = np.random.default_rng(1)
for _ in range(14):
    c = np.array([6.5, 1.0, 6.5]) + rng.uniform(-0.4, 0.4, 3)
    tri(c, c + [0.2, 0, 0], c + [0, 0.2, 0])
arr = np.zeros(len(tris), TRI_DTYPE)
arr["vertex"] = np.array(tris, float)
arr["material"] = -1
arr["line_no"] = np.arange(len(tris))
o = np.array([[7.0, 7.0, 3.0]])
d = np.array([[-1.0, -1.0, -0.125]])
for name, flags in (("default", 0), ("MTB_FLAG_EXACT_OCTREE", MTB_FLAG_EXACT_OCTREE)):
    mt = MythTracer(flags=flags)
    mt.upload(arr, np.zeros(0, MTL_DTYPE))
    r = mt.intersect_rays(o, d)
    print("%-22s triangle %d  t %.15g" % (name, int(r["tri"][0]), float(r["t"][0])))
    mt.close()
