"""Generates the committed golden fixtures by running the UNMODIFIED reference (oracle/_ref) in the build
container, where /root/reference exists.  The fixtures travel to the GPU box, the reference sources do not.

    python tests/golden/make_golden.py

Per case it writes <case>.obj/.mtl(/.ppm) (the exact input files) and <case>.npz with what the reference
returned through its public API: RGB24 frame, PerPixelDebugInfo line numbers and points
(mythtracer.cc:24-36), OctTree::IntersectRay answers for seeded random rays, and the scene AABB.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from mythtracer_b200 import scenegen  # noqa: E402
from oracle import oracle_py  # noqa: E402
from tests import scenes  # noqa: E402

CASES = {
    # name: (builder, width, height, depth)
    "room_small": (96, 72, 3),
    "room_textured": (80, 60, 2),
    "lattice": (65, 49, 5),
}


def build_case(name):
    if name == "room_small":
        files = scenegen.generate_scene(HERE, 600, seed=21, textured=False, name=name)
        return files.obj_path, files.camera, files.lights[:2]
    if name == "room_textured":
        files = scenegen.generate_scene(HERE, 500, seed=22, textured=True, name=name)
        return files.obj_path, (120.5, 40.25, 60.125, -8.0, 35.0, 2.0, 95.0), files.lights[:3]
    if name == "lattice":
        return scenes.lattice_scene(HERE, name)
    raise KeyError(name)


def main():
    assert os.path.isdir("/root/reference/VerStarting"), "golden vectors are made where the reference is mounted"
    for name, (w, h, depth) in CASES.items():
        obj, cam, lights = build_case(name)
        ref = oracle_py.Reference(obj)
        ref.set_lights(lights)
        img = ref.render(cam, w, h, depth=depth)
        aabb = ref.aabb()
        rng = np.random.default_rng(1234)
        o, d = scenes.random_rays(rng, aabb, 3000)
        d[:200] = np.eye(3)[rng.integers(0, 3, 200)]       # axis-parallel: +-inf inverse direction
        hits = ref.intersect(o, d)
        np.savez_compressed(os.path.join(HERE, name + ".npz"), camera=np.array(cam), lights=np.array(lights),
                            width=w, height=h, depth=depth, rgb=img["rgb"], line_no=img["line_no"],
                            points=img["points"], aabb=aabb, ray_o=o, ray_d=d, ray_line_no=hits["line_no"],
                            ray_t=hits["t"], ray_point=hits["point"])
        print(name, "triangles:", sum(1 for l in open(obj) if l.startswith("f ")), "miss px:", int((img["line_no"] < 0).sum()))


if __name__ == "__main__":
    main()
