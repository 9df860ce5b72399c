"""Scene helpers shared by the tests: BASELINE configs at test sizes and small torture scenes."""
import os

import numpy as np

from mythtracer_b200 import scenegen

LIGHT_RIG = scenegen._default_rig()


def config_scene(name, scene_dir, scale=1.0):
    return scenegen.generate_config(name, scene_dir, scale)


def write_obj(path, text, mtl_text=None):
    os.makedirs(os.path.dirname(path), exist_ok=True)
    with open(path, "w") as f:
        f.write(text)
    if mtl_text is not None:
        with open(os.path.splitext(path)[0] + ".mtl", "w") as f:
            f.write(mtl_text)
    return path


BASIC_MTL = """newmtl matte
Ka 0.5 0.4 0.3
Kd 0.6 0.5 0.4
Ks 0.3 0.3 0.3
Ns 20
newmtl mirror
Ka 0.1 0.1 0.1
Kd 0.2 0.2 0.2
Ks 0.9 0.9 0.9
Ns 100
Refl 0.8
newmtl glass
Ka 0.05 0.05 0.05
Kd 0.1 0.1 0.1
Ks 0.7 0.7 0.7
Ns 80
Tr 0.7
Tf 0.9 1.0 0.8
Ni 1.5
Refl 0.1
"""


def lattice_scene(scene_dir, name="lattice"):
    """Axis-aligned boxes on integer coordinates seen by an on-axis camera: direction components that are
    exactly 0, origins that lie exactly on box planes (NaN slab tests), shared-edge ties."""
    lines = ["mtllib %s.mtl" % name]
    verts = []
    faces = []

    def quad(a, b, c, d, mtl):
        base = len(verts)
        verts.extend([a, b, c, d])
        faces.append((mtl, base + 1, base + 2, base + 3))
        faces.append((mtl, base + 3, base + 4, base + 1))

    def box(lo, hi, mtl):
        x0, y0, z0 = lo
        x1, y1, z1 = hi
        quad((x0, y0, z0), (x1, y0, z0), (x1, y1, z0), (x0, y1, z0), mtl)
        quad((x0, y0, z1), (x1, y0, z1), (x1, y1, z1), (x0, y1, z1), mtl)
        quad((x0, y0, z0), (x0, y0, z1), (x0, y1, z1), (x0, y1, z0), mtl)
        quad((x1, y0, z0), (x1, y0, z1), (x1, y1, z1), (x1, y1, z0), mtl)
        quad((x0, y0, z0), (x1, y0, z0), (x1, y0, z1), (x0, y0, z1), mtl)
        quad((x0, y1, z0), (x1, y1, z0), (x1, y1, z1), (x0, y1, z1), mtl)

    box((0, 0, 0), (16, 8, 16), "matte")          # the room
    mats = ["matte", "mirror", "glass"]
    k = 0
    for ix in range(2, 14, 3):
        for iz in range(4, 14, 3):
            box((ix, 0, iz), (ix + 2, 1 + (k % 3), iz + 2), mats[k % 3])
            k += 1
    quad((4, 2, 8), (12, 2, 8), (12, 6, 8), (4, 6, 8), "glass")   # pane through the room centre plane
    for v in verts:
        lines.append("v %.6f %.6f %.6f" % v)
    lines.append("vn 0 1 0")
    cur = None
    for mtl, a, b, c in faces:
        if mtl != cur:
            lines.append("usemtl %s" % mtl)
            cur = mtl
        lines.append("f %d//1 %d//1 %d//1 " % (a, b, c))
    path = os.path.join(scene_dir, name + ".obj")
    write_obj(path, "\n".join(lines) + "\n", BASIC_MTL)
    cam = (8.0, 4.0, 2.0, 0.0, 0.0, 0.0, 90.0)   # on the x = 8 plane, looking down +z: centre column has dir.x == 0
    lights = [(8.0, 7.0, 3.0, 0.2, 0.2, 0.2, 0.9, 0.9, 0.9, 1.0, 1.0, 1.0),
              (3.0, 6.5, 12.0, 0.0, 0.0, 0.0, 0.5, 0.5, 0.5, 0.5, 0.5, 0.5)]
    return path, cam, lights


def random_rays(rng, aabb, n):
    lo, hi = np.asarray(aabb[:3]), np.asarray(aabb[3:])
    o = lo + rng.random((n, 3)) * (hi - lo)
    d = rng.normal(size=(n, 3))
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    return o, d
