"""Full-size golden fixtures, made by the UNMODIFIED reference (oracle/_ref) in the build container.

    python tests/golden/make_golden_full.py [C2] [C3] [C4[:full]] [C5[:full]]

* C2 (textured, 1280x720) and C3 (BASELINE.json configs[2], 1920x1080, depth 5, 2 lights): the WHOLE frame.
* C4 (2 M triangles, 1920x1080): the WHOLE frame (~30 minutes of 8 cores).
* C5 (the C3 scene at 3840x2160, 4 lights, depth 8): the WHOLE frame (~2 hours of 8 cores).
(A config can also be sampled: set its BANDS entry to a number of full-width bands of 8 rows.)

Per config it writes tests/golden/full_<cfg>.npz with what the reference returned through its public API
(RayTrace(WorkChunk*) with output_debug, mythtracer.cc:280-312,24-36): the RGB24 bytes, the per-pixel
debug_line_no plane, sha256 of both and of the hit points, the rows covered, and sha256 of the generated
OBJ / MTL files (the scene generator is seeded; the test refuses to compare against another scene).
The reference sources do not travel to the GPU box; these fixtures do.  Takes ~10-60 minutes of CPU.
"""
import hashlib
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from mythtracer_b200 import scenegen  # noqa: E402
from oracle import oracle_py  # noqa: E402

SCENE_DIR = os.environ.get("MTB_SCENE_DIR", "/tmp/mtb_scenes")
BAND_ROWS = 8
# (number of bands of 8 rows) per config; None = the whole frame
BANDS = {"C2": None, "C3": None, "C4": None, "C5": None}


def sha(a) -> str:
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def file_sha(path) -> str:
    h = hashlib.sha256()
    with open(path, "rb") as f:
        for blk in iter(lambda: f.read(1 << 20), b""):
            h.update(blk)
    return h.hexdigest()


def band_rows(height, n_bands):
    """First rows of n_bands bands of BAND_ROWS rows, spread over the frame, aligned to 8 rows."""
    ys = []
    for i in range(n_bands):
        y = int((i + 0.5) * height / n_bands)
        ys.append(min(height - BAND_ROWS, (y // BAND_ROWS) * BAND_ROWS))
    return ys


def make(name):
    files, cfg = scenegen.generate_config(name, SCENE_DIR)
    W, H, depth = cfg["width"], cfg["height"], cfg["depth"]
    ref = oracle_py.Reference(files.obj_path)
    ref.set_lights(files.lights)
    n_bands = BANDS[name]
    if n_bands is None:
        step = 24
        starts = list(range(0, H, step))
        chunks = [(0, y, W, min(step, H - y)) for y in starts]
    else:
        chunks = [(0, y, W, BAND_ROWS) for y in band_rows(H, n_bands)]
    rows, rgb, line_no, points = [], [], [], []
    t0 = time.time()
    for k, c in enumerate(chunks):
        r = ref.render(files.camera, W, H, chunk=c, depth=depth, debug=True)
        rows.extend(range(c[1], c[1] + c[3]))
        rgb.append(r["rgb"])
        line_no.append(r["line_no"])
        points.append(r["points"])
        print("%s chunk %d/%d rows %d..%d  %.1f s  (elapsed %.0f s)" % (name, k + 1, len(chunks), c[1], c[1] + c[3] - 1,
                                                                       r["seconds"], time.time() - t0), flush=True)
    rgb = np.concatenate(rgb, 0)
    line_no = np.concatenate(line_no, 0)
    points = np.concatenate(points, 0)
    out = os.path.join(HERE, "full_%s.npz" % name)
    np.savez_compressed(out, config=name, width=W, height=H, depth=depth, camera=np.array(files.camera),
                        lights=np.array(files.lights), rows=np.array(rows, np.int32), rgb=rgb, line_no=line_no,
                        rgb_sha256=sha(rgb), line_no_sha256=sha(line_no), points_sha256=sha(points),
                        obj_sha256=file_sha(files.obj_path), mtl_sha256=file_sha(files.mtl_path),
                        n_triangles=files.n_triangles, threads=ref.threads(), seconds=time.time() - t0)
    print(name, "->", out, "%d rows, %.1f%% of the frame, %.0f s, %d bytes" % (
        len(rows), 100.0 * len(rows) / H, time.time() - t0, os.path.getsize(out)), flush=True)


if __name__ == "__main__":
    assert os.path.isdir("/root/reference/VerStarting"), "golden vectors are made where the reference is mounted"
    for n in (sys.argv[1:] or ["C2", "C3", "C5", "C4"]):
        if n.endswith(":full"):  # e.g. C4:full - the whole frame instead of bands
            n = n[:-5]
            BANDS[n] = None
        make(n)
