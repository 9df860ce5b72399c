"""Dumps the per-pixel ray counts of the C3 / C5 frame (development aid: warp-imbalance analysis)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from mythtracer_b200 import MythTracer, Light, scenegen
for name in sys.argv[1].split(","):
    files, cfg = scenegen.generate_config(name, "/tmp/mtb_scenes")
    mt = MythTracer(max_depth=cfg["depth"], flags=16)
    assert mt.LoadObj(files.obj_path)
    mt.GetScene().lights = [Light.from_tuple(l) for l in files.lights]
    w, h = cfg["width"], cfg["height"]
    r = mt.render_chunk(files.camera, w, h, 0, 0, w, h, taps=True)
    os.makedirs("gpurun_out", exist_ok=True)
    np.save("gpurun_out/nrays_%s.npy" % name, r["n_rays"].astype(np.uint16))
    print(name, r["n_rays"].sum(), r["n_rays"].max())
    mt.close()
