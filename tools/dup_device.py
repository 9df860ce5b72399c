"""Experiment: one GPU listed k times in a context = k concurrent sub-frames (interleaved strips) on one device."""
import sys, os, json, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from mythtracer_b200 import MythTracer, Light, scenegen, MTB_FLAG_WAVEFRONT
files, cfg = scenegen.generate_config("C3", "/tmp/mtb_scenes")
W, H = cfg["width"], cfg["height"]
ref = None
for world in (1, 8):
    for k in (1, 2, 3, 4):
        mt = MythTracer(devices=[0] * k, max_depth=cfg["depth"], flags=MTB_FLAG_WAVEFRONT)
        assert mt.LoadObj(files.obj_path)
        mt.GetScene().lights = [Light.from_tuple(l) for l in files.lights]
        mt.set_partition(0, world)
        out = torch.zeros((H, W, 3), dtype=torch.uint8).pin_memory().numpy()
        ts = []
        for it in range(7):
            t0 = time.perf_counter(); r = mt.render_chunk(files.camera, W, H, 0, 0, W, H, out=out); ts.append((time.perf_counter() - t0) * 1e3)
        if world == 1:
            if ref is None: ref = out.copy()
            same = bool(np.array_equal(out, ref))
        else:
            same = None
        print(json.dumps(dict(share="1/%d" % world, contexts_on_gpu=k, wall_ms=round(min(ts[3:]), 2), kernel_ms=round(r["stats"]["kernel_ms"], 2), identical=same)), flush=True)
        mt.close()
