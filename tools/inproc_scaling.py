"""In-process multi-GPU (one context, N devices, tiles stored straight into device 0's frame): frame time of a
BASELINE config through mtb_render_chunk with host buffers; the 1-device frame is the byte reference."""
import sys, os, json, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from mythtracer_b200 import MythTracer, Light, scenegen
name = sys.argv[1] if len(sys.argv) > 1 else "C3"
files, cfg = scenegen.generate_config(name, "/tmp/mtb_scenes")
W, H = cfg["width"], cfg["height"]
ref = None
counts = [int(x) for x in sys.argv[2].split(",")] if len(sys.argv) > 2 else [1, 2, 4, 8]
for n in [x for x in counts if x <= torch.cuda.device_count()]:
    mt = MythTracer(devices=list(range(n)), max_depth=cfg["depth"])
    assert mt.LoadObj(files.obj_path)
    mt.GetScene().lights = [Light.from_tuple(l) for l in files.lights]
    out = torch.empty((H, W, 3), dtype=torch.uint8).pin_memory().numpy()
    ts = []
    for it in range(20):  # the automatic pipeline choice needs thirteen frames
        t0 = time.perf_counter()
        r = mt.render_chunk(files.camera, W, H, 0, 0, W, H, out=out)
        ts.append((time.perf_counter() - t0) * 1e3)
    if ref is None:
        ref = out.copy()
    print(json.dumps(dict(config=name, devices=n, wall_ms_best=round(min(ts[14:]), 2), kernel_ms=round(r["stats"]["kernel_ms"], 2),
                          rays=r["stats"]["rays"], identical=bool(np.array_equal(out, ref)), pipeline=mt.pipeline_in_use()[0])), flush=True)
    mt.close()
