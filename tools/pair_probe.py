"""Two rays per lane: does it pay?  Shadow-like rays of the C3 frame (primary hit point -> light 0 / light 1, all
pixels) through mtb_intersect_rays: one ray per thread; two rays per thread in lock step (MTB_FLAG_PAIR_RAYS, Trace2);
two rays per thread by two ordinary calls; two rays per thread back to back inside one node loop (MTB_FLAG_CHAIN_RAYS,
TraceChain).  Results must be identical; kernel time is what is compared."""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from mythtracer_b200 import MythTracer, Light, scenegen, MTB_FLAG_MEGAKERNEL, MTB_FLAG_PAIR_RAYS, MTB_FLAG_CHAIN_RAYS, MTB_FLAG_COUNT_WORK

name = sys.argv[1] if len(sys.argv) > 1 else "C3"
files, cfg = scenegen.generate_config(name, "/tmp/mtb_scenes")
mt = MythTracer(max_depth=cfg["depth"], flags=MTB_FLAG_MEGAKERNEL)
assert mt.LoadObj(files.obj_path)
mt.GetScene().lights = [Light.from_tuple(l) for l in files.lights]
w, h = cfg["width"], cfg["height"]
r = mt.render_chunk(files.camera, w, h, 0, 0, w, h, debug=True)
pts = r["points"].reshape(-1, 3)
hit = r["line_no"].reshape(-1) >= 0
# pixels in 8x4 groups like a warp of the megakernel
idx = np.arange(w * h).reshape(h // 4, 4, w // 8, 8).transpose(0, 2, 1, 3).reshape(-1)
idx = idx[hit[idx]]
idx = idx[: (len(idx) // 32) * 32]
P = pts[idx]
lights = [np.array(l[0:3], float) for l in files.lights[:2]]
rays = []
for L in lights:
    d = L - P
    d /= np.sqrt((d * d).sum(1))[:, None]
    rays.append((P + d * 0.00001, d))
n = len(P)
# one ray per thread: a warp walks 32 rays towards light 0, the next warp the same pixels towards light 1
o1 = np.stack([rays[0][0].reshape(-1, 32, 3), rays[1][0].reshape(-1, 32, 3)], 1).reshape(-1, 3)
d1 = np.stack([rays[0][1].reshape(-1, 32, 3), rays[1][1].reshape(-1, 32, 3)], 1).reshape(-1, 3)
# two rays per thread: thread i walks pixel i's two rays
o2 = np.stack([rays[0][0], rays[1][0]], 1).reshape(-1, 3)
d2 = np.stack([rays[0][1], rays[1][1]], 1).reshape(-1, 3)
out = {"config": name, "pixels": int(n), "rays": int(2 * n)}
res = {}
for label, flags, o, d in (("single", MTB_FLAG_MEGAKERNEL, o1, d1), ("pair", MTB_FLAG_MEGAKERNEL | MTB_FLAG_PAIR_RAYS, o2, d2),
                           ("two_calls", MTB_FLAG_MEGAKERNEL | MTB_FLAG_PAIR_RAYS | MTB_FLAG_CHAIN_RAYS, o2, d2),
                           ("chain", MTB_FLAG_MEGAKERNEL | MTB_FLAG_CHAIN_RAYS, o2, d2)):
    mt.set_flags(flags)
    ms = []
    for it in range(5):
        g = mt.intersect_rays(o, d, want_stats=True)
        ms.append(round(g["stats"]["kernel_ms"], 3))
    res[label] = g
    mt.set_flags(flags | MTB_FLAG_COUNT_WORK)
    c = mt.intersect_rays(o, d, want_stats=True)["stats"]
    out[label] = {"kernel_ms": ms, "best_ms": min(ms), "mrays_s": round(2 * n / min(ms) / 1e3, 1), "bvh_per_ray": round(c["n_bvh"] / (2 * n), 1),
                  "fallback": c["n_fallback"], "fast": c["n_fast"]}
# same answers: un-permute both to (pixel, light)
t1 = res["single"]["t"].reshape(-1, 2, 32).transpose(0, 2, 1).reshape(-1, 2)
t2 = res["pair"]["t"].reshape(-1, 2)
k1 = res["single"]["tri"].reshape(-1, 2, 32).transpose(0, 2, 1).reshape(-1, 2)
k2 = res["pair"]["tri"].reshape(-1, 2)
out["identical"] = bool(np.array_equal(k1, k2) and np.array_equal(t1, t2, equal_nan=True))
for label in ("two_calls", "chain"):
    out["identical"] = out["identical"] and bool(np.array_equal(k1, res[label]["tri"].reshape(-1, 2)) and
                                                 np.array_equal(t1, res[label]["t"].reshape(-1, 2), equal_nan=True))
out["speedup"] = round(out["single"]["best_ms"] / out["pair"]["best_ms"], 3)
out["chain_vs_two_calls"] = round(out["two_calls"]["best_ms"] / out["chain"]["best_ms"], 3)
print(json.dumps(out))
os.makedirs("gpurun_out", exist_ok=True)
json.dump(out, open("gpurun_out/pair_probe_%s.json" % name, "w"), indent=1)
