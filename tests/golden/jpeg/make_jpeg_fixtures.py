"""Makes the JPEG / interlaced-PNG fixtures of tests/test_host_logic.py::test_jpeg_and_interlaced_png_textures.

Run in the build container (needs Pillow, which wraps libjpeg-turbo: the library whose default arithmetic
csrc/jpeg_decode.cc restates).  Writes small files plus expected.npz = the RGB bytes Pillow decodes from each of them.
    python tests/golden/jpeg/make_jpeg_fixtures.py
"""
import io
import os

import numpy as np
from PIL import Image

HERE = os.path.dirname(os.path.abspath(__file__))
rng = np.random.default_rng(7)


def picture(h, w):
    """Smooth gradients + sharp coloured edges + noise: exercises every AC coefficient and the chroma filters."""
    y, x = np.mgrid[0:h, 0:w]
    img = np.stack([128 + 100 * np.sin(x / 5.0) * np.cos(y / 7.0), 255.0 * x / max(w - 1, 1), 255.0 * y / max(h - 1, 1)], -1)
    img[h // 4: h // 2, w // 3: 2 * w // 3] = (250, 5, 5)
    img[h // 2:, : w // 4] = (5, 5, 250)
    img += rng.normal(0, 12, img.shape)
    return np.clip(img, 0, 255).astype(np.uint8)


cases = {
    "q90_444.jpg": dict(size=(37, 53), quality=90, subsampling="4:4:4"),
    "q75_420.jpg": dict(size=(45, 61), quality=75, subsampling="4:2:0"),
    "q60_422.jpg": dict(size=(33, 50), quality=60, subsampling="4:2:2"),
    "q30_420_big.jpg": dict(size=(130, 171), quality=30, subsampling="4:2:0"),
    "q95_420_narrow.jpg": dict(size=(40, 3), quality=95, subsampling="4:2:0"),   # chroma plane 2 samples wide: no fancy filter
    "q85_420_restart.jpg": dict(size=(64, 80), quality=85, subsampling="4:2:0", restart_marker_blocks=3),
    "q80_grey.jpg": dict(size=(29, 41), quality=80, grey=True),
    "q85_420_progressive.jpg": dict(size=(70, 90), quality=85, subsampling="4:2:0", progressive=True),
    "q50_444_progressive.jpg": dict(size=(41, 47), quality=50, subsampling="4:4:4", progressive=True),
    "q100_16x16.jpg": dict(size=(16, 16), quality=100, subsampling="4:2:0"),
    "q70_420_optimized.jpg": dict(size=(57, 75), quality=70, subsampling="4:2:0", optimize=True),
}
expected = {}
for name, c in cases.items():
    h, w = c["size"]
    img = picture(h, w)
    im = Image.fromarray(img[..., 0] if c.get("grey") else img)
    kw = {k: v for k, v in c.items() if k not in ("size", "grey")}
    buf = io.BytesIO()
    try:
        im.save(buf, "JPEG", **kw)
    except TypeError:
        kw.pop("restart_marker_blocks", None)
        im.save(buf, "JPEG", **kw)
    data = buf.getvalue()
    open(os.path.join(HERE, name), "wb").write(data)
    expected[name] = np.asarray(Image.open(io.BytesIO(data)).convert("RGB"))
np.savez_compressed(os.path.join(HERE, "expected.npz"), **expected)
print({k: (v.shape, os.path.getsize(os.path.join(HERE, k))) for k, v in expected.items()})
