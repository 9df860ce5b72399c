"""The reference-style programs of apps/: a mythtracer_worker-compatible TCP front end driven by a fake
master that speaks the reference's wire protocol (network.h:16-26), and the main_local-style animation driver."""
import os
import shutil
import socket
import struct
import subprocess
import threading

import numpy as np
import pytest

from tests import scenes

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def apps(product_lib):
    if shutil.which("g++") is None or shutil.which("make") is None:
        pytest.skip("g++/make not available")
    subprocess.check_call(["make", "-C", os.path.join(ROOT, "apps")], stdout=subprocess.DEVNULL)
    return os.path.join(ROOT, "apps", "_build")


def _recv_all(conn, n):
    buf = b""
    while len(buf) < n:
        part = conn.recv(n - len(buf))
        assert part, "worker closed the connection"
        buf += part
    return buf


def _recv_packet(conn):
    header = _recv_all(conn, 16)
    tag, ident, length = header[:4], header[4:12], struct.unpack("<I", header[12:])[0]
    return tag, ident, _recv_all(conn, length)


def _send_packet(conn, tag, ident, payload):
    conn.sendall(tag + ident.ljust(8, b"\0")[:8] + struct.pack("<I", len(payload)) + payload)


def test_worker_serves_a_reference_master(apps, scene_dir, tmp_path):
    from mythtracer_b200 import Camera, Light, MythTracer
    files, cfg = scenes.config_scene("C1", scene_dir)
    W, H, T = 300, 200, 128          # 128x128 tiles, clipped at the edges (main_net_master.cc:24-25,202-217)
    cam = Camera.from_tuple(files.camera)
    rig = scenes.LIGHT_RIG[:2]
    lights_path = tmp_path / "rig.txt"
    lights_path.write_text("\n".join(" ".join(repr(float(x)) for x in l) for l in rig) + "\n")
    tiles = [(x, y, min(T, W - x), min(T, H - y)) for y in range(0, H, T) for x in range(0, W, T)]

    srv = socket.socket()
    srv.bind(("127.0.0.1", 0))
    srv.listen(1)
    port = srv.getsockname()[1]
    frame = np.zeros((H, W, 3), np.uint8)
    errors = []

    def master():
        try:
            conn, _ = srv.accept()
            tag, ident, payload = _recv_packet(conn)
            assert tag == b"RDY!" and ident.rstrip(b"\0") == b"b200w" and payload == b""
            _send_packet(conn, b"CAMR", b"b200w", cam.Serialize())
            for (x, y, w, h) in tiles:
                _send_packet(conn, b"WORK", b"b200w", struct.pack("<6I", W, H, x, y, w, h))
                tag, ident, payload = _recv_packet(conn)
                assert tag == b"PXLS"
                n = struct.unpack("<I", payload[:4])[0]
                assert n == w * h * 3 == len(payload) - 4
                frame[y:y + h, x:x + w] = np.frombuffer(payload[4:], np.uint8).reshape(h, w, 3)   # BlitWorkChunk
            conn.close()
        except Exception as e:  # surfaced in the main thread
            errors.append(e)

    t = threading.Thread(target=master, daemon=True)
    t.start()
    res = subprocess.run([os.path.join(apps, "mythtracer_worker_b200"), "b200w", "127.0.0.1", "--obj", files.obj_path, "--port",
                          str(port), "--depth", "2", "--lights", str(lights_path), "--max-chunks", str(len(tiles))],
                         capture_output=True, text=True, timeout=300)
    t.join(timeout=30)
    assert not errors, errors
    assert res.returncode == 0, res.stdout[-1500:] + res.stderr[-1500:]
    mt = MythTracer(max_depth=2)
    assert mt.LoadObj(files.obj_path)
    mt.GetScene().lights = [Light.from_tuple(l) for l in rig]
    assert np.array_equal(frame, mt.RayTrace(W, H, cam))


def test_animation_driver_writes_reference_style_frames(apps, scene_dir, tmp_path):
    from mythtracer_b200 import Camera, Light, MythTracer
    files, cfg = scenes.config_scene("C1", scene_dir)
    out_dir = tmp_path / "anim"
    W, H = 160, 90
    res = subprocess.run([os.path.join(apps, "mythtracer_local_b200"), files.obj_path, str(W), str(H), "74", "76", str(out_dir)],
                         capture_output=True, text=True, timeout=300)
    assert res.returncode == 0, res.stdout[-1500:] + res.stderr[-1500:]
    names = sorted(os.listdir(out_dir))
    assert names == ["dump_00074.raw", "dump_00075.raw", "dump_00076.raw"]
    mt = MythTracer(max_depth=5)
    assert mt.LoadObj(files.obj_path)
    rig = [(231.82174, 81.69966, -27.78259, 0.3, 0.3, 0.3, 1.0, 1.0, 1.0, 1.0, 1.0, 1.0)] + [
        (200.0, 80.0, z, 0.0, 0.0, 0.0, 0.3, 0.3, 0.3, 0.3, 0.3, 0.3) for z in (0.0, 80.0, 160.0)]
    mt.GetScene().lights = [Light.from_tuple(l) for l in rig]
    for frame in (74, 76):   # main_local.cc:51,72-76: angle = 2 * frame, yaw = angle + 90
        cam = Camera((300.0, 107.0, 40.0), 30.0, 2.0 * frame + 90, 0.0, 110.0)
        got = np.fromfile(out_dir / ("dump_%05d.raw" % frame), np.uint8).reshape(H, W, 3)
        assert np.array_equal(got, mt.RayTrace(W, H, cam))
