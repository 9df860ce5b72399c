"""The N > 1 path on CPU: world_size-2 (and 3) gloo runs of the strip partition + gather that bench.py uses
under torchrun.  There is no CPU renderer in the product, so each rank's strips are produced by the oracle
(test infrastructure standing in for the kernels); what is under test is the host-side partition / gather
logic of mythtracer_b200.tiles, which must reproduce the single-process frame byte for byte."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from tests import scenes


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, obj, cam, lights, w, h, depth, out_path):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from mythtracer_b200 import tiles
    from oracle import oracle_py
    orc = oracle_py.Oracle.from_obj(obj)
    orc.set_lights(lights)
    orc.set_threads(2)
    hp = tiles.padded_height(h, world)
    local = torch.full((hp, w, 3), 77, dtype=torch.uint8)   # rows a rank does not own hold garbage
    for s in tiles.owned_strips(h, rank, world):
        y0, y1 = s * 8, min(h, s * 8 + 8)
        part = orc.render(cam, w, h, chunk=(0, y0, w, y1 - y0), depth=depth, debug=False)
        local[y0:y1] = torch.from_numpy(part["rgb"])
    frame = tiles.gather_frame(local, h, w, rank, world, dst=0)
    if rank == 0:
        np.save(out_path, frame.numpy())
    else:
        assert frame is None
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world,h", [(2, 90), (2, 83), (3, 50)])
def test_strip_gather_reproduces_the_frame(oracle_mod, scene_dir, tmp_path, world, h):
    files, cfg = scenes.config_scene("C1", scene_dir)
    w, depth = 120, 2
    out_path = str(tmp_path / "frame.npy")
    mp.spawn(_worker, args=(world, _free_port(), files.obj_path, files.camera, files.lights, w, h, depth, out_path),
             nprocs=world, join=True)
    orc = oracle_mod.Oracle.from_obj(files.obj_path)
    orc.set_lights(files.lights)
    full = orc.render(files.camera, w, h, depth=depth, debug=False)["rgb"]
    got = np.load(out_path)
    assert got.shape == full.shape and np.array_equal(got, full)
