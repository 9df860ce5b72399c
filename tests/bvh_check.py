"""Structural check of a scene BVH (host- or device-built), shared by the CPU and the GPU tests."""
import sys

import numpy as np

sys.setrecursionlimit(10000)


def check_scene_bvh(mt, tris, expect_splits=None):
    nodes, depth, order = mt.scene_bvh()
    n = len(tris)
    if n == 0:
        assert len(nodes) == 0
        return
    refs = np.bincount(order, minlength=n)
    assert len(refs) == n and np.all(refs >= 1), "every triangle must be referenced from a leaf"
    assert mt.scene_info()["n_scene_refs"] == len(order)
    if expect_splits is not None:
        assert (refs.max() > 1) == expect_splits
    v = tris["vertex"].reshape(n, 3, 3)
    lo, hi = v.min(axis=1), v.max(axis=1)
    leaf_boxes = {}  # triangle -> list of the (lo, hi) boxes of the leaves that reference it
    max_depth = 0

    def walk(ref, box, d):
        """`box` = the box the parent stores for this child; returns nothing, asserts containment"""
        nonlocal max_depth
        max_depth = max(max_depth, d)
        b_lo, b_hi = box[:3].astype(np.float64), box[3:].astype(np.float64)
        if ref < 0:
            x = (~ref) & 0xFFFFFFFF
            first, count = x >> 3, x & 7
            for t in order[first:first + count]:
                if refs[t] == 1:
                    assert np.all(b_lo <= lo[t]) and np.all(b_hi >= hi[t]), "leaf box must contain its triangle"
                else:
                    assert np.all(b_lo <= hi[t]) and np.all(b_hi >= lo[t]), "a reference's box must touch its triangle"
                    leaf_boxes.setdefault(int(t), []).append((b_lo, b_hi))
            return
        nd = nodes[ref]
        for cbox, child in ((nd["lbox"], int(nd["left"])), (nd["rbox"], int(nd["right"]))):
            if child >= 0 or ((~child) & 7) > 0:
                assert np.all(box[:3] <= cbox[:3]) and np.all(box[3:] >= cbox[3:]), "a child box must lie inside its parent's"
            walk(child, cbox, d + 1)

    root = nodes[0]
    whole = np.concatenate([np.minimum(root["lbox"][:3], root["rbox"][:3]), np.maximum(root["lbox"][3:], root["rbox"][3:])])
    walk(0, whole, 0)
    assert max_depth <= depth + 1
    # split triangles: points all over the triangle must lie in one of its references' boxes
    rng = np.random.default_rng(3)
    w = rng.dirichlet(np.ones(3), 64)
    w = np.concatenate([w, np.eye(3), [[0.5, 0.5, 0.0], [0.0, 0.5, 0.5], [0.5, 0.0, 0.5]]])
    for t, boxes in list(leaf_boxes.items())[:400]:
        pts = w @ v[t]
        b_lo = np.array([b[0] for b in boxes])
        b_hi = np.array([b[1] for b in boxes])
        inside = np.all((pts[:, None, :] >= b_lo[None]) & (pts[:, None, :] <= b_hi[None]), axis=2).any(axis=1)
        assert inside.all(), "the references of a split triangle must cover it"
