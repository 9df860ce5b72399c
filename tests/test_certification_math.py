"""CPU checks of the arithmetic claims the certified fast traversal rests on (device code in
mythtracer_b200/csrc/device_core.cuh).  Nothing here touches the GPU or the oracle: the device formulas are restated
with Python floats (IEEE double, one rounding per operation, no contraction -- what --fmad=false gives) and numpy
float32, and compared with exact rational arithmetic (fractions.Fraction over the same binary inputs).

  * MollerTrumboreBound: |t_computed - t_exact| <= e for every accepted hit, also for sliver / grazing cases
  * FastBox: the FP32 slab test over padded, outward-rounded boxes accepts whenever the exact FP64 pre-test
    (primitive_triangle.cc:85-108) does, and its entry distance is a lower bound of the FP64 one
  * LimitPrune: the worst-case bound M(T) dominates the error bound of every accepted hit whose box is entered at T
"""
import math
import random
from fractions import Fraction

import numpy as np

U49, U50, U48 = 2.0 ** -49, 2.0 ** -50, 2.0 ** -48


def _sub(a, b):
    return (a[0] - b[0], a[1] - b[1], a[2] - b[2])


def _cross(t, a):  # math3d.h:120-126, this.Cross(a)
    return (t[1] * a[2] - t[2] * a[1], t[2] * a[0] - t[0] * a[2], t[0] * a[1] - t[1] * a[0])


def _dot(a, b):
    return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]


def moller_trumbore_bound(v0, v1, v2, o, d):
    """MollerTrumboreBound of device_core.cuh: (accepted, t, e)."""
    e1, e2 = _sub(v1, v0), _sub(v2, v0)
    pvec = _cross(d, e2)
    det = _dot(e1, pvec)
    if -0.00000001 <= det < 0.00000001:
        return False, 0.0, 0.0
    tvec = _sub(o, v0)
    un = _dot(tvec, pvec)
    us = un if det > 0.0 else -un
    if us < -1e-200 or us > abs(det) * 1.0000000001:
        early = True
    else:
        early = False
    inv_det = 1.0 / det
    u = un * inv_det
    rejected_by_u = u < 0.0 or u > 1.0
    assert not early or rejected_by_u, "the division-free reject must agree with the reference's u test"
    if rejected_by_u:
        return False, 0.0, 0.0
    qvec = _cross(tvec, e1)
    v = _dot(d, qvec) * inv_det
    if v < 0.0 or u + v > 1.0:
        return False, 0.0, 0.0
    t = _dot(e2, qvec) * inv_det
    if t < 0.0:
        return False, 0.0, 0.0
    ad, a1, a2, at = [tuple(abs(x) for x in w) for w in (d, e1, e2, tvec)]
    D = a1[0] * (ad[1] * a2[2] + ad[2] * a2[1]) + a1[1] * (ad[2] * a2[0] + ad[0] * a2[2]) + a1[2] * (ad[0] * a2[1] + ad[1] * a2[0])
    N = a2[0] * (at[1] * a1[2] + at[2] * a1[1]) + a2[1] * (at[2] * a1[0] + at[0] * a1[2]) + a2[2] * (at[0] * a1[1] + at[1] * a1[0])
    e_det, e_num = U49 * D, U49 * N
    den = abs(det) - e_det
    e = (e_num + t * e_det) / den + U50 * t if den > 0.0 else math.inf
    return True, t, e


def exact_t(v0, v1, v2, o, d):
    F = lambda w: tuple(Fraction(x) for x in w)
    v0, v1, v2, o, d = F(v0), F(v1), F(v2), F(o), F(d)
    e1, e2 = _sub(v1, v0), _sub(v2, v0)
    det = _dot(e1, _cross(d, e2))
    if det == 0:
        return None
    return _dot(e2, _cross(_sub(o, v0), e1)) / det


def _random_case(rng, kind):
    R = 400.0
    scale = 10.0 ** rng.uniform(-3, 2)
    c = [rng.uniform(-R, R) for _ in range(3)]
    v = [[c[a] + rng.uniform(-scale, scale) for a in range(3)] for _ in range(3)]
    # aim at a point inside the triangle so that most cases are accepted
    w = [rng.random() for _ in range(3)]
    s = sum(w)
    target = [sum(w[k] / s * v[k][a] for k in range(3)) for a in range(3)]
    if kind == "grazing":  # origin almost in the triangle's plane
        e1 = [v[1][a] - v[0][a] for a in range(3)]
        back = 10.0 ** rng.uniform(0, 2.5)
        n1 = math.sqrt(sum(x * x for x in e1)) or 1.0
        o = [target[a] - e1[a] / n1 * back + rng.uniform(-1, 1) * 10.0 ** rng.uniform(-9, -3) for a in range(3)]
    else:
        o = [rng.uniform(-R, R) for _ in range(3)]
    d = [target[a] - o[a] for a in range(3)]
    n = math.sqrt(sum(x * x for x in d)) or 1.0
    stretch = 1.0 if kind != "unnormalised" else 10.0 ** rng.uniform(-1, 1)  # reflected rays are not unit length
    d = [x / n * stretch for x in d]
    return tuple(v[0]), tuple(v[1]), tuple(v[2]), tuple(o), tuple(d)


def test_moller_trumbore_error_bound_holds():
    rng = random.Random(11)
    accepted = finite = 0
    worst = 0.0
    for i in range(9000):
        kind = ("plain", "grazing", "unnormalised")[i % 3]
        v0, v1, v2, o, d = _random_case(rng, kind)
        ok, t, e = moller_trumbore_bound(v0, v1, v2, o, d)
        if not ok:
            continue
        accepted += 1
        if not math.isfinite(e):
            continue  # an infinite bound makes the ray ambiguous: the exact recursion decides
        finite += 1
        tx = exact_t(v0, v1, v2, o, d)
        err = abs(Fraction(t) - tx)
        assert err <= Fraction(e), (kind, t, float(tx), e, float(err))
        if e > 0:
            worst = max(worst, float(err) / e)
    assert accepted > 4000 and finite > 3500
    assert worst < 0.6, "the bound is supposed to have at least ~2x slack (%.3f)" % worst


def test_division_free_reject_agrees_with_the_u_test():
    """Far misses: `us < -1e-200 || us > |det| * (1 + 1e-10)` must imply the reference's `u < 0 || u > 1` (asserted
    inside moller_trumbore_bound), and must fire for most of them; includes denormal-sized numerators."""
    rng = random.Random(3)
    misses = 0
    for i in range(20000):
        v0, v1, v2, o, d = _random_case(rng, "plain")
        # move the ray sideways so that it misses the triangle by up to a few triangle sizes, or scale everything
        # down towards the denormal range every 10th case
        shift = [rng.uniform(-1, 1) * 10.0 ** rng.uniform(-3, 2) for _ in range(3)]
        o = tuple(o[a] + shift[a] for a in range(3))
        if i % 10 == 0:
            s = 10.0 ** rng.uniform(-160, -100)
            v0, v1, v2, o = [tuple(x * s for x in w) for w in (v0, v1, v2, o)]
        ok, t, e = moller_trumbore_bound(v0, v1, v2, o, d)
        misses += 0 if ok else 1
    assert misses > 5000


def _round_out(lo, hi):
    flo, fhi = np.float32(lo), np.float32(hi)
    if float(flo) > lo:
        flo = np.nextafter(flo, np.float32(-np.inf))
    if float(fhi) < hi:
        fhi = np.nextafter(fhi, np.float32(np.inf))
    return flo, fhi


def _fma32(a, b, c):
    # a * b is exact in double (24 + 24 bits); one more rounding to float: within half an ulp32 of the true fma
    return np.float32(float(a) * float(b) + float(c))


def test_fast_box_is_conservative():
    rng = random.Random(5)
    R = 400.0
    pad = 2.0 ** -16 * R * 1.000001  # SceneBvhBuilder::pad: the whole FP32 error, paid in space at build time
    checked = passed64 = 0
    worst = 0.0
    for i in range(20000):
        size = 10.0 ** rng.uniform(-3, 2)
        c = [rng.uniform(-R, R - size) for _ in range(3)]
        lo = [c[a] for a in range(3)]
        hi = [min(R, c[a] + rng.uniform(0, size)) for a in range(3)]
        o = [rng.uniform(-8 * R, 8 * R) if i % 4 == 0 else rng.uniform(-R, R) for _ in range(3)]
        # aim through the box (or near it), with one component made tiny every other case
        tgt = [rng.uniform(lo[a] - 0.01 * size, hi[a] + 0.01 * size) for a in range(3)]
        d = [tgt[a] - o[a] for a in range(3)]
        n = math.sqrt(sum(x * x for x in d)) or 1.0
        d = [x / n for x in d]
        if i % 2 == 0:
            d[rng.randrange(3)] = rng.choice([-1, 1]) * 10.0 ** rng.uniform(-12, -3)
        if any(x == 0.0 for x in d):
            continue
        inv = [1.0 / x for x in d]
        # the exact FP64 pre-test (SlabRegular == primitive_triangle.cc:85-108 for regular rays)
        near = [((hi[a] if inv[a] < 0 else lo[a]) - o[a]) * inv[a] for a in range(3)]
        far = [((lo[a] if inv[a] < 0 else hi[a]) - o[a]) * inv[a] for a in range(3)]
        tmax, tmin = min(far), max(near)
        pass64 = not (tmax < 0.0) and not (tmin > tmax)
        # FastBox on the padded, outward-rounded float box
        of = [np.float32(x) for x in o]
        fi = [np.float32(x) for x in inv]
        no = [-(of[a] * fi[a]) for a in range(3)]
        box = [_round_out(lo[a] - pad, hi[a] + pad) for a in range(3)]
        # near / far per axis = min / max of the two plane distances (FastBox, device_core.cuh); no widening
        a32 = [_fma32(box[a][0], fi[a], no[a]) for a in range(3)]
        b32 = [_fma32(box[a][1], fi[a], no[a]) for a in range(3)]
        tn = max(min(a32[a], b32[a]) for a in range(3))
        tf = min(max(a32[a], b32[a]) for a in range(3))
        pass32 = bool(tf >= 0.0 and tn <= tf)
        checked += 1
        # per plane: the FP32 distance of the padded plane bounds the FP64 distance of the true plane, and the
        # measured error (in space) stays within the documented 2^-18.9 R
        for a in range(3):
            if abs(inv[a]) > 2.0 ** 100 or abs(inv[a]) < 2.0 ** -100:
                continue
            n64, f64 = near[a], far[a]
            assert float(min(a32[a], b32[a])) <= n64 and float(max(a32[a], b32[a])) >= f64, (a, lo, hi, o, d)
            plane_lo = box[a][1] if inv[a] < 0 else box[a][0]
            exact = (Fraction(float(plane_lo)) - Fraction(o[a])) / Fraction(d[a])
            err_space = abs(Fraction(float(min(a32[a], b32[a]))) - exact) * abs(Fraction(d[a]))
            worst = max(worst, float(err_space) / R)
        if pass64:
            passed64 += 1
            assert pass32, (lo, hi, o, d, tmin, tmax, float(tn), float(tf))
            assert float(tn) <= tmin, (float(tn), tmin)
    assert checked > 15000 and passed64 > 5000
    assert worst <= 2.0 ** -18.9, worst  # the error model of FastBox: at most 2^-18.9 R in space; pad = 2^-16 R


def limit_margin(L, dmax, T):
    """M(T) of LimitPrune (the code prunes at t_limit + 2 M)."""
    k = U48 * 6.0 * L * L
    den = 0.00000001 - k * dmax
    if not den > 0.000000005:
        return math.inf
    return k * (2.0 * dmax * T + L) / den + (2.0 ** -49) * T


def test_limit_margin_dominates_the_error_bound():
    rng = random.Random(23)
    n = 0
    worst = 0.0
    for i in range(9000):
        v0, v1, v2, o, d = _random_case(rng, ("plain", "grazing", "unnormalised")[i % 3])
        ok, t, e = moller_trumbore_bound(v0, v1, v2, o, d)
        if not ok or not math.isfinite(e):
            continue
        L = max(max(v0[a], v1[a], v2[a]) - min(v0[a], v1[a], v2[a]) for a in range(3))
        dmax = max(abs(x) for x in d)
        # entry distance of the triangle's box: T <= t (the hit lies inside the box)
        inv = [1.0 / x for x in d]
        lo = [min(v0[a], v1[a], v2[a]) for a in range(3)]
        hi = [max(v0[a], v1[a], v2[a]) for a in range(3)]
        T = max(max(((hi[a] if inv[a] < 0 else lo[a]) - o[a]) * inv[a] for a in range(3)), 0.0)
        M = limit_margin(L, dmax, T)
        if not math.isfinite(M):
            continue
        n += 1
        assert e <= M, (e, M, L, dmax, T, t)
        worst = max(worst, e / M if M > 0 else 0.0)
    assert n > 3000


def test_unambiguous_winner_is_the_recursions_winner(oracle_mod):
    """The certification rule itself, on the CPU: over ALL triangles whose exact reference test accepts a ray, take
    the smallest t; if no other accepted hit's error interval touches the winner's (lo2 > t* + e*), the octree
    recursion of the reference (oracle restatement, pinned bit for bit to the unmodified reference) must return that
    very triangle with that very t.  The scene is built to provoke near-ties: every triangle also exists as a copy
    displaced by 1e-13 .. 1e-9 and as an exact duplicate; those rays must come out ambiguous, not wrong."""
    from mythtracer_b200.api import MTL_DTYPE, TRI_DTYPE
    rng = random.Random(17)
    nrng = np.random.default_rng(17)
    base = np.zeros(60, TRI_DTYPE)
    base["vertex"] = nrng.uniform(-6, 6, (60, 9))
    base["vertex"][:, 2::3] += 10.0
    near = base.copy()
    near["vertex"] += nrng.uniform(-1, 1, (60, 9)) * (10.0 ** nrng.uniform(-13, -9, (60, 1)))
    tris = np.concatenate([base, near, base[:20]])
    tris["material"] = -1
    tris["line_no"] = np.arange(len(tris))
    orc = oracle_mod.Oracle(tris, np.zeros(0, MTL_DTYPE), [])
    n_rays = 1500
    o = nrng.uniform(-3, 3, (n_rays, 3))
    d = nrng.normal(size=(n_rays, 3))
    d[:, 2] = np.abs(d[:, 2]) + 1.5
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    ref = orc.intersect(o, d)
    verts = [tuple(map(tuple, t.reshape(3, 3))) for t in tris["vertex"]]
    boxes = [(tuple(min(v[k][a] for k in range(3)) for a in range(3)), tuple(max(v[k][a] for k in range(3)) for a in range(3))) for v in verts]
    unambiguous = ambiguous = misses = 0
    for r in range(n_rays):
        ro, rd = tuple(o[r]), tuple(d[r])
        inv = [1.0 / x for x in rd]
        best = None
        lo2 = math.inf
        for k, (v0, v1, v2) in enumerate(verts):
            lo, hi = boxes[k]
            # the reference's AABB pre-test (primitive_triangle.cc:85-108), regular rays
            near_t = max(((hi[a] if inv[a] < 0 else lo[a]) - ro[a]) * inv[a] for a in range(3))
            far_t = min(((lo[a] if inv[a] < 0 else hi[a]) - ro[a]) * inv[a] for a in range(3))
            if far_t < 0.0 or near_t > far_t:
                continue
            ok, t, e = moller_trumbore_bound(v0, v1, v2, ro, rd)
            if not ok:
                continue
            if best is not None and not (t < best[0]):
                lo2 = min(lo2, t - e)
                continue
            if best is not None:
                lo2 = min(lo2, best[0] - best[1])
            best = (t, e, k)
        if best is None:
            misses += 1
            assert ref["tri"][r] == -1
            continue
        if lo2 <= best[0] + best[1]:
            ambiguous += 1  # handed to the exact recursion on the device
            continue
        unambiguous += 1
        assert ref["tri"][r] == best[2], (r, int(ref["tri"][r]), best)
        assert ref["t"][r] == best[0]
    assert ambiguous > 30 and unambiguous > 200, (unambiguous, ambiguous, misses)


def test_stated_edge_case_sibling_entry_tie_is_a_property_of_the_reference(oracle_mod, tmp_path):
    """DESIGN.md section 4, stated edge case (a), made concrete.  A regular ray that passes exactly through an edge of
    the octree grid (computed entry distances of three sibling octants are EQUAL): the reference's recursion visits the
    tied siblings in index order, accepts the hit it finds in the first one (t = 4.90) and stops, although a later
    sibling holds a triangle that is hit at t = 3.0.  The unmodified reference and the oracle restatement agree on the
    farther triangle; brute force over all triangles gives the nearer one -- which is also what a closest-hit search
    (the device's default traversal for regular rays) returns.  MTB_FLAG_EXACT_OCTREE walks the recursion itself."""
    from mythtracer_b200.api import MTL_DTYPE, TRI_DTYPE
    tris = []

    def tri(a, b, c):
        tris.append([*a, *b, *c])
    tri((1, 3, 1.0), (3, 1, 1.0), (2.2, 2.2, 3.9))        # octant 0 (x < 4, y < 4): hit at t = 4.90
    tri((4, 4, 2.0), (4, 4, 3.25), (6, 2, 2.5))           # octant 1 (x >= 4, y <= 4): its edge lies ON the grid edge x = y = 4
    tri((0, 0, 0), (0.3, 0, 0), (0, 0.3, 0))              # the root box is [0, 8]^3, centre (4, 4, 4)
    tri((8, 8, 8), (7.7, 8, 8), (8, 7.7, 8))
    rng = np.random.default_rng(1)
    for _ in range(14):                                   # >= 16 primitives: the root splits (octtree.h:43)
        c = np.array([6.5, 1.0, 6.5]) + rng.uniform(-0.4, 0.4, 3)
        tri(c, c + [0.2, 0, 0], c + [0, 0.2, 0])
    arr = np.zeros(len(tris), TRI_DTYPE)
    arr["vertex"] = np.array(tris, float)
    arr["material"] = -1
    arr["line_no"] = np.arange(len(tris))
    orc = oracle_mod.Oracle(arr, np.zeros(0, MTL_DTYPE), [])
    o = np.array([[7.0, 7.0, 3.0]])
    d = np.array([[-1.0, -1.0, -0.125]])                  # crosses x = 4 and y = 4 at t = 3 exactly
    rec = orc.intersect(o, d)
    brute = orc.intersect(o, d, brute=True)
    assert rec["tri"][0] == 0 and abs(rec["t"][0] - 4.904347826086957) < 1e-12
    assert brute["tri"][0] == 1 and brute["t"][0] == 3.0
    if oracle_mod.Reference.available():
        path = tmp_path / "edge.obj"
        with open(path, "w") as f:
            for t in tris:
                for k in range(3):
                    f.write("v %r %r %r\n" % (float(t[3 * k]), float(t[3 * k + 1]), float(t[3 * k + 2])))
            for i in range(len(tris)):
                f.write("f %d %d %d \n" % (3 * i + 1, 3 * i + 2, 3 * i + 3))
        ref = oracle_mod.Reference(str(path)).intersect(o, d)
        assert ref["line_no"][0] == 3 * len(tris) and abs(ref["t"][0] - rec["t"][0]) == 0.0
