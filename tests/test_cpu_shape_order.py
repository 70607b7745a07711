"""The rule behind the shape hierarchy and the pooled shape tests, checked against the oracle on the host.

Scene::Closest (oracle_render.cpp:38-93) walks the analytic shapes in list order after the mesh, each with the current tMax and a strict
'<'.  The device visits the shapes through a hierarchy, or as (ray, shape) pairs pooled per warp, in an order that has nothing to do with
the list; both rest on the claim that the list-order loop returns

    the lexicographically smallest (t_s, s) over the shapes whose UNBOUNDED first valid root t_s is strictly nearer than the mesh hit,

because BasicIntersect(ray, tMax) is the shape's first valid root truncated at tMax.  Here that claim is tested directly: every shape is
intersected on its own with tMax = FLT_MAX, the rule is evaluated in numpy, and the result must be the oracle's Scene::Closest for every
ray -- with partial sweeps (where the first valid root is the far one), duplicated shapes (exact ties: the earlier one must win), a disk in
the plane of the mesh (within a few ulp of its distance: strict '<' decides) and rays that start inside shapes."""
import numpy as np

import oracle_lib as O
from computational_ray_tracer_b200 import scenes

FLT_MAX = np.finfo(np.float32).max


def _scene():
    rng = np.random.default_rng(11)
    orc = O.OracleScene()
    # a wall behind everything: rays that miss every shape still hit the mesh
    orc.set_model([scenes.quad_mesh((-400, -400, 900), (400, -400, 900), (-400, 400, 900), (400, 400, 900), (0, 0, -1))])
    orc.build_octree()
    shapes = []
    centres = []

    def add(kind, rigid, params):
        shapes.append((kind, rigid, params))
        return orc.add_shape(kind, rigid, params)

    for _ in range(14):
        c = rng.uniform(-150, 150, 3) + np.array([0, 0, 600.0])
        r = float(rng.uniform(15, 60))
        kind = int(rng.integers(0, 3))
        centres.append(c)
        if kind == 0:            # sphere, some clipped in z and swept partially in phi (the far root becomes the first valid one)
            zc = float(rng.uniform(0.2, 1.0))
            add(0, scenes.translation(*c), [r, -r * zc, r * zc, float(rng.choice([360.0, 200.0, 90.0]))])
        elif kind == 1:
            add(1, scenes.translation(*c), [r, -r, r, float(rng.choice([360.0, 180.0]))])
        else:
            add(2, scenes.translation(*c), [0.0, 0.0 if rng.random() < 0.5 else r * 0.3, r, float(rng.choice([360.0, 270.0]))])
    # exact duplicates, later in the list: they tie with their originals at every ray and must lose
    for k in (0, 3, 5):
        add(*shapes[k])
    # a disk in the plane of the wall: its t lands within a few ulp of the mesh hit's
    add(2, scenes.translation(0, 0, 900), [0.0, 0.0, 120.0, 360.0])
    add(3, np.eye(4, dtype=np.float32), [-80, -60, 500, 90, -60, 520, 0, 100, 480])
    return orc, len(shapes), np.array(centres)


def _rays(n, seed, centres):
    rng = np.random.default_rng(seed)
    o = np.zeros((n, 3)); o[:, 2] = rng.uniform(-50, 50, n)
    tgt = rng.uniform(-220, 220, (n, 3)) + np.array([0, 0, 620.0])
    aimed = rng.random(n) < 0.7                                     # most rays are aimed at a shape (through or just past it)
    tgt[aimed] = centres[rng.integers(0, len(centres), aimed.sum())] + rng.normal(0, 25.0, (aimed.sum(), 3))
    inside = rng.random(n) < 0.25                                   # a quarter of the rays start in the middle of the shapes
    o[inside] = rng.uniform(-120, 120, (inside.sum(), 3)) + np.array([0, 0, 600.0])
    d = tgt - o
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    return np.concatenate([o, d], 1).astype(np.float32)


def test_list_order_loop_equals_the_lexicographic_minimum(oracle):
    orc, n_shapes, centres = _scene()
    rays = _rays(6000, 3, centres)
    want = orc.scene_closest(rays)
    mesh = orc.trace(rays)
    t_mesh = np.where(mesh["tri"] >= 0, mesh["t"], np.float32(FLT_MAX)).astype(np.float32)
    best_t = t_mesh.copy()
    best_s = np.full(len(rays), -1, np.int64)
    for s in range(n_shapes):                                        # any order would do: this is the order-free statement of the rule
        r = orc.shape_intersect(s, rays, tmax=FLT_MAX)
        ok = (r["found"] != 0) & (r["t"] >= 0)
        better = ok & ((r["t"] < best_t) | ((r["t"] == best_t) & (best_s >= 0) & (s < best_s)))
        best_t = np.where(better, r["t"], best_t).astype(np.float32)
        best_s = np.where(better, s, best_s)
    shape_wins = best_s >= 0
    assert shape_wins.sum() > 1500 and (~shape_wins).sum() > 500
    assert np.array_equal(want["kind"] == 1, shape_wins)
    assert np.array_equal(want["id0"][shape_wins], best_s[shape_wins])
    assert np.array_equal(want["t"][shape_wins].view(np.uint32), best_t[shape_wins].view(np.uint32))
    # the cases the rule is there for actually occur: duplicates (ties), the far root of a clipped shape, a shape at the mesh distance
    dup_of = {n_shapes - 5: 0, n_shapes - 4: 3, n_shapes - 3: 5}
    assert not np.isin(want["id0"][shape_wins], list(dup_of)).any()                     # an original always beats its later duplicate
    assert np.isin(want["id0"][shape_wins], list(dup_of.values())).any()
    orc.close()


def test_reversed_test_order_gives_the_same_answer(oracle):
    """The same fold run over the shapes in reverse list order: the (t, index) rule makes the visit order irrelevant."""
    orc, n_shapes, centres = _scene()
    rays = _rays(3000, 5, centres)
    want = orc.scene_closest(rays)
    mesh = orc.trace(rays)
    t_mesh = np.where(mesh["tri"] >= 0, mesh["t"], np.float32(FLT_MAX)).astype(np.float32)
    key_t = t_mesh.copy(); key_s = np.full(len(rays), -1, np.int64)
    for s in reversed(range(n_shapes)):
        r = orc.shape_intersect(s, rays, tmax=FLT_MAX)
        ok = (r["found"] != 0) & (r["t"] >= 0) & (r["t"] < t_mesh)                  # strictly nearer than the mesh hit
        better = ok & ((key_s < 0) | (r["t"] < key_t) | ((r["t"] == key_t) & (s < key_s)))
        key_t = np.where(better, r["t"], key_t).astype(np.float32)
        key_s = np.where(better, s, key_s)
    wins = key_s >= 0
    assert np.array_equal(want["kind"] == 1, wins)
    assert np.array_equal(want["id0"][wins], key_s[wins])
    orc.close()


def test_a_wrapped_angle_never_exceeds_a_full_sweep():
    """The integrator's shape code skips atan2f where the sweep is complete (crt_shapes.cuh clip_phi): the reference's wrapped angle
    (Shapes.h:315-318: atan2, and for a negative result `phi += 2 * pi` with a float += double) can never exceed phimax = 360 deg *
    (pi / 180) evaluated in float, so `phi > phimax` cannot reject.  Checked for every kind of float an atan2f may return."""
    phimax = np.float32(360.0) * np.float32(0.01745329251994329576923690768489)
    assert phimax == np.float32(6.28318548) and phimax == np.float32(2 * np.pi)
    pi_f = np.float32(np.pi)                                                # the largest magnitude atan2f returns
    neg = -np.concatenate([np.float32(2.0) ** np.arange(-149, 2, dtype=np.float32), np.linspace(1e-6, float(pi_f), 200001, dtype=np.float32),
                           np.nextafter(np.float32(0), np.float32(1), dtype=np.float32)[None], pi_f[None]]).astype(np.float32)
    neg = neg[(neg < 0) & (neg >= -pi_f)]
    wrapped = (neg.astype(np.float64) + 2 * 3.141592653589793238462643383279502884).astype(np.float32)
    assert wrapped.max() <= phimax and wrapped.min() > 0
    assert not (wrapped > phimax).any() and pi_f < phimax


def test_the_first_valid_root_does_not_depend_on_tmax(oracle):
    """closest_over_shapes keeps the record of the accepting test instead of intersecting the winner again with tMax = FLT_MAX, and the
    pooled version tests every pair against `nextafter(bound)`: both need BasicIntersect(ray, tMax) to return the SAME root (t and hit
    point, bit for bit) for every tMax that admits it, and nothing for a tMax below it."""
    orc, n_shapes, centres = _scene()
    rays = _rays(4000, 9, centres)
    checked = 0
    for s in range(n_shapes):
        full = orc.shape_intersect(s, rays, tmax=FLT_MAX)
        hit = np.flatnonzero((full["found"] != 0) & (full["t"] > 0))
        for i in hit[:: max(1, len(hit) // 40)]:                      # per-ray tMax: a few dozen rays per shape
            t = full["t"][i]
            above = orc.shape_intersect(s, rays[i:i + 1], tmax=float(np.nextafter(t, np.float32(np.inf))))
            far = orc.shape_intersect(s, rays[i:i + 1], tmax=float(t) * 4.0)
            below = orc.shape_intersect(s, rays[i:i + 1], tmax=float(np.nextafter(t, np.float32(0))))
            for r in (above, far):
                assert r["found"][0] and r["t"][0].view(np.uint32) == t.view(np.uint32)
                assert np.array_equal(r["hitp"][0].view(np.uint32), full["hitp"][i].view(np.uint32))
            assert not below["found"][0]
            checked += 1
    assert checked > 300
    orc.close()
