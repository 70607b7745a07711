import sys, os; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__)))); sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'tests'))
"""Traversal statistics per ray type on C2 (instrumented kernels): primary, first shadow, first bounce."""
import numpy as np
from computational_ray_tracer_b200 import api, scenes
import common
ctx = api.Context(0)
m = scenes.heightfield(708); ms = api.MeshSet(m); oc = api.Octtree_Model(ms)
sc = api.Scene(ctx); mm = scenes.c2_materials(sc); sc.set_model(oc, mesh_materials=mm); sc.commit()
w, h = 1920, 1080
r2c, c2w = common.camera_1080p_like(w, h)
film = api.Film(ctx, w, h)
def run(mode, depth, tm):
    st = sc.render(film, api.make_config(w, h, r2c, c2w, mode=mode, xs=8, ys=8, spp_begin=0, spp_end=1, max_depth=depth, trace_mode=tm, collect_stats=1))
    return st
for tm in (0, 1):
    a = run(0, 0, tm)                 # primary only
    b = run(1, 0, tm)                 # primary + shadow of bounce 0
    c = run(1, 1, tm)                 # + closest of bounce 1 + its shadow
    def per(x, rays): return {k: round(x[k] / max(rays, 1), 2) for k in ("nodes_visited", "tris_tested", "leaves_visited")}
    pr = a["closest_rays"]
    sh0 = b["shadow_rays"]
    d = {k: b[k] - a[k] for k in ("nodes_visited", "tris_tested", "leaves_visited")}
    print("trace_mode", tm)
    print("  primary rays", pr, per(a, pr))
    print("  shadow rays (bounce 0)", sh0, per(d, sh0))
    cl1 = c["closest_rays"] - b["closest_rays"]; sh1 = c["shadow_rays"] - b["shadow_rays"]
    e = {k: c[k] - b[k] for k in ("nodes_visited", "tris_tested", "leaves_visited")}
    print("  bounce-1 closest + its shadows", cl1, sh1, per(e, cl1 + sh1))
