import sys, os; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__)))); sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'tests'))
import numpy as np
from computational_ray_tracer_b200 import api, scenes
import common
ctx=api.Context(0)
m=scenes.heightfield(708); ms=api.MeshSet(m); oc=api.Octtree_Model(ms)
sc=api.Scene(ctx); mm=scenes.c2_materials(sc); sc.set_model(oc, mesh_materials=mm); sc.commit()
w,h=1920,1080
r2c,c2w=common.camera_1080p_like(w,h)
film=api.Film(ctx,w,h)
for mode in (0,1):
  for tm in (0,1,2):
    st=sc.render(film, api.make_config(w,h,r2c,c2w,mode=mode,xs=8,ys=8,spp_begin=0,spp_end=2,max_depth=5,trace_mode=tm,collect_stats=1))
    rays=st['closest_rays']+st['shadow_rays']
    print('mode',mode,'tm',tm,'rays',rays,'nodes/ray',st['nodes_visited']/rays,'tris/ray',st['tris_tested']/rays,'leaves/ray',st['leaves_visited']/rays,'maxq',st['max_queue'],'retraced',st['exact_retraced_rays'], 'ms', st['total_ms'])
