import sys, time
sys.path.insert(0,'/root/repo')
from panda3d_b200 import _capi, synth
ctx=_capi.Context(0)
for name,mk in (("c3",synth.config_c3_mesh),("c2r",lambda: synth.config_c2("random")),("c4",synth.config_c4)):
    flat=mk().flat
    t=time.perf_counter(); dm=_capi.Mesh(ctx,flat); ctx.sync(); print(name,"mesh_create s %.3f"%(time.perf_counter()-t), flush=True); dm.close()
