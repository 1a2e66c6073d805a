import sys
sys.path.insert(0,'/root/repo')
import torch
from panda3d_b200 import _capi, synth
from panda3d_b200.casefile import load_case
from panda3d_b200.flat import mesh_from_case
torch.cuda.set_device(0)
stream=torch.cuda.Stream()
with torch.cuda.stream(stream):
    ctx=_capi.Context(0, stream.cuda_stream)
    flush=torch.zeros(256<<20,dtype=torch.uint8,device='cuda')
    for name,flat in (("c1",mesh_from_case(load_case('/root/repo/tests/golden/panda_walk4.p3cs'))),("c2r",synth.config_c2("random").flat)):
        dm=_capi.Mesh(ctx,flat); g=_capi.Group(dm,1)
        g.set_palettes(synth.make_palettes(flat.num_transforms,1,3).reshape(1,-1))
        e0,e1=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
        for mode in ("cold","warm","back2back"):
            for _ in range(5): g.animate()
            stream.synchronize()
            tot=0
            if mode=="back2back":
                e0.record(stream)
                for _ in range(50): g.animate()
                e1.record(stream); stream.synchronize(); tot=e0.elapsed_time(e1)/50*20
            else:
                for _ in range(20):
                    if mode=="cold": flush.add_(1)
                    e0.record(stream); g.animate(); e1.record(stream); stream.synchronize(); tot+=e0.elapsed_time(e1)
            print(name,mode,"us/call %.2f"%(tot/20*1e3),flush=True)
        g.close(); dm.close()
    # an empty kernel for scale
    x=torch.zeros(1,device='cuda')
    for mode in ("cold","warm"):
        tot=0
        for _ in range(20):
            if mode=="cold": flush.add_(1)
            e0.record(stream); x.add_(1); e1.record(stream); stream.synchronize(); tot+=e0.elapsed_time(e1)
        print("tiny torch kernel",mode,"us %.2f"%(tot/20*1e3))
