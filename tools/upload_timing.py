import sys, time; sys.path.insert(0,'/root/repo')
import torch, spfem_b200 as sp
from spfem_b200.assembly import AssemblerElement
m=sp.MeshTri(); m.refine(11)
a=AssemblerElement(m, sp.ElementTriP1())
A=a.iasm(lambda du,dv: du[0]*dv[0]+du[1]*dv[1])
eng=a._get_engine()
for _ in range(3): eng.upload_mesh()
torch.cuda.synchronize()
for rep in range(3):
    t=time.perf_counter()
    for _ in range(5): eng.upload_mesh()
    t1=time.perf_counter()-t
    torch.cuda.synchronize(); t2=time.perf_counter()-t
    print('upload issue %.2f ms  done %.2f ms'%(t1/5*1e3,t2/5*1e3))
