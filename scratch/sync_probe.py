import os,sys,time
sys.path.insert(0,'/root/repo')
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS","32")
import numpy as np, torch
import litiv_b200 as lv
from bench import make_frames, W,H,C
seq, frames = make_frames(4, 6)
sub = lv.BackgroundSubtractorSuBSENSE(seed=0); sub.initialize(frames[0])
hf=[lv.pinned_empty((H,W,C)) for _ in range(4)]
for a,f in zip(hf,frames): a[...]=f
hm=lv.pinned_empty((H,W))
for j in range(70): sub.apply(hf[j%4], 1.0 if j<50 else 0.0, out=hm)
torch.cuda.synchronize()
t0=time.perf_counter()
for j in range(100): sub.apply(hf[j%4], 0.0, out=hm)
torch.cuda.synchronize(); print("sync apply us/frame", (time.perf_counter()-t0)/100*1e6)
# raw copies
d=torch.empty((H,W*C),dtype=torch.uint8,device='cuda'); ht=torch.from_numpy(hf[0].reshape(H,W*C))
dm=torch.empty((H,W),dtype=torch.uint8,device='cuda'); hmt=torch.from_numpy(hm)
for name,fn in (("h2d 6.2MB", lambda: d.copy_(ht,non_blocking=True)), ("d2h 2MB", lambda: hmt.copy_(dm,non_blocking=True))):
    torch.cuda.synchronize(); t0=time.perf_counter()
    for _ in range(100): fn(); torch.cuda.synchronize()
    print(name, (time.perf_counter()-t0)/100*1e6, "us (incl sync)")
# device apply + sync each frame (no copies)
pitch=(W*C+127)//128*128
df=torch.zeros((H,pitch),dtype=torch.uint8,device='cuda'); df[:,:W*C]=torch.from_numpy(frames[1].reshape(H,W*C)).cuda()
torch.cuda.synchronize(); t0=time.perf_counter()
for _ in range(100):
    sub.apply_device(df.data_ptr(), pitch, dm.data_ptr(), 0.0); sub.flush(); torch.cuda.synchronize()
print("device apply + full sync us/frame", (time.perf_counter()-t0)/100*1e6)
