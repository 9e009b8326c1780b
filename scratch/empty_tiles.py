import torch, numpy as np, sys
sys.path.insert(0, '/root/repo')
from moving_object_detector_b200 import Context, ClustererConfig
from moving_object_detector_b200.api import OBJECT_DTYPE, STATS_DTYPE
W,H,B=1280,720,63
dev='cuda:0'
pts=torch.zeros((B,W*H,8),dtype=torch.float32,device=dev)
mask=torch.zeros((B,W*H),dtype=torch.uint8,device=dev)
objs=torch.zeros((B,64*OBJECT_DTYPE.itemsize),dtype=torch.uint8,device=dev)
stats=torch.zeros((B,STATS_DTYPE.itemsize),dtype=torch.uint8,device=dev)
with Context(W,H,max_batch=B) as c:
    s=torch.cuda.current_stream().cuda_stream
    for frac in (0.0,):
        for _ in range(3): c.cluster_batched_dev(W,H,B,pts,mask,ClustererConfig(),objs,stats,64,stream=s)
        a,b=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(20): c.cluster_batched_dev(W,H,B,pts,mask,ClustererConfig(),objs,stats,64,stream=s)
        b.record(); torch.cuda.synchronize()
        print('all-empty clusterer step: %.1f us' % (a.elapsed_time(b)/20*1000))
