import sys,time,os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from elynx_b200 import examine as ex
T,S=500,1_000_000
a=torch.randint(0,20,(T,S),dtype=torch.uint8,device='cuda')
for it in range(3):
    torch.cuda.synchronize(); t0=time.perf_counter()
    r=ex.examine_device(a.data_ptr(),S,T,S,20)
    dt=time.perf_counter()-t0
    print("protein examine %dx%d: %.1f ms, %.3g pair-sites/s, hamming mean %.4f (expect 0.95)"%(T,S,dt*1e3,T*(T-1)/2*S/dt,r['hamming_mean']))
