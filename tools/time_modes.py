import os,sys,torch
ROOT=os.getcwd(); sys.path.insert(0,ROOT); sys.path.insert(0,os.path.join(ROOT,"esp32-wake-word_b200"))
import bench, ww_b200
n=262144
pcm=bench.synth_pcm(n,torch.device("cuda",0),1234)
pf=pcm.float()/32768
def timed(fn):
    for _ in range(3): fn()
    torch.cuda.synchronize(); e0=torch.cuda.Event(enable_timing=True); e1=torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5): fn()
    e1.record(); torch.cuda.synchronize(); return e0.elapsed_time(e1)/5e3
for name,fn in (("py int16",lambda: ww_b200.mfcc_batch(pcm)),("esp int16",lambda: ww_b200.mfcc_batch(pcm,mode="esp",layout="frame_major")),("esp fp32",lambda: ww_b200.mfcc_batch(pf,mode="esp",layout="frame_major")),("py fp32",lambda: ww_b200.mfcc_batch(pf))):
    dt=timed(fn); print(name, round(n/dt/1e6,2),"M clips/s")
