import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np
from mobiclipdecoder_b200 import MobiclipDecoder, MobiParser
from mobiclipdecoder_b200.workloads import CONFIGS, frames
w,h,ver,_ = CONFIGS['moflex_400x240']
fr = [d for d,_ in frames('moflex_400x240', 1, 400)]
p = MobiParser(w,h,ver)
t=time.perf_counter()
for f in fr: p.parse(f)
print('parse only: %.1f us/frame' % ((time.perf_counter()-t)/len(fr)*1e6))
dec = MobiclipDecoder(w,h,ver)
for f in fr[:20]:
    dec.Data, dec.Offset = f, 0; dec.DecodeFrame()
a=b=0.0
out = np.empty((h, w, 4), dtype=np.uint8)
import ctypes as C
for f in fr[20:]:
    dec.Data, dec.Offset = f, 0
    t0=time.perf_counter(); dec.DecodeFrame(False); t1=time.perf_counter()
    dec._lib.mobi_read_bgra(dec._h, out.ctypes.data_as(C.c_void_p), w*4); t2=time.perf_counter()
    a+=t1-t0; b+=t2-t1
n=len(fr)-20
print('decode_frame (parse+pack+H2D+launch, async): %.1f us; read_bgra (convert+D2H+sync): %.1f us' % (a/n*1e6, b/n*1e6))
