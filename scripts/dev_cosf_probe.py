import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from tes_b200 import _lib
from tes_b200.device import ptr, stream_ptr
out = torch.zeros(1, dtype=torch.float64, device="cuda")
L = _lib.lib()
for which in (0, 1):
  for hi in (1.5707964, 3.1415927, 64.0, 512.0, 4096.0):
    for e1 in (0.0, 3e-8, 1.0e-7, 1.7e-7):
        L.tes_cosf_error_probe(0.0, hi, e1, which, ptr(out), stream_ptr())
        print("which=%d hi=%-10g e1=%.1e  max(err - e1|x|) = %.3e" % (which, hi, e1, float(out.cpu()[0])))
