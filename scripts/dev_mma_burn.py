import sys, os, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from tes_b200 import _lib, ops
from tes_b200.device import ptr, stream_ptr
torch.cuda.set_device(0)
L = _lib.lib()
blocks = 148 * 8
sink = torch.empty(blocks * 256, dtype=torch.float64, device="cuda")
fl = ctypes.c_double(0)
for it in (1 << 12, 1 << 16, 1 << 18):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); L.tes_fp64_mma_burn(blocks, it, ptr(sink), ctypes.byref(fl), stream_ptr()); e1.record(); e1.synchronize()
    print("DMMA m8n8k4 iters=%d: %.2f TFLOP/s" % (it, fl.value / (e0.elapsed_time(e1) * 1e-3) * 1e-12))
print("DFMA %.2f  FFMA2 %.2f" % (ops.fp64_fma_peak(0.5), ops.fp32_fma_peak(0.5)))
