import sys, ctypes as C
sys.path.insert(0, '/root/repo')
import torch
from bo_pr_b200 import _capi as capi
lib = capi.load_library()
torch.zeros(1, device='cuda')
for rep in range(2):
    for mode, name in enumerate(["dmma regs", "dfma", "dmma+dfma", "dmma + 12 LDS.64/chunk", "dmma + 6 LDS.128/chunk", "dmma + 24 LDS.64/chunk"]):
        tf = C.c_double(0)
        rc = lib.prb_debug_fp64_peak(mode, C.byref(tf), None)
        print(rep, name, rc, round(tf.value, 2))
