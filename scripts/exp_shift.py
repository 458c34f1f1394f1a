"""Hardware experiment: do UMMA shared-memory descriptors tolerate a start that is 128-byte- but not 1024-byte-aligned
(rows shifted inside a SWIZZLE_128B box), with which base_offset, and with which group strides?"""
import ctypes as C, sys, os
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from streamingflow_b200 import _lib as L
lib = L.load()
torch.manual_seed(0)
n = 64
for rows_a, sbo_rows in ((256, 8), (512, 10), (512, 18), (512, 22), (512, 24)):
    a = torch.randn(rows_a, 64, device="cuda").to(torch.bfloat16)
    b = torch.randn(n, 64, device="cuda").to(torch.bfloat16)
    for shift in (0, 1, 3, 7, 9):
        if shift + 15 * sbo_rows + 8 > rows_a:
            continue
        idx = torch.tensor([shift + (m // 8) * sbo_rows + m % 8 for m in range(128)], device="cuda")
        ref = a[idx].double() @ b.double().t()
        res = []
        for base_mode in (0, 1):
            d = torch.zeros(128, n, device="cuda")
            rc = lib.sf_diag_umma_shift(a.data_ptr(), b.data_ptr(), d.data_ptr(), n, rows_a, shift, sbo_rows, base_mode,
                                        C.c_void_p(torch.cuda.current_stream().cuda_stream))
            torch.cuda.synchronize()
            res.append(((d.double() - ref).abs().max() / ref.abs().max()).item())
        print(f"sbo_rows={sbo_rows:2d} shift={shift:2d}  err(base_offset=0)={res[0]:.2e}  err(base_offset=(addr>>7)&7)={res[1]:.2e}")
