"""Is the fp32 accumulation of tcgen05.mma biased (truncation) or unbiased (round-to-nearest)?  D = A B^T with all-positive
operands through bx_selftest_umma (3xTF32, the same UMMA path as the scoring engines): signed relative error vs K."""
import os, sys, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np, torch
from bayex_b200 import _lib
for K in (32, 128, 512, 2048, 4096):
    rng = np.random.default_rng(K)
    A = rng.uniform(0.5, 1.0, (128, K)).astype(np.float32)
    B = rng.uniform(0.5, 1.0, (128, K)).astype(np.float32)
    dA, dB = torch.from_numpy(A).cuda(), torch.from_numpy(B).cuda()
    dD = torch.zeros((128, 128), device="cuda")
    rc = _lib.lib.bx_selftest_umma(C.c_void_p(dA.data_ptr()), C.c_void_p(dB.data_ptr()), C.c_void_p(dD.data_ptr()), 128, K,
                                   C.c_void_p(torch.cuda.current_stream().cuda_stream))
    torch.cuda.synchronize()
    if rc:
        print(K, "rc", rc, _lib.lib.bx_last_error()); continue
    got = dD.cpu().numpy().astype(np.float64)
    want = A.astype(np.float64) @ B.astype(np.float64).T
    rel = (got - want) / want
    f32 = (A @ B.T).astype(np.float64)
    rel32 = (f32 - want) / want
    print(f"K={K:5d}: tcgen05 3xTF32 signed rel err mean {rel.mean():+.2e} rms {np.sqrt((rel**2).mean()):.2e} | numpy fp32 mean {rel32.mean():+.2e} rms {np.sqrt((rel32**2).mean()):.2e}")
