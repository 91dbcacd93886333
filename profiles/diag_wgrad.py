"""exact-mode tc_wgrad: error vs fp64 and time vs the SIMT TN GEMM it replaces (one GPU)"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import hydragnn_b200 as hb
from hydragnn_b200 import ops

def rel(a, b):
    return float((a.double().cpu() - b).norm() / b.norm())

def t_us(fn, n=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3

for m, n, k, shift in [(172032, 64, 128, 0.0), (172032, 64, 64, 0.0), (172032, 128, 128, 0.5), (860160, 64, 64, 0.5), (200000, 192, 96, 0.0), (9920, 64, 64, 0.0)]:
    g = torch.Generator().manual_seed(1)
    dz, x = torch.randn(m, n, generator=g) + shift, torch.randn(m, k, generator=g) + shift
    dzd, xd = dz.cuda(), x.cuda()
    ref = dz.double().t() @ x.double()
    dw, db = ops.raw_tc_wgrad(dzd, xd)
    s = ops.raw_gemm(dzd, xd, True, False)
    with ops.tensor_cores(True):
        d32, _ = ops.raw_tc_wgrad(dzd, xd)
        t32 = t_us(lambda: ops.raw_tc_wgrad(dzd, xd))
    tx = t_us(lambda: ops.raw_tc_wgrad(dzd, xd))
    ts = t_us(lambda: (ops.raw_gemm(dzd, xd, True, False), ops.raw_colsum(dzd)))
    gb = m * (n + k) * 4 / 1e9
    print(f"m={m} n={n} k={k} shift={shift}: err exact {rel(dw, ref):.2e} simt {rel(s, ref):.2e} tf32 {rel(d32, ref):.2e} | us exact {tx:.1f} ({gb/tx*1e6:.0f} GB/s) tf32 {t32:.1f} ({gb/t32*1e6:.0f} GB/s) simt+colsum {ts:.1f}", flush=True)
