"""Profiling aid (not a test): cycles per tcgen05.mma (M = 128, K = 16, kind::f16) by operand source and N, from
kpf_umma_split_selftest (one CTA, one issuing thread, 3 MMAs per k-step).  python profiles/probe_mma.py"""
import os, sys, torch
sys.path.insert(0, os.getcwd())
from keypointfusion_b200 import ops
dev = "cuda"
K = 128
for a_tmem in (0, 1):
    for N in (16, 32, 64, 128, 256):
        if (128 + N) * K * 4 > 220 * 1024: continue
        A = torch.randn(128, K, device=dev); B = torch.randn(N, K, device=dev); D = torch.empty(128, N, device=dev)
        cyc = torch.zeros(2, dtype=torch.int64, device=dev)
        ops._call("kpf_umma_split_selftest", ops._p(A), ops._p(B), ops._p(D), N, K, 0, a_tmem, 0, ops._p(cyc))
        torch.cuda.synchronize()
        c1, c8 = cyc.tolist()
        n = 3 * K // 16
        print("A from %s  N=%3d : 1 GEMM (%d MMAs) %5d cycles, 8 GEMMs %6d -> %.1f cycles/MMA steady, err %.1e" %
              ("tmem" if a_tmem else "smem", N, n, c1, c8, (c8 - c1) / (7 * n), float((D - A @ B.T).abs().max())))
