"""Hardware probe for the tcgen05 building block (csrc/mp_probe.cu): runs the one-tile GEMM with both
readings of the shared-memory descriptor's LBO/SBO fields and reports which one reproduces A @ W^T.
Run on the GPU box:  python tests/probe_umma.py"""
import ctypes
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import fortattack_b200 as fab  # noqa: E402


def pack_canonical(w):
    """fp16 [N][K] -> canonical K-major core-matrix order [N/8][K/8][8][8] (see csrc/mp_umma.cuh)."""
    n, k = w.shape
    return w.reshape(n // 8, 8, k // 8, 8).permute(0, 2, 1, 3).contiguous()


def run(K, N, lbo, sbo, seed=0):
    L = fab._capi.lib()
    L.mp_probe_gemm.argtypes = [ctypes.c_void_p] * 3 + [ctypes.c_int] * 2 + [ctypes.c_uint32] * 3 + [ctypes.c_void_p] * 2
    g = torch.Generator(device="cpu").manual_seed(seed)
    a = (torch.randn(128, K, generator=g)).half().cuda()
    w = (torch.randn(N, K, generator=g) / K ** 0.5).half().cuda()
    out = torch.full((128, N), float("nan"), device="cuda")
    err = torch.zeros(1, dtype=torch.int32, device="cuda")
    rc = L.mp_probe_gemm(a.data_ptr(), pack_canonical(w).data_ptr(), out.data_ptr(), K, N, lbo, sbo, 0, err.data_ptr(), None)
    torch.cuda.synchronize()
    ref = a.double() @ w.double().t()
    return rc, int(err.item()), float((out.double() - ref).abs().max()), float(ref.abs().max())


if __name__ == "__main__":
    for K, N in ((64, 64), (128, 128), (128, 256), (256, 128), (64, 16)):
        for name, lbo, sbo in (("lbo=128,sbo=K*16", 128, K * 16), ("lbo=K*16,sbo=128", K * 16, 128)):
            print("K=%d N=%d %s -> rc=%d err=%d maxdiff=%.3e (ref max %.2f)" % ((K, N, name) + run(K, N, lbo, sbo)), flush=True)
