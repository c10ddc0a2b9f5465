#!/usr/bin/env python
"""Measured ceiling of the formulation north_star and BASELINE config 5 name: the 1024-tap complex float32
FIR of one filter-bank channel as a block-Toeplitz GEMM on the tensor cores.

  y[n] = sum_k h[k] x[n + K-1 - k]  (filter/FIRFilter.cpp:295-299, K = 1024)

With the stream cut into blocks of B = K samples, Y_j = T0 X_j + T1 X_{j-1}: two B x B Toeplitz (triangular)
matrices per channel applied to all blocks at once -- ONE real GEMM [2B x 4B] x [4B x nblocks] per channel
(complex arithmetic as 2 x 2 real blocks).  Executed flop per output: 16 K = 16 384 (twice the algorithmic 8 K:
half of T0 / T1 is structural zeros; a hand-written kernel that skips zero 128 x 128 tiles gets that down to
1.125 x).  The GEMM itself is run by cuBLAS (torch.matmul), the vendor's best tcgen05 code for the shape:
whatever a hand-written Toeplitz kernel does, its MMA rate is bounded by this.

Precision: the product must hold 1e-5 of output RMS (north_star).  bf16 operands carry 8 bits, tf32 11:
  bf16 x 1 : fastest, error ~ 3e-3  (fails by 300 x)
  tf32 x 1 : error ~ 4e-4           (fails)
  tf32 x 3 : x = x1 + x2, T = t1 + t2, terms (1,1) (1,2) (2,1)   -> ~ 1e-6   (passes; 3 GEMMs)
  bf16 x 6 : three-way split, terms of order <= 2                 -> ~ 1e-6   (passes; 6 GEMMs)
Prints one JSON line per variant: Gsamples/s on one GPU for the C5 shape (sampled over `--channels` channels),
the error against a float64 evaluation, and the rate of the fused overlap-save kernel (fir_os64_kernel) on the
same channels for comparison.   python tools/probe_toeplitz_gemm.py > gpurun_out/r02_probe_toeplitz_gemm.jsonl
"""
import argparse
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def toeplitz_pair(h: np.ndarray):
    """T0[i, j] = h[i - j] (i >= j), T1[i, j] = h[B + i - j] (i < j) for B = len(h): y_blk = T0 x_blk + T1 x_prev."""
    B = len(h)
    i, j = np.meshgrid(np.arange(B), np.arange(B), indexing="ij")
    d = i - j
    T0 = np.where(d >= 0, h[np.clip(d, 0, B - 1)], 0)
    T1 = np.where(d < 0, h[np.clip(B + d, 0, B - 1)], 0)
    return T0, T1


def real_form(T0, T1):
    """[2B x 4B] real matrix acting on [Xr_j; Xi_j; Xr_{j-1}; Xi_{j-1}]."""
    def blk(T):
        return np.block([[T.real, -T.imag], [T.imag, T.real]])
    return np.concatenate([blk(T0), blk(T1)], axis=1)


def split(x: torch.Tensor, kind: str, parts: int):
    out, r = [], x.clone()
    for _ in range(parts):
        if kind == "bf16":
            p = r.to(torch.bfloat16).to(torch.float32)
        else:   # tf32: keep 10 explicit mantissa bits (round to nearest even on the dropped 13)
            i = r.view(torch.int32)
            i = (i + 0x0FFF + ((i >> 13) & 1)) & ~0x1FFF
            p = i.view(torch.float32)
        out.append(p)
        r = r - p
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--channels", type=int, default=16)
    ap.add_argument("--log2-samples", type=int, default=20)
    ap.add_argument("--reps", type=int, default=5)
    args = ap.parse_args()
    from pothoscomms_b200 import workloads as wl
    dev = torch.device("cuda", 0)
    K = B = 1024
    n = 1 << args.log2_samples
    nblk = n // B
    C = args.channels
    A = torch.empty((C, 2 * B, 4 * B), dtype=torch.float32, device=dev)
    for c in range(C):
        h = wl.bank_taps(c, 1024, K)
        A[c] = torch.from_numpy(real_form(*toeplitz_pair(h)).astype(np.float32)).to(dev)
    # stream per channel: [K-1 history | n samples]; the block matrix needs a whole extra block in front
    x = torch.stack([wl.tone_noise_torch(1, (nblk + 1) * B, 0xC0FFEE05 + c, dev) for c in range(C)])     # [C, (nblk+1) B, 2]
    Xb = x.view(C, nblk + 1, B, 2)
    # operand [4B x nblk]: rows (Xr_j, Xi_j, Xr_{j-1}, Xi_{j-1}), built once (a real kernel would read the stream in place)
    cur, prev = Xb[:, 1:], Xb[:, :-1]
    R = torch.cat([cur[..., 0], cur[..., 1], prev[..., 0], prev[..., 1]], dim=2).transpose(1, 2).contiguous()   # [C, 4B, nblk]
    # float64 truth for channel 0 (the same block algebra in double)
    truth = (A[0].double() @ R[0].double())                                                               # [2B, nblk]
    t_rms = truth.pow(2).mean().sqrt().item()

    def timed(fn):
        fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(args.reps):
            y = fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / args.reps, y

    variants = []
    torch.backends.cuda.matmul.allow_tf32 = False
    variants.append(("fp32 (cuBLAS, no tensor cores)", 1, lambda: torch.bmm(A, R)))
    Ab, Rb = A.to(torch.bfloat16), R.to(torch.bfloat16)
    variants.append(("bf16 x 1", 1, lambda: torch.bmm(Ab, Rb, out_dtype=torch.float32)))
    torch.backends.cuda.matmul.allow_tf32 = True
    variants.append(("tf32 x 1", 1, lambda: torch.bmm(A, R)))
    A1, A2 = split(A, "tf32", 2)
    R1, R2 = split(R, "tf32", 2)
    variants.append(("tf32 x 3 (two-way split)", 3, lambda: torch.bmm(A1, R1) + torch.bmm(A1, R2) + torch.bmm(A2, R1)))
    Ab3 = [p.to(torch.bfloat16) for p in split(A, "bf16", 3)]
    Rb3 = [p.to(torch.bfloat16) for p in split(R, "bf16", 3)]

    def bf16x6():
        acc = torch.bmm(Ab3[0], Rb3[0], out_dtype=torch.float32)
        for (i, j) in ((0, 1), (1, 0), (0, 2), (1, 1), (2, 0)):
            acc += torch.bmm(Ab3[i], Rb3[j], out_dtype=torch.float32)
        return acc
    variants.append(("bf16 x 6 (three-way split)", 6, bf16x6))

    outs = C * nblk * B
    for name, gemms, fn in variants:
        torch.backends.cuda.matmul.allow_tf32 = not name.startswith("fp32")
        ms, y = timed(fn)
        err = ((y[0].double() - truth).pow(2).mean().sqrt() / t_rms).item()
        flop = gemms * 2.0 * (2 * B) * (4 * B) * nblk * C
        print(json.dumps({"probe": "toeplitz_gemm", "variant": name, "gemms": gemms, "channels": C, "samples_per_channel": n,
                          "ms": ms, "gsamples_per_s": outs / (ms * 1e-3) / 1e9, "tflops_executed": flop / (ms * 1e-3) / 1e12,
                          "rel_rms_error_vs_float64": err, "meets_1e-5": err < 1e-5,
                          "executed_flop_per_output": gemms * 16.0 * K}))
    # the fused overlap-save kernel on the same channels: one bank launch over (channel, block), as b200c_fir_bank_run does
    from pothoscomms_b200 import FirFilterBank
    bank = FirFilterBank(1, "COMPLEX", C)
    for c in range(C):
        bank.set_taps(c, wl.bank_taps(c, 1024, K))
    xin = x[:, B - (K - 1):].contiguous()                                   # [C, K-1 + n, 2]
    out = torch.empty((C, n, 2), dtype=torch.float32, device=dev)
    ms, _ = timed(lambda: bank.run(xin, out))
    y = out[0].view(nblk, B, 2)
    got = torch.cat([y[..., 0], y[..., 1]], dim=1).t().double()        # [2B, nblk], same layout as truth
    err = ((got - truth).pow(2).mean().sqrt() / t_rms).item()
    print(json.dumps({"probe": "toeplitz_gemm", "variant": "fir_os64_kernel (fused overlap-save, FFMA), one bank launch", "channels": C,
                      "samples_per_channel": n, "ms": ms, "gsamples_per_s": outs / (ms * 1e-3) / 1e9,
                      "rel_rms_error_vs_float64": err, "meets_1e-5": err < 1e-5,
                      "note": f"{C} channels only (a partial wave on 148 SMs); the 1024-channel launch is bench.py's c5_bank line"}))


if __name__ == "__main__":
    main()
