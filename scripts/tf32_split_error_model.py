#!/usr/bin/env python
"""Error model of the split-precision GEMM variants (CPU, NumPy): how far each is from the exact
product, next to plain FP32 accumulation.  Operands as the tensor pipe sees them: a TF32 operand keeps
the top 19 bits of the FP32 container (10 explicit mantissa bits); products are exact, accumulation is
modelled in FP64 and rounded once (the pipe's FP32 accumulation adds its own ~1e-7 on top, identical
for every variant).

  fp32        : plain float32 dot (what the oracle / the SIMT kernel do)
  1xTF32      : a_hi * w_hi                                  (RUVGNN_MODE_TF32, not a parity mode)
  3xTF32      : a_hi*w_hi + a_lo*w_hi + a_hi*w_lo            (RUVGNN_MODE_FP32_TC today)
  3xTF32/bf16 : a_hi*w_hi + a_lo*w_hi + bf16(a)*bf16(w_lo)   (candidate: W_lo shipped as bf16, 20% fewer
                bytes through the SM fill port)
"""
import numpy as np


def tf32_round(x):  # hi part as the library forms it: round to nearest at 13 dropped bits
    b = x.astype(np.float32).view(np.uint32)
    return ((b + np.uint32(0x1000)) & np.uint32(0xFFFFE000)).view(np.float32)


def tf32_trunc(x):  # what the MMA reads from a 32-bit container
    return (x.astype(np.float32).view(np.uint32) & np.uint32(0xFFFFE000)).view(np.float32)


def bf16_round(x):
    b = x.astype(np.float32).view(np.uint32)
    r = b + np.uint32(0x7FFF) + ((b >> np.uint32(16)) & np.uint32(1))
    return (r & np.uint32(0xFFFF0000)).view(np.float32)


def main():
    rng = np.random.default_rng(0)
    print(f"{'K':>5} {'variant':>12} {'max rel err':>12} {'rms rel err':>12}   (relative to |a|.|w| row scale)")
    for K in (128, 384, 768):
        a = rng.standard_normal((4096, K)).astype(np.float32)
        w = (rng.standard_normal((K, 64)) * np.sqrt(2.0 / (2 * K))).astype(np.float32)
        exact = a.astype(np.float64) @ w.astype(np.float64)
        scale = np.abs(a).astype(np.float64) @ np.abs(w).astype(np.float64)  # sum |a||w| per output
        a_hi, w_hi = tf32_round(a), tf32_round(w)
        a_lo, w_lo = tf32_trunc(a - a_hi), tf32_trunc(w - w_hi)  # exact differences, read as TF32
        d = lambda x, y: x.astype(np.float64) @ y.astype(np.float64)
        variants = {
            "fp32": (a @ w).astype(np.float64),
            "1xTF32": d(a_hi, w_hi),
            "3xTF32": d(a_hi, w_hi) + d(a_lo, w_hi) + d(a_hi, w_lo),
            "3xTF32/bf16": d(a_hi, w_hi) + d(a_lo, w_hi) + d(bf16_round(a), bf16_round(w - w_hi)),
        }
        for name, got in variants.items():
            rel = np.abs(got - exact) / scale
            print(f"{K:>5} {name:>12} {rel.max():12.2e} {np.sqrt((rel ** 2).mean()):12.2e}")


if __name__ == "__main__":
    main()
