#!/usr/bin/env python
"""Cycle model of one 128-row tile of the fused K-upd kernel (gemm_upd_fused.cu), fitted to the clock64 timelines in
profiles/r02o_upd_fused_timeline.txt / r02t_upd_fused_timeline.txt, and what it predicts for the candidates of
DESIGN.md section 9.  All constants are measured on this pool's B200s (1.77 GHz under the bench's power cap).

  MMA cost per 128 x 128 x 8 TF32 instruction as the ONE issuing thread sees it, by tcgen05.commit density:
     90 cycles   no commit between k-blocks' MMAs beyond the weight stage's   (C1, C2: 4.3k cycles / 48 MMAs)
    103 cycles   1.5 commits per k-block on average                           (Z, R:   9.9k cycles / 96 MMAs)
    138 cycles   2 commits per k-block (weight stage + A slot)                (S1:    33.1k cycles / 240 MMAs)
     65.5 cycles the tensor peak for this shape (2000 TF32 MAC / clk / SM)
  exposed epilogue between dependent GEMMs: E1 1.9k, Ez 1.2k, Er 1.3k (+0.1k drain before C1)
"""
import argparse

CLOCK_GHZ = 1.77
SMS = 148


def tile_cycles(k1, mma_s1=138.0, mma_gate=103.0, mma_cand=90.0, e1=1900.0, ez=1200.0, er=1300.0, drain=120.0, passes=3):
    per_kb = 4 * passes  # MMAs per 32-float k-block
    s1 = (k1 // 32) * per_kb * mma_s1
    gates = 2 * 8 * per_kb * mma_gate
    cand = 8 * per_kb * mma_cand
    return s1 + gates + cand + e1 + ez + er + drain


def ms_per_million_rows(cycles):
    tiles_per_sm = 1e6 / 128 / SMS
    return tiles_per_sm * cycles / (CLOCK_GHZ * 1e6)


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--heads", type=int, default=4)
    a = ap.parse_args()
    k1 = (a.heads + 1) * 128
    now = tile_cycles(k1)
    print(f"K1 = {k1}: modelled tile {now:.0f} cycles = {ms_per_million_rows(now):.2f} ms per 1M rows"
          + (" (measured 67.3k cycles, 2.02 ms)" if k1 == 640 else ""))
    for name, kw in (("stage 1 with one commit per k-block (shared barrier)", dict(mma_s1=103.0)),
                     ("every GEMM at the commit-free issue rate", dict(mma_s1=90.0, mma_gate=90.0)),
                     ("... and no exposed read-out / split", dict(mma_s1=90.0, mma_gate=90.0, e1=0, ez=0, er=0)),
                     ("tensor peak for 128 x 128 x 8", dict(mma_s1=65.5, mma_gate=65.5, mma_cand=65.5, e1=0, ez=0, er=0))):
        c = tile_cycles(k1, **kw)
        print(f"  {name}: {c:.0f} cycles = {ms_per_million_rows(c):.2f} ms per 1M rows")
