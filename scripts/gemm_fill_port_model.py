#!/usr/bin/env python
"""Fill-port model of the split-precision GEMM kernels.

A persistent CTA per SM walks its (row tile x column tile) list; per 32-wide k-block it must INGEST the A tile
(128 x 32 FP32 = 16 KB) and the W_hi / W_lo tiles (BN x 32 FP32 each) through the SM's fill port (measured
33 B/cycle/SM over the crossbar) and ISSUE three 128 x BN x 32 TF32 MMAs.  Time per k-block = max of the two.
Prints the prediction next to the ncu launch times of profiles/r01l_launches.md / r01g_kernels_full.csv
(config 2: 1M rows, H = 384) and evaluates the candidates of DESIGN section 8.
"""
SM, CLK_GHZ = 148, 1.72           # SM clock under ncu (sm__cycles_elapsed.avg.per_second)
PORT = 32.0                        # bytes / cycle / SM into L1/SMEM
MMA = 2048.0                       # TF32 MACs / cycle / SM (bf16: 2x)


def gemm_ms(rows, n, k, bn, w_bytes_per_elem=8.0, a_bytes=4.0, rows_per_w_fetch=128, passes=3.0):
    tiles = -(-rows // 128) * (n // bn)
    per_sm = tiles / SM
    kblocks = k // 32
    ingest = 128 * 32 * a_bytes + bn * 32 * w_bytes_per_elem * (128 / rows_per_w_fetch)
    mma = 128 * bn * 32 * passes / MMA
    cyc = max(ingest / PORT, mma)
    return per_sm * kblocks * cyc / (CLK_GHZ * 1e6), min(1.0, mma / (ingest / PORT))


def main():
    rows, H = 1_000_000, 384
    shapes = [  # name, N, K, BN used today, measured ms (ncu, r01g/r01l)
        ("W_msg  (EPI_BIAS)", H, H, 192, 1.62),
        ("Q|K|V  (EPI_BIAS)", 3 * H, H, 192, 4.42),
        ("W_o    (BIAS_ADD)", H, H, 192, 1.75),
        ("W_agg  (EPI_BIAS)", H, H, 192, 1.62),
        ("gates  (EPI_GATES)", 2 * H, 2 * H, 256, 5.92),
        ("GRU    (EPI_GRU)", H, 2 * H, 192, 3.30),
    ]
    print(f"{'GEMM':<20}{'N':>6}{'K':>6}{'BN':>5}{'model ms':>10}{'ncu ms':>8}{'pipe busy (model)':>19}")
    tot_m = tot_x = 0.0
    for name, n, k, bn, meas in shapes:
        ms, util = gemm_ms(rows, n, k, bn)
        tot_m += ms
        tot_x += meas
        print(f"{name:<20}{n:>6}{k:>6}{bn:>5}{ms:>10.2f}{meas:>8.2f}{util:>19.2f}")
    print(f"{'sum':<37}{tot_m:>10.2f}{tot_x:>8.2f}")
    print("\ncandidates (same six GEMMs, model only):")
    for label, kw in [
        ("today (BN as above)", {}),
        ("W_lo as bf16 (6 B per weight element)", {"w_bytes_per_elem": 6.0, "passes": 2.5}),
        ("CTA pair, 256 rows per weight fetch, BN = 256", {"rows_per_w_fetch": 256, "bn": 256}),
        ("pair + bf16 W_lo", {"rows_per_w_fetch": 256, "bn": 256, "w_bytes_per_elem": 6.0, "passes": 2.5}),
    ]:
        total = 0.0
        for name, n, k, bn, _ in shapes:
            b = kw.get("bn", bn)
            while n % b:
                b -= 64
            args = {kk: v for kk, v in kw.items() if kk != "bn"}
            total += gemm_ms(rows, n, k, b, **args)[0]
        print(f"  {label:<48}{total:>7.2f} ms")


if __name__ == "__main__":
    main()
