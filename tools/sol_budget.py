#!/usr/bin/env python
"""Speed-of-light budget of one training step (CPU only): what the step would cost if every kernel class ran at its
roofline, under (a) the bytes this implementation moves and (b) the algorithmic minimum of SURVEY.md section 8(d).

    python tools/sol_budget.py [--bars 512] [--measured-ms 41.4] [--gemm-ms 22.9] [--nb-ms 22.5]

Work per bar: contractions 29.494 GFLOP (fwd+bwd); InstanceNorm sites 4.257 M elements
(encoder 0.356 M x2, phrase encoder 1.425 M, decoder 2.120 M), 3.02 M of them under a CBAM; 89.5 M parameters.
Bytes per norm-site element: this implementation 16 forward + 20 backward (16 with the split sweeps on CBAM sites;
DESIGN.md section 4.3); algorithmic minimum 2 read + 2 written forward, 2 + 2 + 2 backward (dout, saved activation, dy).
"""
import argparse
import json
import os

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GFLOP_PER_BAR = 29.494
NORM_ELEMS_PER_BAR = (0.356 * 2 + 1.425 + 2.120) * 1e6
CBAM_ELEMS_PER_BAR = 3.02e6
N_PARAMS = 89.5e6


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--bars", type=int, default=512)
    ap.add_argument("--measured-ms", type=float, default=41.4)
    ap.add_argument("--gemm-ms", type=float, default=22.9, help="contraction launches, serial CUDA-event time")
    ap.add_argument("--nb-ms", type=float, default=22.5, help="norm-block launches, serial CUDA-event time")
    a = ap.parse_args()
    peaks = {"bf16_tflops_sustained": 1383.1, "hbm_gbs": 6547.8}
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        peaks.update(json.load(open(p)))
    tf, bw = peaks["bf16_tflops_sustained"] * 1e12, peaks["hbm_gbs"] * 1e9
    B = a.bars
    t_gemm = GFLOP_PER_BAR * 1e9 * B / tf * 1e3
    plain = NORM_ELEMS_PER_BAR - CBAM_ELEMS_PER_BAR
    impl_bytes = (CBAM_ELEMS_PER_BAR * (16 + 16) + plain * (12 + 14)) * B          # CBAM sites / plain IN sites
    algo_bytes = NORM_ELEMS_PER_BAR * (4 + 6) * B
    t_nb_impl, t_nb_algo = impl_bytes / bw * 1e3, algo_bytes / bw * 1e3
    t_opt = (N_PARAMS * 28 + N_PARAMS * 1.14 * 6) / bw * 1e3                       # Adam 28 B/param + bf16 repack (2 operands)
    rows = [
        ("contractions at the measured sustained tensor peak", t_gemm, a.gemm_ms),
        ("norm blocks at the measured HBM peak, implementation bytes", t_nb_impl, a.nb_ms),
        ("norm blocks at the measured HBM peak, algorithmic bytes", t_nb_algo, a.nb_ms),
        ("Adam + operand repack at the HBM peak", t_opt, 0.81),
    ]
    print("%d bars/step; peaks: %.1f TFLOP/s, %.1f GB/s" % (B, tf / 1e12, bw / 1e9))
    for name, sol, meas in rows:
        print("  %-62s %6.2f ms   measured %5.2f ms  -> %3.0f %% of that bound" % (name, sol, meas, 100 * sol / meas))
    serial = t_gemm + t_nb_impl + t_opt
    print("  serial sum (implementation bytes) %.1f ms; overlapped bound max(tensor, HBM) %.1f ms; measured step %.1f ms "
          "= %.0f %% of the serial sum" % (serial, max(t_gemm, t_nb_impl + t_opt), a.measured_ms,
                                            100 * serial / a.measured_ms))
    print("  with algorithmic norm-block bytes: serial %.1f ms" % (t_gemm + t_nb_algo + t_opt))


if __name__ == "__main__":
    main()
