#!/usr/bin/env python
"""Inputs for tools/ref_golden/ref_golden.m (the unmodified reference under MATLAB / GNU Octave) and the oracle's answer to
the same inputs, so that the two can be compared number by number (tests/test_reference_golden.py).

    python tools/ref_golden/make_inputs.py --out tests/golden/reference/native_seed7 [--seed 7] [--scene v8_3|v8_2|v7_7]
    octave --eval "addpath('tools/ref_golden'); ref_golden('/root/reference/Simulation', 'tests/golden/reference/native_seed7', 'native', 0)"
    python -m pytest tests/test_reference_golden.py

Writes  noise.bin   32 blocks of P*N float64, column-major [P, N]: I then Q of channel 1, I then Q of channel 2, ... --
                    exactly the order in which fun_process_single_frame.m:81-88 calls randn
        targets.txt one "Range Velocity ElevationAngle SNR_dB" row per target
        meta.txt    case name, shape, probe cells (1-based Doppler row, gate)
        oracle_final_targets.txt / oracle_rdm_checks.txt   the oracle's outputs in the format the .m driver writes.
Only the reference's literal configuration can be pinned this way: ref_golden.m takes every constant from the
reference's own set-up block (main_simulate_echoes_with_array_v8_3.m:21-188).
"""
import argparse
import importlib.util
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def load_oracle():
    spec = importlib.util.spec_from_file_location("rsp_oracle", os.path.join(ROOT, "oracle", "rsp_oracle.py"))
    mod = importlib.util.module_from_spec(spec)
    sys.modules["rsp_oracle"] = mod
    spec.loader.exec_module(mod)
    return mod


PROBES = [(167, 321), (167, 322), (180, 322), (200, 1069), (201, 1069), (40, 100), (300, 3300), (1, 1), (332, 3404)]   # (v, g), 1-based


def noise_cube(P, C, N, seed):
    """The cube that the reference builds from the blocks of noise.bin: noise[p, c, n] = (I_c[p, n] + j Q_c[p, n]) * sqrt(1/2)."""
    rng = np.random.default_rng(seed)
    blocks = rng.standard_normal((2 * C, N, P))            # block b = [N][P] in memory == column-major [P, N]
    I, Q = blocks[0::2], blocks[1::2]                      # [C, N, P]
    noise = (I + 1j * Q) * np.sqrt(0.5)
    return blocks, np.ascontiguousarray(np.transpose(noise, (2, 0, 1)))     # raw layout [p, c, n]


def write_checks(path, rdm, probes):
    """rdm[b, g, v] -> the text ref_golden.m writes."""
    with open(path, "w") as fh:
        for b in range(rdm.shape[0]):
            R = rdm[b]
            fh.write("beam %d %.17g %.17g %.17g\n" % (b + 1, R.real.sum(), R.imag.sum(), (np.abs(R) ** 2).sum()))
            for v, g in probes:
                z = R[g - 1, v - 1]
                fh.write("cell %d %d %d %.17g %.17g\n" % (b + 1, v, g, z.real, z.imag))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", required=True)
    ap.add_argument("--seed", type=int, default=7)
    ap.add_argument("--scene", default="v8_3", choices=["v8_3", "v8_2", "v7_7"])
    a = ap.parse_args()
    o = load_oracle()
    cfg = o.make_config("native")
    pre = o.build_precomputed(cfg)
    targets = getattr(o, "targets_" + a.scene)()
    P, C, N = cfg.prtNum, cfg.channel_num, cfg.point_PRT
    os.makedirs(a.out, exist_ok=True)
    blocks, noise = noise_cube(P, C, N, a.seed)
    blocks.astype("<f8").tofile(os.path.join(a.out, "noise.bin"))
    with open(os.path.join(a.out, "targets.txt"), "w") as fh:
        for t in targets:
            fh.write("%.17g %.17g %.17g %.17g\n" % (t.Range, t.Velocity, t.ElevationAngle, t.SNR_dB))
    with open(os.path.join(a.out, "meta.txt"), "w") as fh:
        fh.write("name native\nP %d\nN %d\nC %d\nseed %d\nscene %s\n" % (P, N, C, a.seed, a.scene))
        for v, g in PROBES:
            fh.write("probe %d %d\n" % (v, g))
    raw = o.synthesize_echo(targets, cfg, pre) + noise
    res = o.process_cube(raw, cfg, pre, workers=-1)
    with open(os.path.join(a.out, "oracle_final_targets.txt"), "w") as fh:
        for t in res.final_targets:
            fh.write("%.17g %.17g %.17g %.17g\n" % tuple(t[:4]))
    write_checks(os.path.join(a.out, "oracle_rdm_checks.txt"), res.rdm, PROBES)
    print("wrote", a.out, "-", len(res.final_targets), "final targets")


if __name__ == "__main__":
    main()
