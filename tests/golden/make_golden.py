"""Freeze oracle outputs as golden vectors (run in the build container, commit the .npz files).

    python tests/golden/make_golden.py

The reference is MATLAB and cannot be executed here (no MATLAB/Octave offline), so these are frozen
outputs of oracle/rsp_oracle.py -- regression pins for the oracle and device-independent expected
values for the GPU tests -- NOT outputs of the reference itself (parity unpinned).
Inputs are regenerated from seeds (NumPy PCG64) at test time; only small outputs are stored.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(HERE)), "oracle"))
import rsp_oracle as o  # noqa: E402


def freeze(name: str, seed: int = 0):
    cfg, pre, raw = o.make_cube(name, seed)
    res = o.process_cube(raw.astype(np.complex128), cfg, pre, workers=-1)
    det = res.raw_detections
    # a window of the maps around the strongest detection of the first pair that has one
    i = int(np.argmax(det[:, 3]))
    v0, r0, p0 = int(det[i, 0]) - 1, int(det[i, 1]) - 1, int(det[i, 2]) - 1
    gs = slice(max(r0 - 8, 0), r0 + 8)
    out = dict(
        raw_detections=det, parameterized=res.parameterized, stage1=res.stage1, final_targets=res.final_targets,
        window_origin=np.array([p0, gs.start, 0]),
        rdm_window=res.rdm[p0:p0 + 2, gs, :].astype(np.complex64),
        pc_window=res.pc[:, p0, gs].astype(np.complex64),
        beam_window=res.beam[0, :, 1000:1064].astype(np.complex64),
        rdm_abs_sum=np.array([np.abs(res.rdm).sum()]),
        rdm_sum=np.array([res.rdm.sum()]),
        rdm_peak=np.array([np.abs(res.rdm).max()]),
        raw_checksum=np.array([raw.astype(np.complex128).sum()]),
    )
    path = os.path.join(HERE, f"{name}_seed{seed}.npz")
    np.savez_compressed(path, **out)
    print(name, "detections", len(det), "final", len(res.final_targets), "->", os.path.getsize(path), "bytes")


if __name__ == "__main__":
    for nm in ("cfg1", "cfg2", "native"):
        freeze(nm)
