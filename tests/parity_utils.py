"""Comparison helpers shared by the GPU parity tests, smoke() and the golden-vector tests.

Tolerances (BASELINE.json north_star): complex range-Doppler map <= 1e-4 relative in fp32 --
defined peak-normalised and Frobenius-normalised (a bare per-element ratio is meaningless on the
near-zero cells of a noise map, SURVEY.md section 7 hard part 7); detection cell indices identical
except cells within 1e-4 (relative) of the CFAR threshold; angles <= 0.01 deg.
"""
import numpy as np

RDM_REL_TOL = 1e-4
THRESH_BAND = 1e-4
ANGLE_TOL_DEG = 0.01


def rel_errors(got: np.ndarray, ref: np.ndarray):
    """(max|d|/max|ref|, ||d||_F/||ref||_F)."""
    d = got.astype(np.complex128) - ref
    return float(np.abs(d).max() / np.abs(ref).max()), float(np.linalg.norm(d) / np.linalg.norm(ref))


def compare_detections(dets, ref_rows, margin, ref_par, pre, tol_angle=ANGLE_TOL_DEG, tol_power=1e-5):
    """dets: structured device table; ref_rows: oracle rows [v, r, pair, amp] (1-based);
    margin[pair, g, v]: |S - T*noise|/(T*noise); ref_par: oracle rows [R, V, A, P, pair].
    Returns a dict of statistics; raises AssertionError on a parity violation."""
    got = {(int(d["pair_idx"]), int(d["r_idx"]), int(d["v_idx"])): d for d in dets}
    want = {(int(r[2]), int(r[1]), int(r[0])): i for i, r in enumerate(ref_rows)}
    only_got = set(got) - set(want)
    only_want = set(want) - set(got)
    for (p, r, v) in only_got | only_want:
        m = margin[p - 1, r - 1, v - 1]
        assert m < THRESH_BAND, f"cell (pair {p}, r {r}, v {v}) differs and is {m:.3g} from the threshold"
    # order: the device list must come back in the reference's find order
    keys = [(int(d["pair_idx"]), int(d["r_idx"]), int(d["v_idx"])) for d in dets]
    assert keys == sorted(keys), "detections are not in (pair, range, Doppler) order"
    n_common, max_da, max_dp, n_step = 0, 0.0, 0.0, 0
    dR, dV = pre["deltaR"], pre["deltaV"]
    for key, i in want.items():
        if key not in got:
            continue
        n_common += 1
        d = got[key]
        R, V, A, Pw = ref_par[i][:4]
        max_da = max(max_da, abs(d["angle"] - A))
        max_dp = max(max_dp, abs(d["power"] - Pw) / Pw)
        # spline argmax lives on a 1/8 (range) and 1/4 (Doppler) cell grid: either identical, or --
        # when two grid samples tie to fp32 precision -- one grid step apart
        er, ev = abs(d["range"] - R), abs(d["velocity"] - V)
        if er > 1e-6 * max(1.0, abs(R)) or ev > 1e-6 * max(1.0, abs(V)):
            n_step += 1
            assert er <= dR / 8 * 1.0001 + 1e-9 and ev <= dV / 4 * 1.0001 + 1e-9, (key, er, ev)
    assert max_da <= tol_angle, f"angle differs by {max_da} deg"
    assert max_dp <= tol_power, f"detection power differs by {max_dp} relative"
    assert n_step <= max(1, n_common // 50), f"{n_step} of {n_common} spline peaks moved by a grid step"
    return dict(n_ref=len(want), n_dev=len(got), n_common=n_common, only_dev=len(only_got), only_ref=len(only_want),
                max_angle_err=max_da, max_power_rel=max_dp, spline_step_moves=n_step)


def compare_targets(final_dev, final_ref, pre=None, loose=False, rtol=1e-5, atol_angle=ANGLE_TOL_DEG):
    """Clustered targets.  ``loose`` (a spline peak tied and moved one grid step, or a near-threshold
    cell differed) widens range/velocity to one interpolation step of the weighted mean."""
    assert len(final_dev) == len(final_ref), (len(final_dev), len(final_ref))
    tr = pre["deltaR"] / 8 if (loose and pre is not None) else 0.0
    tv = pre["deltaV"] / 4 if (loose and pre is not None) else 0.0
    for d, r in zip(final_dev, final_ref):
        assert abs(d["range"] - r[0]) <= 1e-3 + rtol * abs(r[0]) + tr, (d, r)
        assert abs(d["velocity"] - r[1]) <= 1e-3 + rtol * abs(r[1]) + tv, (d, r)
        assert abs(d["angle"] - r[2]) <= atol_angle, (d, r)
        assert abs(d["power"] - r[3]) <= 1e-4 * abs(r[3]), (d, r)
