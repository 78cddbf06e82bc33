"""Pin the oracle (CPU): known-answer facts of the reference's configuration, built-in
equivalences, and the frozen golden vectors.  The reference ships no tests or vectors and cannot be
run offline (MATLAB), so these are the pins that exist (SURVEY.md sections 4 and 8(c))."""
import os

import numpy as np
import pytest

from conftest import oracle as o

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def test_native_derived_constants():
    """constants implied by v8_3:67-84,121-183 at the literal configuration."""
    cfg = o.native_config()
    pre = o.build_precomputed(cfg)
    assert cfg.point_PRT == 5819 and pre["N_total_gate"] == 3404
    assert pre["fir_delay"] == 17
    assert pre["N_fft_med"] == 8192 and pre["N_fft_long"] == 8192
    assert (pre["seg_start_narrow"], pre["seg_start_medium"], pre["seg_start_long"]) == (5, 490, 1985)
    assert 11.4e-6 * 25e6 == 285.0 and 31.8e-6 * 25e6 == 795.0          # used un-rounded as indices
    tx = pre["tx_pulse"]
    nz = np.nonzero(tx)[0]
    assert nz[0] == 0 and tx[289] != 0 and tx[288] == 0 and tx[1284] != 0 and tx[1283] == 0 and nz[-1] == 1983
    assert pre["P_signal_unscaled"] == pytest.approx(1.0, abs=1e-12)
    assert pre["deltaR"] == pytest.approx(5.99584916, abs=1e-8)
    assert pre["deltaV"] == pytest.approx(0.2052638, abs=1e-6)
    assert pre["v_max"] == pytest.approx(68.1476, abs=1e-4)
    assert len(pre["MF_medium_win"]) == 200 and len(pre["MF_long_win"]) == 700
    assert pre["MF_narrow"].max() == 6.0 and np.allclose(pre["MF_narrow"], pre["MF_narrow"][::-1])


def test_dbf_csv_steers_to_documented_beam_angles():
    """W' must steer to the beam_angles_deg table (v8_3:178) the author read off plot_beam_patterns.m."""
    cfg = o.native_config()
    W = o.load_dbf_csv()
    assert W.shape == (13, 16)
    th = np.linspace(-30, 80, 2201)
    steer = np.stack([o.channel_phasors(t, cfg) for t in th], axis=1)      # [C, n_theta]
    pattern = np.abs(np.conj(W) @ steer)                                    # x * W' response
    peaks = th[np.argmax(pattern, axis=1)]
    assert np.all(np.abs(peaks - o.BEAM_ANGLES_DEG) <= 0.9), peaks


def test_generalised_shapes():
    for name, G in (("cfg1", 1681), ("cfg2", 5777), ("cfg3", 13969)):
        cfg = o.make_config(name)
        assert cfg.n_gates == G == cfg.point_PRT - 2415
        pre = o.build_precomputed(cfg)
        assert pre["N_fft_long"] == cfg.point_PRT
        assert pre["DBF_coeffs_data_C"].shape == (cfg.beam_num, cfg.channel_num)
    assert o.make_config("cfg1").guardCells_V == 2 and o.make_config("cfg1").refCells_V == 4


def test_fft_pulse_compression_equals_linear_convolution():
    """fsf:115-120: N_fft >= linear length, so the FFT method is a plain linear convolution and a
    target with delay d peaks at gate d (1-based) in the medium and long segments."""
    cfg = o.make_config("cfg1")
    pre = o.build_precomputed(cfg)
    rng = np.random.default_rng(0)
    y = rng.standard_normal((1, 1, cfg.point_PRT)) + 1j * rng.standard_normal((1, 1, cfg.point_PRT))
    pc = o.pulse_compress(y, pre)[0, 0]
    g1, g2 = pre["N_gate_narrow"], pre["N_gate_medium"]
    med = np.convolve(y[0, 0, pre["seg_start_medium"] - 1:], pre["MF_medium_win"])
    lng = np.convolve(y[0, 0, pre["seg_start_long"] - 1:], pre["MF_long_win"])
    assert np.allclose(pc[g1:g1 + g2], med[g1:g1 + g2], rtol=0, atol=1e-9)
    assert np.allclose(pc[g1 + g2:], lng[g1 + g2:cfg.n_gates], rtol=0, atol=1e-9)
    for d in (500, 1500):
        raw = o.synthesize_echo([o.Target(d * pre["deltaR"], 0.0, 0.0, 30.0)], cfg, pre)
        line = o.pulse_compress(o.dbf(raw[:1], pre["DBF_coeffs_data_C"]), pre)[0, 2]
        assert int(np.argmax(np.abs(line))) + 1 == d


def test_spline_is_matlab_not_a_knot():
    """5 points of a cubic are reproduced exactly by a not-a-knot spline (interp1 'spline')."""
    x = np.arange(5.0)
    y = 0.3 * x ** 3 - 2.0 * x ** 2 + 3.0 * x + 1.0
    q = np.arange(0, 33) / 8
    want = q[int(np.argmax(0.3 * q ** 3 - 2.0 * q ** 2 + 3.0 * q + 1.0))]
    assert o._spline_peak(y, 8) == want


def test_clustering_order_and_merge_rules():
    cfg = o.Config()
    par = np.array([[1000.0, 10.0, 5.0, 2.0, 1], [1010.0, 10.1, 6.0, 6.0, 1], [1005.0, 10.2, 12.0, 1.0, 2],
                    [5000.0, -3.0, 0.0, 4.0, 3]])
    s1 = o.cluster_stage1(par, cfg)
    assert len(s1) == 3                                     # third point is > 5 deg away in angle
    assert s1[0, 3] == 8.0 and s1[0, 0] == pytest.approx((1000 * 2 + 1010 * 6) / 8)
    fin = o.cluster_stage2(s1, cfg)
    assert len(fin) == 2 and fin[0, 3] == 8.0               # winner takes all on (R, V)
    assert len(o.cluster_stage1(np.zeros((0, 5)), cfg)) == 0 and len(o.cluster_stage2(np.zeros((0, 4)), cfg)) == 0


@pytest.mark.parametrize("name", ["cfg1", "cfg2"])
def test_oracle_reproduces_golden(name):
    g = np.load(os.path.join(GOLDEN, f"{name}_seed0.npz"))
    cfg, pre, raw = o.make_cube(name, 0)
    assert raw.astype(np.complex128).sum() == pytest.approx(g["raw_checksum"][0], rel=1e-12)
    res = o.process_cube(raw.astype(np.complex128), cfg, pre, workers=-1)
    assert np.array_equal(res.raw_detections[:, :3], g["raw_detections"][:, :3])
    assert np.allclose(res.raw_detections[:, 3], g["raw_detections"][:, 3], rtol=1e-9)
    assert np.allclose(res.parameterized, g["parameterized"], rtol=1e-9, atol=1e-9)
    assert np.allclose(res.final_targets, g["final_targets"], rtol=1e-9, atol=1e-9)
    p0, g0, _ = g["window_origin"]
    assert np.allclose(res.rdm[p0:p0 + 2, g0:g0 + g["rdm_window"].shape[1], :], g["rdm_window"], rtol=1e-5, atol=1e-3)
    assert np.abs(res.rdm).sum() == pytest.approx(g["rdm_abs_sum"][0], rel=1e-9)


def test_known_answer_native_targets():
    """v8_3:30-37 targets through the frozen native run: two final targets at the configured
    range / radial velocity / 10 deg elevation."""
    g = np.load(os.path.join(GOLDEN, "native_seed0.npz"))
    fin = g["final_targets"]
    assert len(fin) == 2
    assert fin[0, 0] == pytest.approx(2991.9, abs=1.0) and fin[0, 1] == pytest.approx(20.18, abs=0.05)
    assert fin[1, 0] == pytest.approx(9995.1, abs=1.0) and fin[1, 1] == pytest.approx(25.19, abs=0.05)
    assert np.all(np.abs(fin[:, 2] - 10.0) < 0.05)
    assert 150 <= len(g["raw_detections"]) <= 250


def test_stage2_specification_is_a_per_segment_correlation():
    """The semantics this repo specifies for process_stage2_mtd (the reference's callee is not shipped):
    pc = per-segment cross-correlation with the reference pulse, mtd = fftshift(fft) with a zero-Doppler notch."""
    rng = np.random.default_rng(0)
    gates, P, B = [40, 90, 150], 16, 2
    pulses = [np.ones(4, complex), np.exp(1j * rng.uniform(0, 6, 20)), np.exp(1j * rng.uniform(0, 6, 50))]
    iq = rng.standard_normal((P, sum(gates), B)) + 1j * rng.standard_normal((P, sum(gates), B))
    mtd, pc = o.stage2_mtd(iq, gates, pulses, None, 2)
    g0 = 0
    for ng, pulse in zip(gates, pulses):
        seg = iq[3, g0:g0 + ng, 1]
        want = np.correlate(np.concatenate([seg, np.zeros(len(pulse) - 1)]), pulse, mode="valid")
        assert np.allclose(pc[3, g0:g0 + ng, 1], want)
        g0 += ng
    full = np.fft.fftshift(np.fft.fft(pc, axis=0), axes=0)
    assert not mtd[P // 2 - 2: P // 2 + 3].any()
    keep = np.r_[0:P // 2 - 2, P // 2 + 3:P]
    assert np.allclose(mtd[keep], full[keep])


def test_range_cfar_1d_literal_restatement_properties():
    """f-3 (debug_simulated_data_processing_v2.m:419-511): hand-checkable facts of the literal restatement."""
    V, seg = 12, [80, 44, 60]
    R = sum(seg)
    amp = np.ones((V, R))
    amp[3, 20] = 50.0            # narrow segment
    amp[8, 80 + 2] = 50.0        # third column of the medium segment: the left window leaves the segment
    amp[9, R - 1] = 50.0         # last gate of the long segment: the right window leaves the segment
    amp[6, 150] = 50.0           # on the zero-velocity row round(12/2)+1 = 7 (1-based) -> never tested
    flag, thr = o.local_execute_cfar(amp, seg, 1, 5, 14, 3.0, 0)
    assert flag[3, 20] == 1 and flag[8, 82] == 1 and flag[9, R - 1] == 1
    assert flag[5:8].sum() == 0 and thr[5:8].sum() == 0          # rows 6..8 (1-based) are masked for MTD_0v_num = 1
    assert flag.sum() == 3
    assert np.all(thr[0] == 3.0) and thr[3, 0] == 3.0 and thr[3, 20] == 3.0
    # a strong cell raises the threshold of exactly the cells whose windows contain it: columns 20 +- (15..19) in its segment
    hit = np.flatnonzero(thr[3, :seg[0]] > 3.0)
    assert set(hit) == {1, 2, 3, 4, 5, 35, 36, 37, 38, 39}
    assert np.allclose(thr[3, hit], 3.0 * (4 * 1.0 + 50.0) / 5)
    # smallest-of ignores it where the other window is clean (columns 35..39); at columns 1..5 the left window leaves the
    # segment and is REPLACED by the right one (:481-485), so both hold the strong cell and the level stays raised
    _, thr_so = o.local_execute_cfar(amp, seg, 1, 5, 14, 3.0, 1)
    assert np.all(thr_so[3, 35:40] == 3.0) and np.allclose(thr_so[3, 1:6], 32.4) and np.all(thr_so[3, 6:35] == 3.0)
    with pytest.raises(IndexError):                                  # the reference would index outside a 30-gate segment
        o.local_execute_cfar(np.ones((4, 30)), [30, 0, 0], 0, 5, 14, 3.0, 0)
