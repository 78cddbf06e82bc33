"""CPU checks of the *kernel bodies* (csrc/rsp_phases.cuh) compiled for the host.

The CUDA kernels are written as sequences of host/device "phases"; csrc/host_emul.cpp runs the same
phases with loops instead of barriers.  Here they are compared with NumPy / the oracle, so the
butterflies, digit reversal, overlap-save bookkeeping, Doppler plan, CFAR window arithmetic and the
spline peak search are verified without a GPU.  (The oracle is only the checker.)
"""
import ctypes

import numpy as np
import pytest

from conftest import emul_lib, oracle as o

fp = ctypes.POINTER(ctypes.c_float)
dp = ctypes.POINTER(ctypes.c_double)
ip = ctypes.POINTER(ctypes.c_int)


@pytest.fixture(scope="module")
def lib():
    return emul_lib()


@pytest.fixture(scope="module")
def cfg1():
    cfg, pre, raw = o.make_cube("cfg1", 0)
    beam = o.dbf(raw.astype(np.complex128), pre["DBF_coeffs_data_C"])
    pc = o.pulse_compress(beam, pre)
    return cfg, pre, beam, pc


@pytest.mark.parametrize("R", [2, 4, 8, 16, 32, 64])
@pytest.mark.parametrize("sign", [-1, 1])
def test_small_dft(lib, R, sign):
    rng = np.random.default_rng(R * 7 + sign)
    x = (rng.standard_normal(R) + 1j * rng.standard_normal(R)).astype(np.complex64)
    v = x.copy()
    assert lib.emul_small_dft(R, sign, v.ctypes.data_as(fp)) == 0
    ref = np.fft.fft(x.astype(np.complex128)) if sign < 0 else np.fft.ifft(x.astype(np.complex128)) * R
    assert np.abs(v - ref).max() < 5e-6 * np.abs(ref).max()


def test_digit_reverse_is_permutation(lib):
    for L, rad in ((4096, [16, 16, 16]), (2048, [8, 16, 16]), (1024, [4, 16, 16]), (64, [8, 8]), (32, [8, 4])):
        r = (ctypes.c_int * len(rad))(*rad)
        pos = [lib.emul_digit_reverse(f, L, r, len(rad)) for f in range(L)]
        assert sorted(pos) == list(range(L))


@pytest.mark.parametrize("seg", ["medium", "long"])
@pytest.mark.parametrize("L", [0, 1024, 2048, 4096])
def test_pc_overlap_save_matches_reference_fft_convolution(lib, cfg1, seg, L):
    cfg, pre, beam, pc = cfg1
    N, G = cfg.point_PRT, cfg.n_gates
    g1, g2 = pre["N_gate_narrow"], pre["N_gate_medium"]
    if seg == "medium":
        ss, gate0, ng, taps = pre["seg_start_medium"] - 1, g1, g2, pre["MF_medium_win"]
    else:
        ss, gate0, ng, taps = pre["seg_start_long"] - 1, g1 + g2, G - g1 - g2, pre["MF_long_win"]
    if L and L < len(taps):
        pytest.skip("block shorter than the filter")
    t = np.ascontiguousarray(np.stack([taps.real, taps.imag], -1))
    for (p, b) in ((3, 2), (0, 12), (31, 0)):
        line = np.ascontiguousarray(beam[p, b].astype(np.complex64))
        out = np.zeros(G, np.complex64)
        Lu, nb = ctypes.c_int(), ctypes.c_int()
        rc = lib.emul_pc_segment(line.ctypes.data_as(fp), N, int(ss), int(gate0), int(ng), t.ctypes.data_as(dp),
                                 len(taps), L, out.ctypes.data_as(fp), ctypes.byref(Lu), ctypes.byref(nb))
        assert rc == 0
        ref = pc[p, b, gate0:gate0 + ng]
        assert np.abs(out[gate0:gate0 + ng] - ref).max() <= 2e-6 * np.abs(ref).max()
        # the block writer must not touch gates owned by the other segments
        assert not out[:gate0].any() and not out[gate0 + ng:].any()


@pytest.mark.parametrize("N,expect", [(8192, (1, 1, 1)), (4096, (0, 1, 0)), (16384, (4, 0, 0)), (5819, None)])
def test_pc_mixed_block_plan_covers_the_segment_with_fewest_points(lib, cfg1, N, expect):
    """choose_pc_mix: the long segment of config 2 (4826 gates, 700 taps) is one 4096 + one 2048 + one 1024 block
    (7168 points instead of four 2048-point blocks); the parts laid end to end equal a plain linear convolution."""
    cfg, pre, beam, pc = cfg1
    taps = pre["MF_long_win"]
    ss, gate0 = pre["seg_start_long"] - 1, pre["N_gate_narrow"] + pre["N_gate_medium"]
    ng = (N - 2415) - gate0                                  # G = N - 2415
    rng = np.random.default_rng(N)
    line = (rng.standard_normal(N) + 1j * rng.standard_normal(N)).astype(np.complex64)
    out = np.zeros(N, np.complex64)
    counts = (ctypes.c_int * 3)()
    t = np.ascontiguousarray(np.stack([taps.real, taps.imag], -1))
    pts = lib.emul_pc_segment_mixed(line.ctypes.data_as(fp), N, int(ss), int(gate0), int(ng), t.ctypes.data_as(dp),
                                    len(taps), out.ctypes.data_as(fp), counts)
    assert pts > 0
    c = tuple(counts)
    if expect is not None:
        assert c == expect
    valid = [L - (len(taps) - 1) for L in (4096, 2048, 1024)]
    assert sum(n * v for n, v in zip(c, valid)) >= ng
    assert pts == sum(n * L for n, L in zip(c, (4096, 2048, 1024)))
    single = min(-(-ng // v) * L for v, L in zip(valid, (4096, 2048, 1024)) if v > 0)
    assert pts <= single
    full = np.convolve(line[ss:].astype(np.complex128), taps)            # full[j] = sum_k taps[k] y[ss + j - k]
    ref = full[gate0:gate0 + ng]
    assert np.abs(out[gate0:gate0 + ng] - ref).max() <= 3e-6 * np.abs(ref).max()
    assert not out[:gate0].any() and not out[gate0 + ng:].any()


@pytest.mark.parametrize("C,B,N", [(16, 8, 96), (16, 13, 70), (32, 16, 64), (10, 5, 40)])
def test_dbf_weight_fragments_and_lane_mapping(lib, C, B, N):
    """dbf_mma2_kernel's index algebra on the host: weight fragments as the MMA A operand, the B-fragment sample order,
    the D-fragment store addresses.  Result must be x . W' (fun_process_single_frame.m:95)."""
    rng = np.random.default_rng(C * 100 + B)
    x = (rng.standard_normal((C, N)) + 1j * rng.standard_normal((C, N))).astype(np.complex64)
    W = rng.standard_normal((B, C)) + 1j * rng.standard_normal((B, C))
    Wri = np.ascontiguousarray(np.stack([W.real, W.imag], -1))
    out = np.zeros((B, N), np.complex64)
    rc = lib.emul_dbf_wa(np.ascontiguousarray(x).ctypes.data_as(fp), C, N, Wri.ctypes.data_as(dp), B, out.ctypes.data_as(fp))
    assert rc == 0
    ref = (x.astype(np.complex128).T @ W.conj().T).T                     # [B][N]: sum_c x[c][n] conj(W[b][c])
    assert np.abs(out - ref).max() <= 5e-6 * np.abs(ref).max()


def test_cfar_pitch_is_conflict_free_across_rows(lib):
    for need in (32, 48, 64, 80, 128, 144, 272):
        p = lib.emul_cfar4_pitch(need)
        assert p >= need and p % 4 == 0 and p % 32 == 8 and p - need < 32


def test_pc_narrow_fir(lib, cfg1):
    cfg, pre, beam, pc = cfg1
    g1 = pre["N_gate_narrow"]
    fir = pre["MF_narrow"].astype(np.float32)
    line = np.ascontiguousarray(beam[5, 7].astype(np.complex64))
    out = np.zeros(g1, np.complex64)
    assert 0 == lib.emul_pc_narrow(line.ctypes.data_as(fp), cfg.point_PRT, pre["seg_start_narrow"] - 1,
                                   fir.ctypes.data_as(fp), len(fir), pre["fir_delay"], g1, out.ctypes.data_as(fp))
    ref = pc[5, 7, :g1]
    assert np.abs(out - ref).max() <= 2e-6 * np.abs(ref).max()


@pytest.mark.parametrize("P", [8, 16, 32, 64, 128, 256, 512])
def test_mtd_tile(lib, P):
    TG = 32
    rng = np.random.default_rng(P)
    x = (rng.standard_normal((P, TG)) + 1j * rng.standard_normal((P, TG))).astype(np.complex64)
    win = np.kaiser(P, 4.5).astype(np.float32)
    out = np.zeros((TG, P), np.complex64)
    rad = (ctypes.c_int * 4)()
    assert lib.emul_mtd_tile(x.ctypes.data_as(fp), P, TG, win.ctypes.data_as(fp), out.ctypes.data_as(fp), rad) == 0
    ref = np.fft.fftshift(np.fft.fft(x.astype(np.complex128) * win.astype(np.float64)[:, None], axis=0), axes=0).T
    assert np.abs(out - ref).max() <= 3e-6 * np.abs(ref).max(), list(rad)


@pytest.mark.parametrize("P", [332, 83, 20, 166, 24, 7, 96])
def test_mtd_generic_dft_tile(lib, P):
    """mtd_dft_kernel's work items for P = R * Q (the native 332 = 4 * 83, odd P, P with a factor 8)."""
    TG = 16
    rng = np.random.default_rng(P)
    x = (rng.standard_normal((P, TG)) + 1j * rng.standard_normal((P, TG))).astype(np.complex64)
    win = np.kaiser(P, 4.5).astype(np.float32)
    out = np.zeros((TG, P), np.complex64)
    assert lib.emul_mtd_dft_tile(x.ctypes.data_as(fp), P, TG, win.ctypes.data_as(fp), out.ctypes.data_as(fp)) == 0
    ref = np.fft.fftshift(np.fft.fft(x.astype(np.complex128) * win.astype(np.float64)[:, None], axis=0), axes=0).T
    assert np.abs(out - ref).max() <= 5e-6 * np.abs(ref).max()


@pytest.mark.parametrize("KT", [4, 6, 11])
@pytest.mark.parametrize("P", [332, 83, 20, 166, 24, 7, 96, 45])
def test_mtd_generic_dft_tile_k_tiled_is_bit_identical(lib, P, KT):
    """mtd_dft_kernel<TG, R, KT>: KT bins per work item share every input load; same sums in the same order as
    the one-bin item, so the two must agree bit for bit (including a last, partly filled group of bins)."""
    TG = 8
    rng = np.random.default_rng(P + KT)
    x = (rng.standard_normal((P, TG)) + 1j * rng.standard_normal((P, TG))).astype(np.complex64)
    win = np.kaiser(P, 4.5).astype(np.float32)
    one, kt = np.zeros((TG, P), np.complex64), np.zeros((TG, P), np.complex64)
    assert lib.emul_mtd_dft_tile(x.ctypes.data_as(fp), P, TG, win.ctypes.data_as(fp), one.ctypes.data_as(fp)) == 0
    rc = lib.emul_mtd_dft_tile_kt(x.ctypes.data_as(fp), P, TG, KT, win.ctypes.data_as(fp), kt.ctypes.data_as(fp))
    if KT > 4 and P % 8 == 0:
        assert rc == -1                      # 8 x 11 accumulators are not instantiated
        return
    assert rc == 0
    assert np.array_equal(one.view(np.uint32), kt.view(np.uint32))


@pytest.mark.parametrize("KP", [3, 4, 6])
@pytest.mark.parametrize("P", [332, 83, 20, 166, 7, 45, 28])
def test_mtd_generic_dft_tile_folded_form(lib, P, KP):
    """mtd_dft_kernel<TG, R, 100 + KP> for odd Q = P / R (the native 332 = 4 x 83): even / odd folding (mtd_dft_fold_phase) and
    bin pairs (mtd_dft_item_sym) against NumPy, every output row written exactly once (a last, partly filled group included)."""
    TG = 8
    rng = np.random.default_rng(P + KP)
    x = (rng.standard_normal((P, TG)) + 1j * rng.standard_normal((P, TG))).astype(np.complex64)
    win = np.kaiser(P, 4.5).astype(np.float32)
    out = np.zeros((TG, P), np.complex64)
    rc = lib.emul_mtd_dft_tile_sym(x.ctypes.data_as(fp), P, TG, KP, win.ctypes.data_as(fp), out.ctypes.data_as(fp))
    assert rc == 0
    ref = np.fft.fftshift(np.fft.fft(x.astype(np.complex128) * win.astype(np.float64)[:, None], axis=0), axes=0).T
    assert np.abs(out - ref).max() <= 5e-6 * np.abs(ref).max()


@pytest.mark.parametrize("tg", [16, 32, 64])
@pytest.mark.parametrize("shape", [(300, 64, 10, 10, 5, 5), (200, 32, 10, 2, 5, 4), (150, 48, 3, 4, 2, 3)])
def test_cfar_tiles_match_oracle(lib, shape, tg):
    G, P, gR, gV, rR, rV = shape
    rng = np.random.default_rng(G)
    S = rng.rayleigh(1.0, (1, G, P))
    for (g, v) in ((100, P // 2), (40, P // 2 + 3), (G - 20 - 5, P // 2 - 2)):
        S[0, g, v] += 60.0
    S = S.astype(np.float32)
    cfg = o.Config(guardCells_R=gR, guardCells_V=gV, refCells_R=rR, refCells_V=rV, T_CFAR=4.0)
    det = np.zeros((G, P), np.uint8)
    lib.emul_cfar_map(S[0].ctypes.data_as(fp), G, P, gR, gV, rR, rV, ctypes.c_float(4.0), tg,
                      det.ctypes.data_as(ctypes.POINTER(ctypes.c_ubyte)))
    ref = o.cfar_detect(S.astype(np.float64), cfg)
    margin = o.cfar_margin(S.astype(np.float64), cfg)[0]
    got = {(int(v) + 1, int(g) + 1) for g, v in zip(*np.nonzero(det))}
    want = {(int(r[0]), int(r[1])) for r in ref}
    for (v1, g1) in got ^ want:                      # only near-threshold cells may differ
        assert margin[g1 - 1, v1 - 1] < 1e-5
    assert len(want) >= 2
    # and the vectorised oracle equals the literal scalar loops
    assert np.array_equal(ref, o.cfar_detect_scalar(S.astype(np.float64), cfg))


@pytest.mark.parametrize("use_template", [0, 1])
@pytest.mark.parametrize("tg", [16, 32, 64])
@pytest.mark.parametrize("shape", [(300, 64, 10, 10, 5, 5), (200, 32, 10, 2, 5, 4), (150, 48, 3, 4, 2, 3), (120, 332, 10, 10, 5, 5)])
def test_cfar_vectorised_quads_match_oracle(lib, shape, tg, use_template):
    G, P, gR, gV, rR, rV = shape
    rng = np.random.default_rng(G + P)
    S = rng.rayleigh(1.0, (1, G, P))
    for (g, v) in ((100, P // 2), (40, P // 2 + 3), (G - gR - rR - 1, P - gV - rV - 1), (gR + rR, gV + rV)):
        S[0, g, v] += 60.0
    S = S.astype(np.float32)
    cfg = o.Config(guardCells_R=gR, guardCells_V=gV, refCells_R=rR, refCells_V=rV, T_CFAR=4.0)
    det = np.zeros((G, P), np.uint8)
    rc = lib.emul_cfar4_map(S[0].ctypes.data_as(fp), G, P, gR, gV, rR, rV, ctypes.c_float(4.0), tg, use_template,
                            det.ctypes.data_as(ctypes.POINTER(ctypes.c_ubyte)))
    assert rc == 0
    ref = o.cfar_detect(S.astype(np.float64), cfg)
    margin = o.cfar_margin(S.astype(np.float64), cfg)[0]
    got = {(int(v) + 1, int(g) + 1) for g, v in zip(*np.nonzero(det))}
    want = {(int(r[0]), int(r[1])) for r in ref}
    for (v1, g1) in got ^ want:
        assert margin[g1 - 1, v1 - 1] < 1e-5
    assert len(want) >= 4          # including the two corner CUTs


@pytest.mark.parametrize("tg", [20, 40, 80, 120])
@pytest.mark.parametrize("shape", [(300, 64, 10, 10, 5, 5), (200, 32, 10, 2, 5, 4), (260, 128, 10, 10, 5, 5), (173, 32, 10, 10, 5, 5), (150, 16, 4, 2, 5, 4), (120, 332, 10, 10, 5, 5), (130, 48, 10, 2, 5, 4)])
def test_cfar_marching_kernel_phases_match_oracle(lib, shape, tg):
    """cfar5_march / cfar5_doppler (the phases of cfar5_kernel) with NaN in the never-initialised pad columns."""
    G, P, gR, gV, rR, rV = shape
    rng = np.random.default_rng(G + P + tg)
    S = rng.rayleigh(1.0, (1, G, P))
    for (g, v) in ((100, P // 2), (40, P // 2 + 1), (G - gR - rR - 1, P - gV - rV - 1), (gR + rR, gV + rV), (101, P // 2)):
        S[0, g, v] += 60.0
    # cells that pass the range test but fail the Doppler test (a ridge along Doppler), and the reverse
    S[0, 60, :] += 25.0
    S[0, 70:96, P // 2] += 25.0
    S = S.astype(np.float32)
    cfg = o.Config(guardCells_R=gR, guardCells_V=gV, refCells_R=rR, refCells_V=rV, T_CFAR=4.0)
    det = np.zeros((G, P), np.uint8)
    rc = lib.emul_cfar5_map(S[0].ctypes.data_as(fp), G, P, gR, gV, rR, rV, ctypes.c_float(4.0), tg,
                            det.ctypes.data_as(ctypes.POINTER(ctypes.c_ubyte)))
    assert rc == 0
    ref = o.cfar_detect(S.astype(np.float64), cfg)
    margin = o.cfar_margin(S.astype(np.float64), cfg)[0]
    got = {(int(v) + 1, int(g) + 1) for g, v in zip(*np.nonzero(det))}
    want = {(int(r[0]), int(r[1])) for r in ref}
    for (v1, g1) in got ^ want:
        assert margin[g1 - 1, v1 - 1] < 1e-5
    assert len(want) >= 2


def test_spline_peak_matches_scipy(lib):
    lib.emul_spline5_peak.restype = ctypes.c_double
    rng = np.random.default_rng(3)
    for _ in range(200):
        y = rng.rayleigh(1.0, 5)
        y[2] += rng.uniform(0, 5)
        for os_ in (8, 4):
            got = lib.emul_spline5_peak(y.ctypes.data_as(dp), os_)
            assert got == pytest.approx(o._spline_peak(y, os_), abs=1e-12)
