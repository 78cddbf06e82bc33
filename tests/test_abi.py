"""CPU checks of the drop-in boundary: librsp.so loads without a GPU, exports every symbol
include/rsp.h declares, refuses to run without a device (no CPU fallback), and its host-side
pieces (detection ordering, S10/S11 clustering, struct layouts) agree with the oracle."""
import ctypes as C
import os
import re

import numpy as np
import pytest

import rsp_b200 as rsp
from rsp_b200 import _abi
from conftest import ROOT, has_gpu, oracle as o


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "rsp.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(rsp_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    lib = _abi.load()
    declared = _declared_symbols()
    assert len(declared) >= 20
    bound = {name for name, _, _ in _abi.SYMBOLS}
    assert set(declared) == bound, set(declared) ^ bound
    for name in declared:
        assert hasattr(lib, name), name
    assert lib.rsp_abi_version() == _abi.RSP_ABI_VERSION


def test_struct_layouts_match_header(tmp_path):
    """Compile a C program against include/rsp.h and compare sizeof/offsetof with the ctypes mirror."""
    import subprocess
    src = tmp_path / "sz.c"
    src.write_text('''#include <stdio.h>
#include <stddef.h>
#include "rsp.h"
int main(void) {
  printf("%zu %zu %zu %zu %zu %zu %zu %zu %zu %zu\\n", sizeof(rsp_params), sizeof(rsp_constants), sizeof(rsp_detection),
         sizeof(rsp_target), sizeof(rsp_cluster_params), sizeof(rsp_info), offsetof(rsp_params, t_cfar),
         offsetof(rsp_constants, delta_r), offsetof(rsp_detection, range), offsetof(rsp_info, algorithmic_bytes_per_cpi));
  return 0; }''')
    exe = tmp_path / "sz"
    subprocess.check_call(["gcc", "-std=c99", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)])
    got = [int(x) for x in subprocess.check_output([str(exe)]).split()]
    want = [C.sizeof(_abi.rsp_params), C.sizeof(_abi.rsp_constants), C.sizeof(_abi.rsp_detection),
            C.sizeof(_abi.rsp_target), C.sizeof(_abi.rsp_cluster_params), C.sizeof(_abi.rsp_info),
            _abi.rsp_params.t_cfar.offset, _abi.rsp_constants.delta_r.offset, _abi.rsp_detection.range.offset,
            _abi.rsp_info.algorithmic_bytes_per_cpi.offset]
    assert got == want
    assert C.sizeof(_abi.rsp_detection) == 40 == np.dtype(rsp.DETECTION_DTYPE).itemsize
    assert C.sizeof(_abi.rsp_target) == 32 == np.dtype(rsp.TARGET_DTYPE).itemsize


@pytest.mark.skipif(has_gpu(), reason="checks the no-device behaviour")
def test_no_device_is_a_loud_error_not_a_fallback():
    config, cfar_params, cluster_params = rsp.named_config("cfg1")
    pd = rsp.build_precomputed_data(config)
    with pytest.raises(rsp.RspError) as ei:
        rsp.RadarChain(config, cfar_params, pd)
    assert ei.value.code == _abi.RSP_ERR_NO_DEVICE
    assert "no CPU fallback" in str(ei.value)


def test_create_rejects_bad_arguments():
    lib = _abi.load()
    p = _abi.rsp_params()
    ctx = C.c_void_p()
    p.abi_version = 99
    assert lib.rsp_create(C.byref(p), C.byref(ctx)) == _abi.RSP_ERR_INVALID_ARG
    p.abi_version = _abi.RSP_ABI_VERSION
    p.n_channels, p.n_beams, p.n_pulses, p.n_samples = 64, 13, 32, 4096
    assert lib.rsp_create(C.byref(p), C.byref(ctx)) == _abi.RSP_ERR_UNSUPPORTED
    assert b"channels" in lib.rsp_last_error(None)


def test_entry_points_reject_a_null_context_without_touching_a_device():
    """Every entry point of the frame / stream paths answers a NULL context with RSP_ERR_INVALID_ARG (no crash, no device
    call) -- what a host binding sees when rsp_create failed and its return code was ignored."""
    lib = _abi.load()
    cp = _abi.rsp_cluster_params(30.0, 0.4, 5.0)
    n = C.c_int32(0)
    null = C.c_void_p()
    assert lib.rsp_process_frames(null, null, null, 1, 1.0, null, C.byref(cp), 0, 0, null, 0, null, null, 0, null) == _abi.RSP_ERR_INVALID_ARG
    assert lib.rsp_submit_targets(null, null, 0, 1.0, 0, 0) == _abi.RSP_ERR_INVALID_ARG
    assert lib.rsp_fetch_targets(null, 0, C.byref(cp), null, 0, C.byref(n), null, 0, C.byref(n)) == _abi.RSP_ERR_INVALID_ARG
    assert lib.rsp_stream_fetch(null, 0, null, 0, C.byref(n)) == _abi.RSP_ERR_INVALID_ARG
    assert lib.rsp_submit_cpi(null, null, null, 0) == _abi.RSP_ERR_INVALID_ARG
    assert lib.rsp_stream_enqueue(null, null, 0, null, 0, 0, 0) == _abi.RSP_ERR_INVALID_ARG


def test_sort_is_reference_find_order():
    rng = np.random.default_rng(0)
    d = np.zeros(500, dtype=rsp.DETECTION_DTYPE)
    d["pair_idx"], d["r_idx"], d["v_idx"] = rng.integers(1, 13, 500), rng.integers(16, 3390, 500), rng.integers(16, 317, 500)
    s = rsp.sort_detections(d)
    keys = list(zip(s["pair_idx"].tolist(), s["r_idx"].tolist(), s["v_idx"].tolist()))
    assert keys == sorted(keys)


@pytest.mark.parametrize("name", ["cfg1", "cfg2", "native"])
def test_host_clustering_matches_oracle_on_golden_detections(name):
    g = np.load(os.path.join(ROOT, "tests", "golden", f"{name}_seed0.npz"))
    par, raw = g["parameterized"], g["raw_detections"]
    d = np.zeros(len(par), dtype=rsp.DETECTION_DTYPE)
    d["v_idx"], d["r_idx"], d["pair_idx"] = raw[:, 0], raw[:, 1], raw[:, 2]
    d["range"], d["velocity"], d["angle"], d["power"] = par[:, 0], par[:, 1], par[:, 2], par[:, 3]
    _, _, cluster_params = rsp.named_config(name)
    s1, fin = rsp.cluster(d, cluster_params)
    assert len(s1) == len(g["stage1"]) and len(fin) == len(g["final_targets"])
    for got, want in ((s1, g["stage1"]), (fin, g["final_targets"])):
        m = np.stack([got["range"], got["velocity"], got["angle"], got["power"]], 1)
        assert np.allclose(m, want, rtol=2e-7, atol=1e-6)       # power travels as fp32
    s1, fin = rsp.cluster(d[:0], cluster_params)
    assert len(s1) == 0 and len(fin) == 0


def test_clustering_is_order_dependent_like_the_reference():
    """Chain A-B-C where A~B and B~C but A!~C: BFS from A collects all three (fsf:313-336)."""
    d = np.zeros(4, dtype=rsp.DETECTION_DTYPE)
    d["range"] = [1000, 1025, 1050, 3000]
    d["velocity"] = [5.0, 5.2, 5.4, 5.0]
    d["angle"] = [10, 10, 10, 10]
    d["power"] = [1, 2, 3, 9]
    cp = rsp.Struct(max_range_sep=30.0, max_vel_sep=0.4, max_angle_sep=5.0)
    s1, fin = rsp.cluster(d, cp)
    ref1 = o.cluster_stage1(np.stack([d["range"], d["velocity"], d["angle"], d["power"], np.ones(4)], 1),
                            o.Config())
    assert len(s1) == len(ref1) == 2 and s1["power"][0] == 6.0
    assert np.allclose(np.stack([s1["range"], s1["velocity"], s1["angle"], s1["power"]], 1), ref1)


@pytest.mark.parametrize("seed", range(6))
def test_clustering_equals_the_literal_bfs_on_random_scenes(seed):
    """rsp_cluster labels connected components with a sorted sweep; the oracle runs the reference's FIFO search
    literally (fsf:313-336).  Random scenes with chains, ties on the gate, duplicates and unsorted input."""
    rng = np.random.default_rng(seed)
    n = int(rng.integers(1, 400))
    d = np.zeros(n, dtype=rsp.DETECTION_DTYPE)
    centres = rng.uniform(500, 9000, size=max(1, n // 25))
    d["range"] = rng.choice(centres, n) + np.round(rng.uniform(-60, 60, n) / 7.5) * 7.5     # gate ties are common
    d["velocity"] = rng.choice([-8.0, 0.0, 5.0, 5.5], n) + rng.normal(0, 0.3, n)
    d["angle"] = rng.choice([-5.0, 8.0, 10.0], n) + rng.normal(0, 0.8, n)
    d["power"] = rng.uniform(1, 1e4, n).astype(np.float32)
    if seed % 2:
        d[: n // 3] = d[: n // 3][::-1]
        d[n // 2] = d[0]                                                                      # exact duplicate
    cfg = o.Config()
    cp = rsp.Struct(max_range_sep=cfg.max_range_sep, max_vel_sep=cfg.max_vel_sep, max_angle_sep=cfg.max_angle_sep)
    s1, fin = rsp.cluster(d, cp)
    par = np.stack([d["range"], d["velocity"], d["angle"], d["power"].astype(np.float64), np.ones(n)], 1)
    ref1 = o.cluster_stage1(par, cfg)
    ref2 = o.cluster_stage2(ref1, cfg)
    assert len(s1) == len(ref1) and len(fin) == len(ref2)
    for got, want in ((s1, ref1), (fin, ref2)):
        m = np.stack([got["range"], got["velocity"], got["angle"], got["power"]], 1)
        assert np.allclose(m, want, rtol=1e-12, atol=0)


@pytest.mark.parametrize("case", ["equal_keys", "zero_sep", "nan_range", "dense_1600", "wide_sep"])
def test_clustering_cell_sweep_degenerate_scenes(case):
    """The cell sweep of rsp_cluster (cells a quarter of max_range_sep wide) against the literal search where the cell
    arithmetic degenerates: one cell, zero-width gate, a non-finite key, a dense 64-target frame, a gate wider than the scene."""
    rng = np.random.default_rng(11)
    cfg = o.Config()
    n = 1600 if case == "dense_1600" else 120
    d = np.zeros(n, dtype=rsp.DETECTION_DTYPE)
    centres = rng.uniform(4000, 34000, size=64 if case == "dense_1600" else 6)
    which = rng.integers(0, len(centres), n)
    d["range"] = centres[which] + np.round(rng.uniform(-14, 14, n) / 6.0) * 6.0
    d["velocity"] = (which % 7) * 3.0 + rng.choice([0.0, 2.4], n) + rng.normal(0, 0.05, n)
    d["angle"] = (which % 5) * 4.0 + rng.normal(0, 0.5, n)
    d["power"] = rng.uniform(1, 1e4, n).astype(np.float32)
    sep = cfg.max_range_sep
    if case == "equal_keys":
        d["range"] = 5000.0
    elif case == "zero_sep":
        sep = 0.0
    elif case == "nan_range":
        d["range"][7] = np.nan
        d["range"][31] = np.inf
    elif case == "wide_sep":
        sep = 1e6
    cfg.max_range_sep = sep
    cp = rsp.Struct(max_range_sep=sep, max_vel_sep=cfg.max_vel_sep, max_angle_sep=cfg.max_angle_sep)
    s1, fin = rsp.cluster(d, cp)
    par = np.stack([d["range"], d["velocity"], d["angle"], d["power"].astype(np.float64), np.ones(n)], 1)
    with np.errstate(invalid="ignore"):
        ref1 = o.cluster_stage1(par, cfg)
        ref2 = o.cluster_stage2(ref1, cfg)
    assert len(s1) == len(ref1) and len(fin) == len(ref2)
    for got, want in ((s1, ref1), (fin, ref2)):
        m = np.stack([got["range"], got["velocity"], got["angle"], got["power"]], 1)
        assert np.allclose(m, want, rtol=1e-12, atol=0, equal_nan=True)


def test_product_precompute_equals_oracle_precompute():
    for name in ("native", "cfg1", "cfg2", "cfg3"):
        config, cfar_params, _ = rsp.named_config(name)
        pd = rsp.build_precomputed_data(config)
        ocfg = o.make_config(name)
        pre = o.build_precomputed(ocfg)
        for k in ("tx_pulse", "MF_narrow", "MF_medium_win", "MF_long_win", "MF_medium_fft", "MF_long_fft", "MTD_win",
                  "velocity_axis", "range_axis", "beam_angles_deg", "k_slopes_LUT", "DBF_coeffs_data_C"):
            assert np.allclose(np.asarray(pd[k]), np.asarray(pre[k]), rtol=1e-13, atol=1e-13), (name, k)
        for k in ("fir_delay", "N_fft_med", "N_fft_long", "N_gate_narrow", "N_gate_medium", "N_gate_long",
                  "N_total_gate", "seg_start_narrow", "seg_start_medium", "seg_start_long"):
            assert pd[k] == pre[k], (name, k)
        assert pd.deltaR == pre["deltaR"] and pd.deltaV == pre["deltaV"]
        assert (cfar_params.guardCells_V, cfar_params.refCells_V) == (ocfg.guardCells_V, ocfg.refCells_V)


def test_host_echo_synthesis_equals_oracle():
    config, _, _ = rsp.named_config("cfg1")
    pd = rsp.build_precomputed_data(config)
    ocfg = o.make_config("cfg1")
    pre = o.build_precomputed(ocfg)
    tg = o.targets_t3(ocfg, pre)
    mine = rsp.synthesize_echo([dict(Range=t.Range, Velocity=t.Velocity, ElevationAngle=t.ElevationAngle,
                                     SNR_dB=t.SNR_dB, const_H=1.0) for t in tg], config, pd)
    assert np.allclose(mine, o.synthesize_echo(tg, ocfg, pre), rtol=1e-12, atol=1e-12)


def test_mex_gateways_compile_against_the_prototype_shim():
    """MATLAB/Octave/mex.h are absent here; the gateway sources must at least be valid C++ against the
    declared subset of the MEX API and against include/rsp.h."""
    import subprocess
    for f in ("fun_process_single_frame_mex.cpp", "process_stage2_mtd_mex.cpp"):
        subprocess.check_call(["g++", "-std=c++17", "-fsyntax-only", "-Wall", "-Werror", "-I", os.path.join(ROOT, "mex", "mex_shim"),
                               "-I", os.path.join(ROOT, "include"), "-I", os.path.join(ROOT, "mex"), os.path.join(ROOT, "mex", f)])
