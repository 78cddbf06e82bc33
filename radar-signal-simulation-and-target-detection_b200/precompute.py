"""Host-side construction of the chain's inputs: the MATLAB structs ``config``, ``cfar_params``,
``cluster_params`` and ``precomputed_data`` with the reference's field names.

Mirrors main_simulate_echoes_with_array_v8_3.m:44-84 (configuration literals) and :121-183 (the
precompute block).  This is one-time host set-up (small vectors), not the hot path; the hot path
consumes these tables through rsp_upload_constants (include/rsp.h).
"""
from __future__ import annotations

import math
import os
from typing import Optional

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
DBF_CSV = os.path.join(HERE, "data", "dbf_coef_x8_250522.csv")      # the reference's X8..._DBFcoef.csv


class Struct(dict):
    """A MATLAB-struct look-alike: fields are attributes and dict keys."""
    __getattr__ = dict.__getitem__
    __setattr__ = dict.__setitem__

    def copy(self):
        return Struct({k: (v.copy() if isinstance(v, (Struct, np.ndarray)) else v) for k, v in self.items()})


# literals of v8_3:141, :178, :179
FIR_COEFFS = (794, 1403, 2143, 2672, 2591, 1711, -58, -2351, -4592, -5855, -5338, -2389, 3005, 10341, 18410,
              25779, 30907, 32768, 30907, 25779, 18410, 10341, 3005, -2389, -5338, -5855, -4592, -2351, -58,
              1711, 2591, 2672, 2143, 1403, 794)
BEAM_ANGLES_DEG = (-16, -9.6, -3.2, 3.2, 9.6, 16, 22.6, 29.2, 36.1, 43.3, 51, 59.6, 70.3)
K_SLOPES_LUT = (-4.6391, -4.6888, -4.7578, -4.7891, -4.7214, -4.7513, -5.2343, -5.4529, -5.7323, -6.1685,
                -7.0256, -8.7612)


def default_config(channel_num: int = 16, beam_num: int = 13, prtNum: int = 332, point_PRT: Optional[int] = None):
    """``(config, cfar_params, cluster_params)``.

    With no arguments: the reference's literal set-up (v8_3:44-84; N = round(prt*fs) = 5819,
    gates [228, 723, 2453]).  With ``point_PRT`` given: the generalised shapes used by the
    benchmark configurations (fs, fc, waveform and segment starts fixed; prt = N/fs; gates
    [228, 723, N-3366]; Doppler CFAR window guard 2 / ref 4 when P == 32)."""
    sc = Struct(c=2.99792458e8, fs=25e6, fc=9450e6, prtNum=int(prtNum), prt=232.76e-6, B=20e6,
                tao=[0.16e-6, 8e-6, 28e-6], gap_duration=[11.4e-6, 31.8e-6, 153.4e-6],
                point_prt_segments=[228, 723, 2453], channel_num=int(channel_num), beam_num=int(beam_num))
    if point_PRT is None:
        sc.point_PRT = int(round(sc.prt * sc.fs))                                  # v8_3:82
    else:
        sc.point_PRT = int(point_PRT)
        sc.prt = sc.point_PRT / sc.fs
        sc.point_prt_segments = [228, 723, sc.point_PRT - 3366]
    sc.wavelength = sc.c / sc.fc                                                   # v8_3:80
    config = Struct(Sig_Config=sc, Array=Struct(element_spacing=0.0138))           # v8_3:79
    cfar_params = Struct(refCells_V=5, guardCells_V=10, refCells_R=5, guardCells_R=10, T_CFAR=8.0,
                         method="GOCA")                                            # v8_3:45-50
    if point_PRT is not None and sc.prtNum == 32:
        cfar_params.guardCells_V, cfar_params.refCells_V = 2, 4
    cluster_params = Struct(max_range_sep=30.0, max_vel_sep=0.4, max_angle_sep=5.0)   # v8_3:52-54
    return config, cfar_params, cluster_params


NAMED_SHAPES = {            # BASELINE.json configs + the reference's literal shape
    "native": dict(channel_num=16, beam_num=13, prtNum=332, point_PRT=None),
    "cfg1": dict(channel_num=16, beam_num=13, prtNum=32, point_PRT=4096),
    "cfg2": dict(channel_num=16, beam_num=8, prtNum=64, point_PRT=8192),
    "cfg3": dict(channel_num=32, beam_num=16, prtNum=128, point_PRT=16384),
}


def named_config(name: str):
    return default_config(**NAMED_SHAPES[name])


def read_dbf_csv(path: str = DBF_CSV) -> np.ndarray:
    """``readmatrix`` + ``D(:,1:2:end) + 1j*D(:,2:2:end)`` (v8_3:182-183)."""
    with open(path, "rb") as fh:
        text = fh.read().decode("ascii")
    rows = []
    for line in text.replace("\r", "").split("\n"):
        cells = [x for x in line.split(",") if x.strip()]
        if cells:
            rows.append([float(x) for x in cells])
    d = np.asarray(rows, dtype=np.float64)
    return d[:, 0::2] + 1j * d[:, 1::2]


def dbf_tables(config):
    """(W [B,C], beam_angles_deg [B], k_slopes_LUT [B-1]).  16 channels and <= 13 beams: rows of the
    shipped CSV and the literal tables (v8_3:178-183).  Other shapes have no reference data:
    Hamming-tapered steering vectors to angles uniform in [-16, 60] deg, k = -4.7."""
    sc = config.Sig_Config
    C, B = sc.channel_num, sc.beam_num
    if C == 16 and B <= 13:
        return read_dbf_csv()[:B], np.array(BEAM_ANGLES_DEG[:B], float), np.array(K_SLOPES_LUT[:B - 1], float)
    angles = np.linspace(-16.0, 60.0, B)
    taper = np.hamming(C)
    dphi = 2 * np.pi * config.Array.element_spacing * np.sin(np.deg2rad(angles)) / sc.wavelength
    W = taper[None, :] * np.exp(1j * np.arange(C)[None, :] * dphi[:, None]) / taper.sum()
    return W, angles, np.full(B - 1, -4.7)


def _grpdelay_mean(b: np.ndarray, n: int = 512) -> float:
    """mean(grpdelay(b)) over n frequencies in [0, pi) (MATLAB default), via Re{FFT(k.b)/FFT(b)}."""
    k = np.arange(len(b))
    H = np.fft.fft(b, 2 * n)[:n]
    Hk = np.fft.fft(k * b, 2 * n)[:n]
    ok = np.abs(H) > 1e-12 * np.abs(H).max()
    gd = np.zeros(n)
    gd[ok] = np.real(Hk[ok] / H[ok])
    return float(gd.mean())


def build_precomputed_data(config) -> Struct:
    """The ``precomputed_data`` struct of v8_3:90-188 (same field names)."""
    sc = config.Sig_Config
    fs, N, P = sc.fs, sc.point_PRT, sc.prtNum
    tau1, tau2, tau3 = sc.tao
    gap1, gap2 = sc.gap_duration[0], sc.gap_duration[1]
    k2, k3 = -sc.B / tau2, sc.B / tau3                                       # v8_3:123
    ns1, ns2, ns3 = (int(round(t * fs)) for t in (tau1, tau2, tau3))         # v8_3:124-126
    t2 = np.linspace(-tau2 / 2, tau2 / 2, ns2)
    t3 = np.linspace(-tau3 / 2, tau3 / 2, ns3)
    pulse2 = np.exp(1j * 2 * np.pi * (0.5 * k2 * t2 ** 2))                   # v8_3:130-131
    pulse3 = np.exp(1j * 2 * np.pi * (0.5 * k3 * t3 ** 2))
    tx_pulse = np.zeros(N, dtype=np.complex128)                              # v8_3:132-137
    tx_pulse[:ns1] = 1.0
    offset1 = int(round((tau1 + gap1) * fs))
    tx_pulse[offset1:offset1 + ns2] = pulse2
    offset2 = offset1 + int(round((tau2 + gap2) * fs))
    tx_pulse[offset2:offset2 + ns3] = pulse3
    pd = Struct()
    pd.T_frame = P * sc.prt                                                  # v8_3:93
    pd.tx_pulse = tx_pulse
    pd.P_signal_unscaled = float(np.mean(np.abs(tx_pulse[tx_pulse != 0]) ** 2))   # v8_3:139
    fir = np.asarray(FIR_COEFFS, dtype=np.float64)
    fir = 6 * fir / fir.max()                                                # v8_3:142
    pd.MF_narrow = fir
    pd.fir_delay = int(math.floor(_grpdelay_mean(fir) + 0.5))                # v8_3:144
    pd.MF_medium_win = np.conj(pulse2 * np.kaiser(ns2, 4.5))[::-1].copy()    # v8_3:145-148
    pd.MF_long_win = np.conj(pulse3 * np.kaiser(ns3, 4.5))[::-1].copy()
    gap1_num, gap2_num = gap1 * fs, gap2 * fs                                # v8_3:152-153 (not rounded)
    seg_m = ns1 + gap1_num + ns2 + 1
    seg_l = ns1 + gap1_num + ns2 + gap2_num + ns3 + 1
    if seg_m != int(seg_m) or seg_l != int(seg_l):
        raise ValueError("segment starts are not integers (the reference uses them as indices)")
    seg_m, seg_l = int(seg_m), int(seg_l)
    pd.N_fft_med = 2 ** int(math.ceil(math.log2((N - seg_m + 1) + ns2 - 1)))   # v8_3:158-159
    pd.N_fft_long = 2 ** int(math.ceil(math.log2((N - seg_l + 1) + ns3 - 1)))
    pd.MF_medium_fft = np.fft.fft(pd.MF_medium_win, pd.N_fft_med)            # v8_3:160-161
    pd.MF_long_fft = np.fft.fft(pd.MF_long_win, pd.N_fft_long)
    g = [int(x) for x in sc.point_prt_segments]
    pd.N_gate_narrow, pd.N_gate_medium, pd.N_gate_long = g                   # v8_3:163-165
    pd.N_total_gate = sum(g)
    pd.seg_start_narrow, pd.seg_start_medium, pd.seg_start_long = ns1 + 1, seg_m, seg_l   # v8_3:167-169
    pd.MTD_win = np.kaiser(P, 4.5)                                           # v8_3:171
    v_max = sc.wavelength / (2 * sc.prt)                                     # v8_3:173
    pd.velocity_axis = np.linspace(-v_max / 2, v_max / 2, P)                 # v8_3:174
    pd.range_axis = np.arange(pd.N_total_gate) * (sc.c / (2 * fs))           # v8_3:175
    pd.deltaR = sc.c * (1.0 / fs) / 2                                        # v8_3:176
    pd.deltaV = v_max / P                                                    # v8_3:177
    W, angles, slopes = dbf_tables(config)
    pd.beam_angles_deg = angles                                              # v8_3:178
    pd.k_slopes_LUT = slopes                                                 # v8_3:179
    pd.DBF_coeffs_data_C = W                                                 # v8_3:183
    return pd
