"""Host-side mirror of ``[MTD_results, PC_results] = process_stage2_mtd(iq_data, angle, config)``
(/root/reference/Simulation/process_stage2_mtd.m:1) over rsp_stage2_configure / rsp_stage2_mtd.

The per-beam callee ``fun_MTD_produce`` (and its three callees) are not shipped by the reference, so the
arithmetic below is SPECIFIED BY THIS REPO, following the only surviving copy of the
set-up code (debug_simulated_data_processing_v2.m:259-351) and the config schema of
main_test_with_simulated_data.m:46-140:

  * reference pulses   pulse1 = sin(2*pi*t1 + pi/2), pulse2/3 = exp(j*2*pi*(0.5*K*t.^2)),
                       t = -tau/2 : ts : tau/2 - ts                      (debug_..._v2.m:309-317)
  * per-segment matched filter (segments = config.Sig_Config.point_prt(2:4) gates of iq_data)
        pc(p, g, b) = sum_k iq(p, g + k, b) * conj(pulse_s(k)),   g, g + k inside segment s
  * Doppler            mtd(:, g, b) = fftshift(fft(pc(:, g, b) .* win)),  win = ones unless given
  * zero-velocity notch rows within +-config.cfar.MTD_0v_num of zero Doppler are cleared
                       (main_test_with_simulated_data.m:106,124); 0 / absent = off
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _abi
from .frame import _field


def reference_pulses(config):
    """The three reference pulses of debug_simulated_data_processing_v2.m:309-317."""
    sc = _field(config, "Sig_Config")
    fs = float(_field(sc, "fs"))
    ts = 1.0 / fs
    tao = [float(x) for x in _field(sc, "tao")]
    Bw = float(_field(sc, "B"))
    K2, K3 = -Bw / tao[1], Bw / tao[2]
    out = []
    for i, tau in enumerate(tao):
        n = int(round(tau / ts))
        t = -tau / 2 + ts * np.arange(n)                     # -tau/2 : ts : tau/2 - ts
        if i == 0:
            out.append(np.sin(2 * np.pi * t + np.pi / 2).astype(np.complex128))
        else:
            out.append(np.exp(1j * 2 * np.pi * (0.5 * (K2 if i == 1 else K3) * t ** 2)))
    return out


class Stage2Chain:
    def __init__(self, config, device: int = 0, mtd_win=None):
        self._lib = _abi.load()
        self._ctx = C.c_void_p()
        self._config = config
        sc = _field(config, "Sig_Config")
        mtd = config["mtd"] if "mtd" in config else sc
        self.P = int(_field(sc, "prtNum"))
        self.B = int(_field(mtd, "beam_num"))
        pts = [int(x) for x in _field(sc, "point_prt")]
        self.gates = pts[1:4]
        self.G = sum(self.gates)
        p = _abi.rsp_params()
        p.abi_version = _abi.RSP_ABI_VERSION
        p.n_channels, p.n_beams, p.n_pulses, p.n_samples = 1, self.B, self.P, max(self.G, 64)
        p.seg_start[:] = [1, 1, 1]
        p.n_gates[:] = self.gates
        p.fir_delay, p.t_cfar = 0, 8.0
        p.guard_r = p.guard_v = 1
        p.ref_r = p.ref_v = 1
        p.max_detections, p.monopulse_complex, p.device = 16, 0, int(device)
        _abi.check(self._lib.rsp_create(C.byref(p), C.byref(self._ctx)))
        pulses = [np.ascontiguousarray(x, dtype=np.complex128) for x in reference_pulses(config)]
        cfg = _abi.rsp_stage2_config()
        for i in range(3):
            cfg.pulse[i] = pulses[i].ctypes.data
            cfg.n_pulse[i] = len(pulses[i])
        win = None if mtd_win is None else np.ascontiguousarray(mtd_win, dtype=np.float64)
        cfg.mtd_win = win.ctypes.data if win is not None else None
        cfar = config["cfar"] if "cfar" in config else {}
        cfg.zero_vel_bins = int(cfar["MTD_0v_num"]) if "MTD_0v_num" in cfar else 0
        _abi.check(self._lib.rsp_stage2_configure(self._ctx, C.byref(cfg)), self._ctx)

    def __call__(self, iq_data: np.ndarray):
        """iq_data[p, g, b] (any strides, complex) -> (MTD_results, PC_results), both [P, G, B] complex128."""
        if iq_data.shape != (self.P, self.G, self.B):
            raise ValueError(f"iq_data must be {(self.P, self.G, self.B)}, got {iq_data.shape}")
        src = np.asfortranarray(iq_data.astype(np.complex128, copy=False))       # MATLAB byte order
        mtd = np.empty((self.P, self.G, self.B), np.complex128, order="F")
        pc = np.empty((self.P, self.G, self.B), np.complex128, order="F")
        _abi.check(self._lib.rsp_stage2_mtd(self._ctx, C.c_void_p(src.ctypes.data), _abi.RSP_C128,
                                            C.c_void_p(mtd.ctypes.data), C.c_void_p(pc.ctypes.data)), self._ctx)
        return mtd, pc

    def cfar(self, config=None):
        """The 1-D range CFAR (local_execute_cfar) on the Doppler maps the last call left on the device, all beams at once,
        without a host round trip of the maps.  Returns (cfar_flag, threshold_matrix) as [P, G, B] float64."""
        cfgd = config if config is not None else self._config
        p = _cfar1d_params(cfgd["cfar"], self.gates)
        flags = np.empty((self.B, self.G, self.P), np.uint8)
        thr = np.empty((self.B, self.G, self.P), np.float32)
        _abi.check(self._lib.rsp_stage2_cfar(self._ctx, C.byref(p), C.c_void_p(flags.ctypes.data), C.c_void_p(thr.ctypes.data)),
                   self._ctx)
        return np.transpose(flags, (2, 1, 0)).astype(np.float64), np.transpose(thr, (2, 1, 0)).astype(np.float64)

    def close(self):
        if getattr(self, "_ctx", None) and self._ctx.value:
            self._lib.rsp_destroy(self._ctx)
            self._ctx = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def _cfar1d_params(cfar, seg_len):
    p = _abi.rsp_cfar1d_params()
    p.ref_cells, p.save_cells = int(cfar["refCells_R"]), int(cfar["saveCells_R"])
    p.method = int(cfar["CFARmethod_R"]) if "CFARmethod_R" in cfar else 0
    p.zero_vel_bins = int(cfar["MTD_0v_num"]) if "MTD_0v_num" in cfar else 0
    p.t_cfar = float(cfar["T_CFAR"])
    p.seg_len[:] = [int(x) for x in seg_len]
    return p


def local_execute_cfar(mtd_amplitude_map, cfar_params, config, device: int = 0):
    """Drop-in for local_execute_cfar, debug_simulated_data_processing_v2.m:419: per-segment 1-D range GOCA / SOCA CFAR
    with the zero-velocity rows left out.  ``mtd_amplitude_map``: [V, R] (one beam) or [V, R, B]; ``config.cfar`` holds
    refCells_R, saveCells_R, T_CFAR, CFARmethod_R, MTD_0v_num and ``config.Sig_Config.point_prt`` = [total, narrow, medium,
    long] gates; like the reference, ``cfar_params`` is accepted and config.cfar is what is read (:427-429).
    Returns (cfar_flag, threshold_matrix), float64 arrays shaped like the input."""
    amp = np.asarray(mtd_amplitude_map)
    one = amp.ndim == 2
    if one:
        amp = amp[:, :, None]
    V, R, B = amp.shape
    seg = list(config["Sig_Config"]["point_prt"])[1:4]
    p = _cfar1d_params(config["cfar"], seg)
    dev_amp = np.ascontiguousarray(np.transpose(amp, (2, 1, 0)), dtype=np.float32)          # [B][R][V]
    flags = np.empty((B, R, V), np.uint8)
    thr = np.empty((B, R, V), np.float32)
    lib = _abi.load()
    _abi.check(lib.rsp_cfar1d(int(device), C.c_void_p(dev_amp.ctypes.data), V, R, B, C.byref(p), C.c_void_p(flags.ctypes.data),
                              C.c_void_p(thr.ctypes.data)))
    f = np.transpose(flags, (2, 1, 0)).astype(np.float64)
    t = np.transpose(thr, (2, 1, 0)).astype(np.float64)
    return (f[:, :, 0], t[:, :, 0]) if one else (f, t)


_cache = {}


def process_stage2_mtd(iq_data, angle, config, device: int = 0):
    """Drop-in for process_stage2_mtd.m:1.  ``angle`` is accepted and unused, like the reference
    (process_stage2_mtd.m:26)."""
    key = (int(device), _config_digest(config))      # content, not id(): the reference re-reads config on every call
    ch = _cache.get(key)
    if ch is None:
        for old in _cache.values():
            old.close()
        _cache.clear()
        ch = _cache[key] = Stage2Chain(config, device=device)
    return ch(np.asarray(iq_data))


def _config_digest(obj) -> str:
    """Stable digest of a (nested) struct of scalars / sequences / arrays."""
    import hashlib
    h = hashlib.blake2b(digest_size=16)

    def walk(x):
        if isinstance(x, dict):
            for k in sorted(x.keys()):
                h.update(str(k).encode())
                walk(x[k])
        elif isinstance(x, (list, tuple)):
            for v in x:
                walk(v)
        elif isinstance(x, np.ndarray):
            h.update(np.ascontiguousarray(x).tobytes())
        else:
            h.update(repr(x).encode())
    walk(obj)
    return h.hexdigest()
