"""Host-side mirror of the reference's per-frame interface on top of the C ABI (include/rsp.h).

    final_targets = fun_process_single_frame(targets, config, cfar_params, cluster_params,
                                             precomputed_data, frame_idx)        # fun_process_single_frame.m:13

keeps the reference's argument names, meaning and "[] when nothing is detected" behaviour.  All of
S5..S9 (DBF, pulse compression, MTD, CFAR, monopulse) runs in librsp.so's CUDA kernels; S10/S11
(clustering, order-dependent BFS over a few hundred detections) runs in the library's host C++.
Nothing in this module computes the chain on the CPU.
"""
from __future__ import annotations

import ctypes as C
import math
from typing import Iterable, List, Optional, Sequence

import os
import numpy as np

from . import _abi
from ._abi import DETECTION_DTYPE, TARGET_DTYPE


def _as_c128(a) -> np.ndarray:
    return np.ascontiguousarray(np.asarray(a, dtype=np.complex128))


def _as_f64(a) -> np.ndarray:
    return np.ascontiguousarray(np.asarray(a, dtype=np.float64))


def _field(obj, name):
    return obj[name] if isinstance(obj, dict) else getattr(obj, name)


def _ptr_of(x):
    """(pointer, is_device) of a numpy array or a torch tensor."""
    if hasattr(x, "data_ptr"):                       # torch tensor
        return C.c_void_p(x.data_ptr()), bool(x.is_cuda)
    return C.c_void_p(x.ctypes.data), False


class RadarChain:
    """One librsp context: shapes + constants resident on one GPU, one CUDA stream."""

    def __init__(self, config, cfar_params, precomputed_data, device: int = 0, max_detections: int = 8192,
                 monopulse_complex: bool = False):
        self._lib = _abi.load()
        self._ctx = C.c_void_p()
        sc = _field(config, "Sig_Config")
        pd = precomputed_data
        self.C, self.B = int(_field(sc, "channel_num")), int(_field(sc, "beam_num"))
        self.P, self.N = int(_field(sc, "prtNum")), int(_field(sc, "point_PRT"))
        gates = [int(_field(pd, "N_gate_narrow")), int(_field(pd, "N_gate_medium")), int(_field(pd, "N_gate_long"))]
        self.G = sum(gates)
        if int(_field(pd, "N_total_gate")) != self.G:
            raise ValueError("N_total_gate does not equal the sum of the three gate counts")
        self.max_detections = int(max_detections)
        p = _abi.rsp_params()
        p.abi_version = _abi.RSP_ABI_VERSION
        p.n_channels, p.n_beams, p.n_pulses, p.n_samples = self.C, self.B, self.P, self.N
        p.seg_start[:] = [int(_field(pd, "seg_start_narrow")), int(_field(pd, "seg_start_medium")),
                          int(_field(pd, "seg_start_long"))]
        p.n_gates[:] = gates
        p.fir_delay = int(_field(pd, "fir_delay"))
        p.t_cfar = float(_field(cfar_params, "T_CFAR"))
        p.guard_r, p.guard_v = int(_field(cfar_params, "guardCells_R")), int(_field(cfar_params, "guardCells_V"))
        p.ref_r, p.ref_v = int(_field(cfar_params, "refCells_R")), int(_field(cfar_params, "refCells_V"))
        p.max_detections = self.max_detections
        p.monopulse_complex = int(bool(monopulse_complex))
        p.device = int(device)
        self._device = int(device)
        _abi.check(self._lib.rsp_create(C.byref(p), C.byref(self._ctx)))
        self._upload(pd)

    # -- set-up -------------------------------------------------------------------------------
    def _upload(self, pd):
        W = _as_c128(_field(pd, "DBF_coeffs_data_C"))
        if W.shape != (self.B, self.C):
            raise ValueError(f"DBF_coeffs_data_C must be [{self.B},{self.C}], got {W.shape}")
        fir = _as_f64(_field(pd, "MF_narrow"))
        try:
            mf_m, mf_l = _as_c128(_field(pd, "MF_medium_win")), _as_c128(_field(pd, "MF_long_win"))
        except (KeyError, AttributeError):
            # only the spectra were supplied (the fields fun_process_single_frame.m:26-27 reads):
            # the taps are the leading non-zero part of their inverse transform.
            mf_m = _taps_from_spectrum(_field(pd, "MF_medium_fft"))
            mf_l = _taps_from_spectrum(_field(pd, "MF_long_fft"))
        win, ra, va = _as_f64(_field(pd, "MTD_win")).ravel(), _as_f64(_field(pd, "range_axis")).ravel(), \
            _as_f64(_field(pd, "velocity_axis")).ravel()
        ang, ks = _as_f64(_field(pd, "beam_angles_deg")).ravel(), _as_f64(_field(pd, "k_slopes_LUT")).ravel()
        if len(win) != self.P or len(va) != self.P or len(ra) != self.G or len(ang) < self.B or len(ks) < self.B - 1:
            raise ValueError("precomputed_data table sizes do not match the configuration")
        k = _abi.rsp_constants()
        keep = (W, fir, mf_m, mf_l, win, ra, va, ang, ks)        # keep alive during the call
        k.dbf_weights, k.fir, k.n_fir = W.ctypes.data, fir.ctypes.data, len(fir)
        k.mf_medium, k.n_mf_medium = mf_m.ctypes.data, len(mf_m)
        k.mf_long, k.n_mf_long = mf_l.ctypes.data, len(mf_l)
        k.mtd_win, k.range_axis, k.velocity_axis = win.ctypes.data, ra.ctypes.data, va.ctypes.data
        k.delta_r, k.delta_v = float(_field(pd, "deltaR")), float(_field(pd, "deltaV"))
        k.beam_angles_deg, k.k_slopes = ang.ctypes.data, ks.ctypes.data
        _abi.check(self._lib.rsp_upload_constants(self._ctx, C.byref(k)), self._ctx)
        del keep

    def close(self):
        if getattr(self, "_ctx", None) and self._ctx.value:
            self._lib.rsp_destroy(self._ctx)
            self._ctx = C.c_void_p()

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_stream(self, cuda_stream_handle: int):
        _abi.check(self._lib.rsp_set_stream(self._ctx, C.c_void_p(cuda_stream_handle)), self._ctx)

    def synchronize(self):
        _abi.check(self._lib.rsp_synchronize(self._ctx), self._ctx)

    def info(self) -> dict:
        i = _abi.rsp_info()
        _abi.check(self._lib.rsp_get_info(self._ctx, C.byref(i)), self._ctx)
        return {f: getattr(i, f) for f, _ in i._fields_}

    def set_profiling(self, enable: bool):
        _abi.check(self._lib.rsp_set_profiling(self._ctx, int(bool(enable))), self._ctx)

    def kernel_times(self) -> dict:
        """{kernel class: (total device ms, launches)} since the last call (profiling must be on)."""
        kt = _abi.rsp_kernel_times()
        _abi.check(self._lib.rsp_get_kernel_times(self._ctx, C.byref(kt)), self._ctx)
        return {kt.name[i].decode(): (kt.total_ms[i], kt.launches[i]) for i in range(kt.n) if kt.launches[i]}

    def fused_trace(self) -> np.ndarray:
        """[ctas, 8] int64 phase timestamps of the last dbf_pc launch (context created under RSP_FUSED_DEBUG)."""
        buf = np.zeros((self.P * self.B, 8), np.int64)
        n = C.c_int32(0)
        _abi.check(self._lib.rsp_get_fused_trace(self._ctx, buf.ctypes.data, buf.shape[0], C.byref(n)), self._ctx)
        return buf[: n.value]

    # -- S5..S9 ---------------------------------------------------------------------------------
    def process_cpi(self, raw, layout: str = "pcn", rdm_out=None) -> np.ndarray:
        """Run DBF -> PC -> MTD -> CFAR -> monopulse on one cube.

        raw: numpy array or torch tensor (host or CUDA), complex64 or complex128;
             layout "pcn" = raw[p][c][n]; "matlab" = MATLAB [P,N,C] column-major bytes.
        rdm_out: optional numpy array / CUDA tensor receiving the complex64 map [B][G][P].
        Returns the detection table (structured array, reference order, 1-based indices)."""
        n_el = self.P * self.C * self.N
        ptr, on_dev = _ptr_of(raw)
        dt = str(raw.dtype).replace("torch.", "")
        if dt not in ("complex64", "complex128"):
            raise TypeError(f"raw must be complex64/complex128, got {raw.dtype}")
        size = raw.numel() if hasattr(raw, "numel") else raw.size
        if size != n_el:
            raise ValueError(f"raw has {size} elements, expected P*C*N = {n_el}")
        if hasattr(raw, "is_contiguous"):
            if not raw.is_contiguous():
                raise ValueError("raw must be contiguous")
        elif not raw.flags["C_CONTIGUOUS"]:
            raise ValueError("raw must be C-contiguous")
        lay = {"pcn": _abi.RSP_LAYOUT_PCN, "matlab": _abi.RSP_LAYOUT_MATLAB}[layout]
        dets = np.zeros(self.max_detections, dtype=DETECTION_DTYPE)
        n = C.c_int32(0)
        rptr, rmem = C.c_void_p(), _abi.RSP_MEM_DEVICE
        if rdm_out is not None:
            rptr, rdev = _ptr_of(rdm_out)
            rmem = _abi.RSP_MEM_DEVICE if rdev else _abi.RSP_MEM_HOST
        _abi.check(self._lib.rsp_process_cpi(
            self._ctx, ptr, lay, _abi.RSP_C64 if dt == "complex64" else _abi.RSP_C128,
            _abi.RSP_MEM_DEVICE if on_dev else _abi.RSP_MEM_HOST, rptr, rmem,
            C.c_void_p(dets.ctypes.data), self.max_detections, C.byref(n)), self._ctx)
        return dets[:n.value].copy()

    # -- S4 on the device ----------------------------------------------------------------------------
    def set_waveform(self, config, precomputed_data):
        """Upload tx_pulse and the scalars of fun_process_single_frame.m:47-77 (needed once)."""
        sc = _field(config, "Sig_Config")
        tx = _as_c128(_field(precomputed_data, "tx_pulse")).ravel()
        if len(tx) != self.N:
            raise ValueError("tx_pulse length must equal point_PRT")
        w = _abi.rsp_waveform()
        w.tx_pulse = tx.ctypes.data
        w.c, w.fs, w.wavelength, w.prt = (float(_field(sc, k)) for k in ("c", "fs", "wavelength", "prt"))
        w.element_spacing = float(_field(_field(config, "Array"), "element_spacing"))
        w.p_signal_unscaled = float(_field(precomputed_data, "P_signal_unscaled"))
        _abi.check(self._lib.rsp_set_waveform(self._ctx, C.byref(w)), self._ctx)
        self._has_waveform = True

    @staticmethod
    def _pack_targets(targets):
        arr = (_abi.rsp_target_in * max(len(targets), 1))()
        for i, t in enumerate(targets):
            arr[i].range, arr[i].velocity = float(_field(t, "Range")), float(_field(t, "Velocity"))
            arr[i].elevation_deg, arr[i].snr_db = float(_field(t, "ElevationAngle")), float(_field(t, "SNR_dB"))
        return arr

    def synthesize(self, targets, noise_power: float = 1.0, seed: int = 0, out=None):
        """S4 + S4.1 on the GPU.  ``out``: CUDA tensor [P, C, N] complex64 to fill (default: the context's
        own cube, which process_targets then consumes)."""
        arr = self._pack_targets(targets)
        optr = C.c_void_p(out.data_ptr()) if out is not None else C.c_void_p()
        _abi.check(self._lib.rsp_synthesize(self._ctx, arr, len(targets), float(noise_power), int(seed) & (2 ** 64 - 1), optr),
                   self._ctx)
        return out

    def process_targets(self, targets, cluster_params, noise_power: float = 1.0, seed: int = 0):
        """fun_process_single_frame on the device end to end -> (final targets, detections)."""
        arr = self._pack_targets(targets)
        cp = _abi.rsp_cluster_params(float(_field(cluster_params, "max_range_sep")), float(_field(cluster_params, "max_vel_sep")),
                                     float(_field(cluster_params, "max_angle_sep")))
        fin = np.zeros(4096, dtype=TARGET_DTYPE)
        dets = np.zeros(self.max_detections, dtype=DETECTION_DTYPE)
        nf, nd = C.c_int32(0), C.c_int32(0)
        _abi.check(self._lib.rsp_process_targets(self._ctx, arr, len(targets), float(noise_power), int(seed) & (2 ** 64 - 1),
                                                 C.byref(cp), C.c_void_p(fin.ctypes.data), len(fin), C.byref(nf),
                                                 C.c_void_p(dets.ctypes.data), len(dets), C.byref(nd)), self._ctx)
        return fin[:nf.value].copy(), dets[:nd.value].copy()

    def submit_targets(self, targets, slot: int, noise_power: float = 1.0, seed: int = 0):
        """Pipelined frame path: enqueue S4 (device synthesis) + S5..S9 of one frame on lane slot % lanes and
        return at once; collect with fetch_targets(slot)."""
        arr = self._pack_targets(targets)
        _abi.check(self._lib.rsp_submit_targets(self._ctx, arr, len(targets), float(noise_power), int(seed) & (2 ** 64 - 1),
                                                int(slot)), self._ctx)

    def fetch_targets(self, slot: int, cluster_params):
        """Wait for a submitted frame -> (final targets, detections)."""
        cp = _abi.rsp_cluster_params(float(_field(cluster_params, "max_range_sep")), float(_field(cluster_params, "max_vel_sep")),
                                     float(_field(cluster_params, "max_angle_sep")))
        fin, dets, pf, pd_ = self._result_buffers()
        nf, nd = C.c_int32(0), C.c_int32(0)
        _abi.check(self._lib.rsp_fetch_targets(self._ctx, int(slot), C.byref(cp), pf, len(fin), C.byref(nf),
                                               pd_, len(dets), C.byref(nd)), self._ctx)
        return fin[:nf.value].copy(), dets[:nd.value].copy()

    def _result_buffers(self):
        """Reusable host buffers for one frame's targets and detections (allocating them per call costs more
        than the fetch itself)."""
        b = getattr(self, "_res_buf", None)
        if b is None:
            fin = np.empty(4096, dtype=TARGET_DTYPE)
            dets = np.empty(self.max_detections, dtype=DETECTION_DTYPE)
            b = self._res_buf = (fin, dets, C.c_void_p(fin.ctypes.data), C.c_void_p(dets.ctypes.data))
        return b

    def process_targets_batch(self, target_lists, cluster_params, noise_power: float = 1.0, seeds=None, depth: int = 0,
                              host_threads: int = 0, return_detections: bool = True, native: bool = True):
        """Many independent frames (Monte-Carlo trials, a block of a frame stream), pipelined `depth` deep over
        the lanes: frame i+1 is synthesised on the GPU while frame i runs S5..S9, and only target lists go in
        and detection lists come back.  Returns [(final targets, detections), ...] in input order (detections None
        with ``return_detections=False`` -- the reference function returns final_targets only).

        ``native`` (default): one rsp_process_frames call per block of 256 frames -- submission, fetch and a pool of
        ``host_threads`` workers for sorting + clustering all inside librsp, so the host half of a dense frame does
        not hold the GPU back.  ``native=False``: the same pipeline driven frame by frame from Python
        (rsp_submit_targets / rsp_fetch_targets)."""
        n = len(target_lists)
        seeds = list(range(n)) if seeds is None else list(seeds)
        if native:
            return self._process_frames_native(target_lists, cluster_params, noise_power, seeds, depth, host_threads, return_detections)
        slots = self.stream_slots()
        depth = min(depth or 2 * max(self.info()["lanes"], 1), slots)
        out = []
        for i in range(n + depth):
            if i >= depth:
                out.append(self.fetch_targets((i - depth) % slots, cluster_params))
            if i < n:
                self.submit_targets(target_lists[i], i % slots, noise_power, seeds[i])
        return out[:n]

    _FRAME_BLOCK = 256           # frames per rsp_process_frames call
    _FRAME_TARGET_CAP = 512      # final targets per frame
    _FRAME_DET_CAP = 4096        # detections per frame the flat buffer is sized for (on average)

    def _frame_buffers(self, with_detections: bool):
        b = getattr(self, "_frames_buf", None)
        if b is None:
            b = self._frames_buf = [np.empty((self._FRAME_BLOCK, self._FRAME_TARGET_CAP), dtype=TARGET_DTYPE), None]
        if with_detections and b[1] is None:
            b[1] = np.empty(self._FRAME_BLOCK * min(self.max_detections, self._FRAME_DET_CAP), dtype=DETECTION_DTYPE)
        return b[0], b[1]

    def _process_frames_native(self, target_lists, cluster_params, noise_power, seeds, depth, host_threads, return_detections):
        cp = _abi.rsp_cluster_params(float(_field(cluster_params, "max_range_sep")), float(_field(cluster_params, "max_vel_sep")),
                                     float(_field(cluster_params, "max_angle_sep")))
        if host_threads <= 0:
            host_threads = max(2, min(8, (os.cpu_count() or 4) - 1))
        out = []
        packed = {}                                   # a list object that appears many times (one scene, many noise seeds) is packed once
        for b0 in range(0, len(target_lists), self._FRAME_BLOCK):
            lists = target_lists[b0:b0 + self._FRAME_BLOCK]
            nb = len(lists)
            rows = []
            for tl in lists:
                r = packed.get(id(tl))
                if r is None:
                    r = np.array([(float(_field(t, "Range")), float(_field(t, "Velocity")), float(_field(t, "ElevationAngle")),
                                   float(_field(t, "SNR_dB"))) for t in tl], dtype=np.float64).reshape(len(tl), 4)
                    packed[id(tl)] = r
                rows.append(r)
            n_tg = np.array([len(r) for r in rows], dtype=np.int32)
            tg = np.ascontiguousarray(np.concatenate(rows, axis=0)) if n_tg.sum() else np.zeros((1, 4), dtype=np.float64)
            sd = np.array([int(x) & (2 ** 64 - 1) for x in seeds[b0:b0 + nb]], dtype=np.uint64)
            fin, dets_all = self._frame_buffers(return_detections)     # reused from call to call (42 MB: first touch costs 10 ms)
            n_fin = np.zeros(nb, dtype=np.int32)
            if return_detections:
                cap_total = nb * min(self.max_detections, self._FRAME_DET_CAP)
                dets = dets_all[:cap_total]
                offs = np.zeros(nb + 1, dtype=np.int64)
                pd_, po = C.c_void_p(dets.ctypes.data), C.c_void_p(offs.ctypes.data)
            else:
                cap_total, dets, offs, pd_, po = 0, None, None, C.c_void_p(), C.c_void_p()
            _abi.check(self._lib.rsp_process_frames(self._ctx, C.c_void_p(tg.ctypes.data), C.c_void_p(n_tg.ctypes.data), nb,
                                                    float(noise_power), C.c_void_p(sd.ctypes.data), C.byref(cp), int(depth),
                                                    int(host_threads), C.c_void_p(fin.ctypes.data), self._FRAME_TARGET_CAP,
                                                    C.c_void_p(n_fin.ctypes.data), pd_, cap_total, po), self._ctx)
            # one copy of the block's results out of the reused buffers; the per-frame arrays are views into it
            nf = n_fin.tolist()
            fin_b = fin[:nb, :max(nf) if nf else 0].copy()
            if return_detections:
                o = offs.tolist()
                dets_b = dets[:o[nb]].copy()
                out.extend((fin_b[i, :nf[i]], dets_b[o[i]:o[i + 1]]) for i in range(nb))
            else:
                out.extend((fin_b[i, :nf[i]], None) for i in range(nb))
        return out

    # -- device-resident stream -----------------------------------------------------------------
    def stream_slots(self) -> int:
        return int(self._lib.rsp_stream_slots(self._ctx))

    def stream_enqueue(self, raw_dev_ptr: int, raw_pool: int, rdm_dev_ptr: int, rdm_pool: int, n_cpi: int,
                       first_slot: int = 0):
        _abi.check(self._lib.rsp_stream_enqueue(self._ctx, C.c_void_p(raw_dev_ptr), raw_pool,
                                                C.c_void_p(rdm_dev_ptr) if rdm_dev_ptr else C.c_void_p(), rdm_pool,
                                                n_cpi, first_slot), self._ctx)

    def submit_cpi(self, raw_host, slot: int, rdm_dev_ptr: int = 0):
        """Pipelined host-input path: enqueue H2D + chain for one pinned PCN complex64 cube; collect with
        stream_fetch(slot).  The caller must keep ``raw_host`` alive until the fetch."""
        ptr, on_dev = _ptr_of(raw_host)
        if on_dev:
            raise ValueError("submit_cpi takes a host cube")
        _abi.check(self._lib.rsp_submit_cpi(self._ctx, ptr, C.c_void_p(rdm_dev_ptr) if rdm_dev_ptr else C.c_void_p(), slot),
                   self._ctx)

    def stream_fetch(self, slot: int) -> np.ndarray:
        _, dets, _, pd_ = self._result_buffers()
        n = C.c_int32(0)
        _abi.check(self._lib.rsp_stream_fetch(self._ctx, slot, pd_, self.max_detections, C.byref(n)), self._ctx)
        return dets[:n.value].copy()

    def stream_device_buffers(self):
        a, b = C.c_void_p(), C.c_void_p()
        _abi.check(self._lib.rsp_stream_device_buffers(self._ctx, C.byref(a), C.byref(b)), self._ctx)
        return a.value, b.value

    # -- intermediates (parity) -------------------------------------------------------------------
    def get_beam(self) -> np.ndarray:
        out = np.empty((self.P, self.B, self.N), np.complex64)
        _abi.check(self._lib.rsp_get_beam(self._ctx, C.c_void_p(out.ctypes.data)), self._ctx)
        return out

    def get_pc(self) -> np.ndarray:
        out = np.empty((self.P, self.B, self.G), np.complex64)
        _abi.check(self._lib.rsp_get_pc(self._ctx, C.c_void_p(out.ctypes.data)), self._ctx)
        return out

    def get_rdm(self) -> np.ndarray:
        out = np.empty((self.B, self.G, self.P), np.complex64)
        _abi.check(self._lib.rsp_get_rdm(self._ctx, C.c_void_p(out.ctypes.data)), self._ctx)
        return out

    def get_amp(self) -> np.ndarray:
        out = np.empty((self.B, self.G, self.P), np.float32)
        _abi.check(self._lib.rsp_get_amp(self._ctx, C.c_void_p(out.ctypes.data)), self._ctx)
        return out


def _taps_from_spectrum(spec) -> np.ndarray:
    h = np.fft.ifft(np.asarray(spec, dtype=np.complex128))
    nz = np.nonzero(np.abs(h) > 1e-9 * np.abs(h).max())[0]
    return np.ascontiguousarray(h[: int(nz.max()) + 1])


def sort_detections(dets: np.ndarray) -> np.ndarray:
    """Reference order: pair ascending, then range, then Doppler (MATLAB find, fsf:215-221)."""
    d = np.ascontiguousarray(dets, dtype=DETECTION_DTYPE)
    _abi.check(_abi.load().rsp_sort_detections(C.c_void_p(d.ctypes.data), len(d)))
    return d


def cluster(dets: np.ndarray, cluster_params):
    """S10 + S11 (fun_process_single_frame.m:302-407) -> (stage1, final) structured arrays."""
    lib = _abi.load()
    d = np.ascontiguousarray(dets, dtype=DETECTION_DTYPE)
    n = len(d)
    cp = _abi.rsp_cluster_params(float(_field(cluster_params, "max_range_sep")),
                                 float(_field(cluster_params, "max_vel_sep")),
                                 float(_field(cluster_params, "max_angle_sep")))
    s1 = np.zeros(max(n, 1), dtype=TARGET_DTYPE)
    fin = np.zeros(max(n, 1), dtype=TARGET_DTYPE)
    n1, nf = C.c_int32(0), C.c_int32(0)
    _abi.check(lib.rsp_cluster(C.c_void_p(d.ctypes.data), n, C.byref(cp), C.c_void_p(s1.ctypes.data), C.byref(n1),
                               C.c_void_p(fin.ctypes.data), C.byref(nf)))
    return s1[:n1.value].copy(), fin[:nf.value].copy()


# ------------------------------------------------------------------------------------------------
# S4 / S4.1 on the host (inputs of the chain; fun_process_single_frame.m:47-88)
# ------------------------------------------------------------------------------------------------
def _matlab_round(x: float) -> int:
    return int(math.floor(x + 0.5)) if x >= 0 else -int(math.floor(-x + 0.5))


def synthesize_echo(targets: Sequence, config, precomputed_data) -> np.ndarray:
    """Noise-free echoes raw[p][c][n] (complex128) for a list of targets with fields Range,
    Velocity, ElevationAngle, SNR_dB (extra fields are ignored), fun_process_single_frame.m:47-77."""
    sc = _field(config, "Sig_Config")
    P, N, Cn = int(_field(sc, "prtNum")), int(_field(sc, "point_PRT")), int(_field(sc, "channel_num"))
    c0, fs, lam, prt = (float(_field(sc, k)) for k in ("c", "fs", "wavelength", "prt"))
    d_el = float(_field(_field(config, "Array"), "element_spacing"))
    tx = np.asarray(_field(precomputed_data, "tx_pulse"))
    p_sig = float(_field(precomputed_data, "P_signal_unscaled"))
    raw = np.zeros((P, Cn, N), dtype=np.complex128)
    m = np.arange(P)
    for t in targets:
        ts = 1.0 / fs                                                  # fsf:18
        delay_samples = _matlab_round((2 * float(_field(t, "Range")) / c0) / ts)   # fsf:55-56: round(delay / ts), not delay * fs
        doppler = np.exp(1j * 2 * np.pi * (2 * float(_field(t, "Velocity")) / lam) * m * prt)
        amplitude = math.sqrt(10 ** (float(_field(t, "SNR_dB")) / 10) / p_sig)
        base = np.zeros(N, dtype=np.complex128)
        if 0 < delay_samples < N:
            n_echo = min(len(tx), N - delay_samples)
            base[delay_samples:delay_samples + n_echo] = tx[:n_echo]
        dphi = 2 * np.pi * d_el * math.sin(math.radians(float(_field(t, "ElevationAngle")))) / lam
        phasors = np.exp(1j * np.arange(Cn) * dphi)
        raw += (amplitude * doppler)[:, None, None] * phasors[None, :, None] * base[None, None, :]
    return raw


def add_noise(raw: np.ndarray, rng: np.random.Generator, P_noise_floor: float = 1.0) -> np.ndarray:
    """fun_process_single_frame.m:81-88 with a NumPy generator in place of MATLAB's randn."""
    noise = rng.standard_normal(raw.shape) + 1j * rng.standard_normal(raw.shape)
    return raw + noise * math.sqrt(P_noise_floor / 2)


_chain_cache = {}


def _context_key(config, cfar_params, precomputed_data, device: int):
    """What defines a device context: shape, CFAR parameters, device and the CONTENT of every uploaded table.  The reference
    re-reads its structs on every call, so a caller may change T_CFAR or swap a filter between two calls with the same
    objects; keying the cache on id() would silently reuse a stale context (and ids are recycled after garbage collection)."""
    import hashlib
    h = hashlib.blake2b(digest_size=16)
    sc = _field(config, "Sig_Config")
    for name in ("prtNum", "point_PRT", "channel_num", "beam_num", "fs", "c", "wavelength", "prt"):
        h.update(repr((name, _field(sc, name))).encode())
    h.update(repr(tuple(_field(sc, "point_prt_segments"))).encode())
    h.update(repr(float(_field(_field(config, "Array"), "element_spacing"))).encode())
    for name in ("T_CFAR", "guardCells_R", "guardCells_V", "refCells_R", "refCells_V"):
        h.update(repr((name, _field(cfar_params, name))).encode())
    for name in sorted(_fields_of(precomputed_data)):
        v = _field(precomputed_data, name)
        h.update(name.encode())
        h.update(np.ascontiguousarray(np.asarray(v)).tobytes() if not isinstance(v, (str, bytes)) else str(v).encode())
    return (int(device), h.hexdigest())


def _fields_of(obj):
    if isinstance(obj, dict):
        return list(obj.keys())
    return [k for k in vars(obj) if not k.startswith("_")]


def _cached_chain(config, cfar_params, precomputed_data, device: int):
    key = _context_key(config, cfar_params, precomputed_data, device)
    chain = _chain_cache.get(key)
    if chain is None:
        for old in _chain_cache.values():
            old.close()
        _chain_cache.clear()
        chain = _chain_cache[key] = RadarChain(config, cfar_params, precomputed_data, device=device)
    return chain


def fun_process_single_frame(targets, config, cfar_params, cluster_params, precomputed_data, frame_idx=1, *,
                             rng: Optional[np.random.Generator] = None, noise: bool = True,
                             chain: Optional[RadarChain] = None, device: int = 0, host_synthesis: bool = False) -> List[dict]:
    """Drop-in for fun_process_single_frame.m:13.  Returns a list of dicts with the reference's
    fields Range, Velocity, Angle, Power (stage-2 cluster order); ``[]`` when nothing is detected.

    Default: S4 (echo synthesis + noise) runs on the GPU (rsp_process_targets), so only the target list
    crosses the bus; the noise comes from the device Philox generator seeded from ``rng`` (the reference
    draws from MATLAB's global randn stream, which cannot be reproduced).  ``host_synthesis=True`` builds
    the cube with NumPy (``rng`` noise) and hands it to the device chain instead."""
    if chain is None:
        chain = _cached_chain(config, cfar_params, precomputed_data, device)
    if host_synthesis:
        raw = synthesize_echo(targets, config, precomputed_data)
        if noise:
            raw = add_noise(raw, rng if rng is not None else np.random.default_rng())
        dets = chain.process_cpi(np.ascontiguousarray(raw.astype(np.complex64)))
        _, final = cluster(dets, cluster_params)
    else:
        if not getattr(chain, "_has_waveform", False):
            chain.set_waveform(config, precomputed_data)
        seed = int((rng if rng is not None else np.random.default_rng()).integers(0, 2 ** 63))
        final, _ = chain.process_targets(list(targets), cluster_params, 1.0 if noise else 0.0, seed)
    return [dict(Range=float(t["range"]), Velocity=float(t["velocity"]), Angle=float(t["angle"]),
                 Power=float(t["power"])) for t in final]


def fun_process_frames(target_lists, config, cfar_params, cluster_params, precomputed_data, first_frame_idx=1, *,
                       rng: Optional[np.random.Generator] = None, noise: bool = True,
                       chain: Optional[RadarChain] = None, device: int = 0) -> List[List[dict]]:
    """fun_process_single_frame for a whole block of independent frames at once (the frame loop of
    main_simulate_echoes_with_array_v8_3.m:200-248, whose target kinematics do not depend on the detections):
    same per-frame results as calling fun_process_single_frame frame by frame with the same ``rng`` (one seed
    is drawn per frame, in order), but the frames are pipelined over the device lanes."""
    if chain is None:
        chain = _cached_chain(config, cfar_params, precomputed_data, device)
    if not getattr(chain, "_has_waveform", False):
        chain.set_waveform(config, precomputed_data)
    rng = rng if rng is not None else np.random.default_rng()
    seeds = [int(rng.integers(0, 2 ** 63)) for _ in target_lists]
    res = chain.process_targets_batch([list(t) for t in target_lists], cluster_params, 1.0 if noise else 0.0, seeds)
    return [[dict(Range=float(t["range"]), Velocity=float(t["velocity"]), Angle=float(t["angle"]), Power=float(t["power"]))
             for t in final] for final, _ in res]

