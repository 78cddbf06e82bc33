"""The MEX gateways (mex/*.cpp) linked against librsp.so and EXECUTED against a stub MEX runtime.

MATLAB / Octave do not exist here, so the gateways cannot run under the real interpreter; mex/mex_shim/mex_stub_runtime.cpp
implements the few mx* / mex* functions they use (test infrastructure).  CPU tests: both gateways link, export mexFunction
and reject bad calls with the reference-style error ids before touching the device.  GPU test: the reference's argument
structs are built through the stub, mexFunction runs S4 in C++ + the stub's randn stream + S5..S11 on the device, and the
struct array it returns must equal what the ctypes path returns for the same cube."""
import ctypes as C
import importlib.util
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "radar-signal-simulation-and-target-detection_b200")
LIB_DIR = os.path.join(PKG, "lib")


def _build(gateway: str) -> str:
    """g++ the gateway + the stub runtime into lib/lib<gateway>_stub.so, linked against the in-tree librsp.so."""
    import __graft_entry__ as entry
    entry.build()
    out = os.path.join(LIB_DIR, f"lib{gateway}_stub.so")
    srcs = [os.path.join(ROOT, "mex", gateway + ".cpp"), os.path.join(ROOT, "mex", "mex_shim", "mex_stub_runtime.cpp")]
    deps = srcs + [os.path.join(ROOT, "mex", "rsp_mex_common.hpp"), os.path.join(ROOT, "include", "rsp.h"), os.path.join(LIB_DIR, "librsp.so")]
    if not os.path.exists(out) or any(os.path.getmtime(d) > os.path.getmtime(out) for d in deps):
        subprocess.check_call(["g++", "-std=c++17", "-O1", "-Wall", "-Werror", "-shared", "-fPIC", "-I", os.path.join(ROOT, "mex", "mex_shim"),
                               "-I", os.path.join(ROOT, "include"), "-I", os.path.join(ROOT, "mex")] + srcs +
                              ["-L", LIB_DIR, "-lrsp", "-Wl,-rpath," + LIB_DIR, "-o", out])
    return out


class MexStub:
    """ctypes view of one gateway + the stub runtime: builds mxArrays from Python values and calls mexFunction."""

    def __init__(self, gateway: str):
        self.lib = L = C.CDLL(_build(gateway))
        vp, sz = C.c_void_p, C.c_size_t
        L.mxCreateDoubleMatrix.restype = vp; L.mxCreateDoubleMatrix.argtypes = [sz, sz, C.c_int]
        L.mxCreateNumericArray.restype = vp; L.mxCreateNumericArray.argtypes = [sz, C.POINTER(sz), C.c_int, C.c_int]
        L.mxCreateStructMatrix.restype = vp; L.mxCreateStructMatrix.argtypes = [sz, sz, C.c_int, C.POINTER(C.c_char_p)]
        L.mxCreateDoubleScalar.restype = vp; L.mxCreateDoubleScalar.argtypes = [C.c_double]
        L.mxSetField.argtypes = [vp, sz, C.c_char_p, vp]
        L.mxGetField.restype = vp; L.mxGetField.argtypes = [vp, sz, C.c_char_p]
        L.mxGetDoubles.restype = C.POINTER(C.c_double); L.mxGetDoubles.argtypes = [vp]
        L.mxGetScalar.restype = C.c_double; L.mxGetScalar.argtypes = [vp]
        L.mxGetNumberOfElements.restype = sz; L.mxGetNumberOfElements.argtypes = [vp]
        L.mxGetNumberOfDimensions.restype = sz; L.mxGetNumberOfDimensions.argtypes = [vp]
        L.mxGetDimensions.restype = C.POINTER(sz); L.mxGetDimensions.argtypes = [vp]
        L.mxIsStruct.argtypes = [vp]; L.mxIsComplex.argtypes = [vp]
        L.mxDestroyArray.argtypes = [vp]
        L.stub_set_random_stream.argtypes = [C.POINTER(C.c_double), sz]
        L.stub_random_consumed.restype = sz
        L.stub_last_error.restype = C.c_char_p
        L.stub_printed.restype = C.c_char_p
        L.stub_call_mex.argtypes = [C.c_int, C.POINTER(vp), C.c_int, C.POINTER(vp)]

    def array(self, a) -> int:
        """NumPy array -> double (complex) mxArray with MATLAB's column-major element order."""
        a = np.asarray(a)
        cplx = np.iscomplexobj(a)
        a = a.astype(np.complex128 if cplx else np.float64)
        if a.ndim == 0:
            a = a.reshape(1, 1)
        if a.ndim == 1:
            a = a.reshape(1, -1)                    # row vector, like the reference's precomputed vectors
        dims = (C.c_size_t * a.ndim)(*a.shape)
        m = self.lib.mxCreateNumericArray(a.ndim, dims, 6, 1 if cplx else 0)
        flat = np.ascontiguousarray(a.ravel(order="F"))
        dst = self.lib.mxGetDoubles(m)
        C.memmove(dst, flat.ctypes.data, flat.nbytes)
        return m

    def struct(self, d) -> int:
        """dict (MATLAB struct look-alike) -> 1 x 1 struct mxArray, recursively; strings and None are skipped."""
        items = [(k, v) for k, v in d.items() if not isinstance(v, (str, type(None)))]
        names = (C.c_char_p * len(items))(*[k.encode() for k, _ in items])
        m = self.lib.mxCreateStructMatrix(1, 1, len(items), names)
        for k, v in items:
            self.lib.mxSetField(m, 0, k.encode(), self.value(v))
        return m

    def struct_array(self, dicts) -> int:
        keys = list(dicts[0].keys()) if dicts else []
        names = (C.c_char_p * len(keys))(*[k.encode() for k in keys])
        m = self.lib.mxCreateStructMatrix(1, len(dicts), len(keys), names)
        for i, d in enumerate(dicts):
            for k in keys:
                self.lib.mxSetField(m, i, k.encode(), self.value(d[k]))
        return m

    def value(self, v) -> int:
        if isinstance(v, dict):
            return self.struct(v)
        if isinstance(v, (list, tuple)) and v and isinstance(v[0], dict):
            return self.struct_array(v)
        return self.array(v)

    def call(self, nlhs: int, args):
        plhs = (C.c_void_p * max(nlhs, 1))()
        prhs = (C.c_void_p * max(len(args), 1))(*args)
        rc = self.lib.stub_call_mex(nlhs, plhs, len(args), prhs)
        return rc, [plhs[i] for i in range(max(nlhs, 1))], (self.lib.stub_last_error() or b"").decode()

    def read_struct_array(self, m, fields):
        n = self.lib.mxGetNumberOfElements(m)
        if not self.lib.mxIsStruct(m):
            return []
        return [{f: self.lib.mxGetScalar(self.lib.mxGetField(m, i, f.encode())) for f in fields} for i in range(n)]


def _rsp():
    import rsp_b200 as rsp
    return rsp


def test_gateways_link_against_librsp_and_reject_bad_calls():
    """Link + load + the reference-style argument errors (raised before any device call, so no GPU is needed)."""
    rsp = _rsp()
    for gw in ("fun_process_single_frame_mex", "process_stage2_mtd_mex"):
        s = MexStub(gw)
        assert hasattr(s.lib, "mexFunction")
        rc, _, err = s.call(1, [])
        assert rc == 1 and err.startswith("rsp:nargin"), err
    s = MexStub("fun_process_single_frame_mex")
    config, cfar_params, cluster_params = rsp.named_config("cfg1")
    pd = rsp.build_precomputed_data(config)
    broken = dict(config)
    broken.pop("Sig_Config")
    args = [s.struct_array([]), s.struct(broken), s.struct(cfar_params), s.struct(cluster_params), s.struct(pd)]
    rc, _, err = s.call(1, args)
    assert rc == 1 and err.startswith("rsp:missingField") and "Sig_Config" in err, err
    rc, _, err = s.call(2, args)
    assert rc == 1 and err.startswith("rsp:nargout"), err


def _noise_tools():
    spec = importlib.util.spec_from_file_location("make_inputs", os.path.join(ROOT, "tools", "ref_golden", "make_inputs.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["cfg1", "native"])
def test_mex_gateway_runs_and_equals_the_ctypes_path(name):
    """fun_process_single_frame(targets, config, cfar_params, cluster_params, precomputed_data, frame_idx) through the MEX
    gateway: echo synthesis in the gateway's C++ loops, randn blocks in the reference's order (per channel I then Q,
    column-major [P, N]) from the stub's stream, rsp_process_frame on the MATLAB-ordered complex-double cube -- against
    the NumPy synthesis + the same noise through RadarChain.process_cpi + rsp.cluster."""
    rsp = _rsp()
    config, cfar_params, cluster_params = rsp.named_config(name)
    pd = rsp.build_precomputed_data(config)
    sc = config.Sig_Config
    P, N, Cn = int(sc.prtNum), int(sc.point_PRT), int(sc.channel_num)
    v_max = sc.wavelength / (2 * sc.prt)
    targets = [dict(Range=3000.0, Velocity=0.10 * v_max, ElevationAngle=10.0, SNR_dB=10.0),
               dict(Range=900.0, Velocity=-0.08 * v_max, ElevationAngle=-5.0, SNR_dB=20.0),
               dict(Range=8000.0 if name != "cfg1" else 6000.0, Velocity=0.05 * v_max, ElevationAngle=15.0, SNR_dB=12.0)]
    blocks, noise = _noise_tools().noise_cube(P, Cn, N, seed=3)           # blocks: randn call order; noise: [p, c, n]

    s = MexStub("fun_process_single_frame_mex")
    stream = np.ascontiguousarray(blocks.reshape(-1))
    s.lib.stub_set_random_stream(stream.ctypes.data_as(C.POINTER(C.c_double)), stream.size)
    args = [s.struct_array(targets), s.struct(config), s.struct(cfar_params), s.struct(cluster_params), s.struct(pd), s.array(7.0)]
    rc, out, err = s.call(1, args)
    assert rc == 0, err
    assert s.lib.stub_random_consumed() == stream.size                     # 2 C randn(P, N) calls, nothing else
    assert b"frame 7 done" in s.lib.stub_printed()
    got = s.read_struct_array(out[0], ["Range", "Velocity", "Angle", "Power"])

    raw = rsp.synthesize_echo(targets, config, pd) + noise                 # complex128 [p, c, n]
    with rsp.RadarChain(config, cfar_params, pd) as chain:
        dets = chain.process_cpi(np.ascontiguousarray(np.transpose(raw, (1, 2, 0))), layout="matlab")   # memory order of MATLAB [P, N, C]
        _, final = rsp.cluster(dets, cluster_params)
    want = [dict(Range=float(t["range"]), Velocity=float(t["velocity"]), Angle=float(t["angle"]), Power=float(t["power"])) for t in final]
    assert len(got) == len(want) >= 3
    for g, w in zip(got, want):
        assert abs(g["Range"] - w["Range"]) <= 1e-3 and abs(g["Velocity"] - w["Velocity"]) <= 1e-4
        assert abs(g["Angle"] - w["Angle"]) <= 1e-4 and abs(g["Power"] - w["Power"]) <= 1e-5 * w["Power"]
    # an empty scene without noise returns the reference's 0 x 0 double, not a struct
    zero = np.zeros(stream.size)
    s.lib.stub_set_random_stream(zero.ctypes.data_as(C.POINTER(C.c_double)), zero.size)
    rc, out, err = s.call(1, [s.struct_array([])] + args[1:5])
    assert rc == 0 and not s.lib.mxIsStruct(out[0]) and s.lib.mxGetNumberOfElements(out[0]) == 0, err
    s.lib.stub_run_at_exit()


@pytest.mark.gpu
def test_stage2_mex_gateway_runs_and_equals_the_ctypes_path():
    """[MTD_results, PC_results] = process_stage2_mtd(iq_data, angle, config) through its MEX gateway (complex double
    [P, 3404, B] in, two complex double cubes out, process_stage2_mtd.m:1-52) against rsp.process_stage2_mtd."""
    rsp = _rsp()
    P, B, gates = 16, 2, [228, 723, 2453]
    config = rsp.Struct(Sig_Config=rsp.Struct(fs=25e6, prtNum=P, tao=[0.16e-6, 8e-6, 28e-6], B=20e6, point_prt=[sum(gates)] + gates),
                        mtd=rsp.Struct(beam_num=B), cfar=rsp.Struct(MTD_0v_num=2))
    rng = np.random.default_rng(5)
    iq = (rng.standard_normal((P, sum(gates), B)) + 1j * rng.standard_normal((P, sum(gates), B))) * np.sqrt(0.5)
    s = MexStub("process_stage2_mtd_mex")
    rc, out, err = s.call(2, [s.array(iq), s.array(0.0), s.struct(config)])
    assert rc == 0, err
    want_mtd, want_pc = rsp.process_stage2_mtd(iq, None, config)
    for m, want in zip(out, (want_mtd, want_pc)):
        assert s.lib.mxIsComplex(m) and s.lib.mxGetNumberOfDimensions(m) == 3
        assert [s.lib.mxGetDimensions(m)[i] for i in range(3)] == [P, sum(gates), B]
        flat = np.ctypeslib.as_array(s.lib.mxGetDoubles(m), shape=(2 * iq.size,)).view(np.complex128)
        got = flat.reshape((P, sum(gates), B), order="F")
        assert np.array_equal(got, want)                                   # the same library calls on the same bytes
    rc, _, err = s.call(2, [s.array(iq.real), s.array(0.0), s.struct(config)])
    assert rc == 1 and err.startswith("rsp:type"), err
    s.lib.stub_run_at_exit()
