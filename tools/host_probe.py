import sys, time, os
sys.path.insert(0, os.getcwd())
import numpy as np, torch, ctypes as C
import rsp_b200 as rsp
from rsp_b200 import _abi
config, cfar_params, cluster_params = rsp.named_config("cfg2")
pd = rsp.build_precomputed_data(config)
chain = rsp.RadarChain(config, cfar_params, pd, max_detections=32768)
chain.set_waveform(config, pd)
v_max = config.Sig_Config.wavelength / (2 * config.Sig_Config.prt)
rng = np.random.default_rng(1); dR = float(pd.deltaR); vb = (chain.P / 2 - 16) / chain.P * v_max
tl = [dict(Range=float(rng.uniform(700 * dR, (chain.G - 16) * dR)), Velocity=float(rng.uniform(-vb, vb)), ElevationAngle=float(rng.uniform(-15.0, 60.0)), SNR_dB=float(rng.uniform(-10.0, 20.0))) for _ in range(64)]
fin, dets = chain.process_targets(tl, cluster_params, 1.0, 1)
print("dets", len(dets), "final", len(fin))
lib = _abi.load()
cp = _abi.rsp_cluster_params(float(cluster_params.max_range_sep), float(cluster_params.max_vel_sep), float(cluster_params.max_angle_sep))
out = np.zeros(len(dets), rsp.frame.TARGET_DTYPE); nf = C.c_int32(0); n1 = C.c_int32(0)
d = np.ascontiguousarray(dets)
for rep in range(3):
    t0 = time.perf_counter()
    for _ in range(300):
        lib.rsp_cluster(C.c_void_p(d.ctypes.data), len(d), C.byref(cp), None, C.byref(n1), C.c_void_p(out.ctypes.data), C.byref(nf))
    print("rsp_cluster us", (time.perf_counter() - t0) / 300 * 1e6, nf.value)
sh = d.copy(); np.random.default_rng(0).shuffle(sh)
t0 = time.perf_counter()
for _ in range(300): rsp.sort_detections(sh)
print("python sort_detections us", (time.perf_counter() - t0) / 300 * 1e6)
slots = 64
for i in range(slots): chain.submit_targets(tl, i, 1.0, i)
chain.synchronize(); torch.cuda.synchronize()
t0 = time.perf_counter()
for i in range(slots): chain.fetch_targets(i, cluster_params)
print("fetch_targets us", (time.perf_counter() - t0) / slots * 1e6)
t0 = time.perf_counter()
for i in range(slots): chain._pack_targets(tl)
print("pack us", (time.perf_counter() - t0) / slots * 1e6)
