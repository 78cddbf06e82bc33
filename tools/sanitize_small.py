"""Smallest end-to-end exercise of every kernel for compute-sanitizer (cfg1-like, tiny P)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import rsp_b200 as rsp

def run(cfgname, **shape):
    config, cfar_params, cluster_params = rsp.default_config(**shape) if shape else rsp.named_config(cfgname)
    pd = rsp.build_precomputed_data(config)
    ch = rsp.RadarChain(config, cfar_params, pd, max_detections=2048)
    ch.set_waveform(config, pd)
    v_max = config.Sig_Config.wavelength / (2 * config.Sig_Config.prt)
    tg = [dict(Range=900.0, Velocity=0.1 * v_max, ElevationAngle=-5.0, SNR_dB=20.0), dict(Range=8000.0, Velocity=0.05 * v_max, ElevationAngle=15.0, SNR_dB=10.0)]
    fin, dets = ch.process_targets(tg, cluster_params, 1.0, seed=1)
    cube = torch.empty((2, ch.P, ch.C, ch.N), dtype=torch.complex64, device="cuda")
    ch.synthesize(tg, 1.0, 2, out=cube[0]); ch.synthesize(tg, 1.0, 3, out=cube[1])
    rdm = torch.empty((2, ch.B, ch.G, ch.P), dtype=torch.complex64, device="cuda")
    ch.stream_enqueue(cube.data_ptr(), 2, 0, 0, 5, 0); ch.synchronize()                 # no ring: every lane its own map
    n = [len(ch.stream_fetch(i)) for i in range(5)]
    host = cube[0].cpu().numpy()
    a = ch.process_cpi(host); b = ch.process_cpi(np.ascontiguousarray(np.transpose(host, (1, 2, 0)).astype(np.complex128)), layout="matlab")
    assert np.array_equal(a, b)
    ch.submit_cpi(torch.from_numpy(host).pin_memory().numpy(), 1); ch.stream_fetch(1)
    print(cfgname, shape, "final", len(fin), "dets", len(dets), n)
    ch.close()

run("cfg1")                                                         # P=32, vectorised CFAR (5,4,2), pow2 MTD
run("odd", channel_num=16, beam_num=5, prtNum=20, point_PRT=4097)   # odd N, generic DFT MTD, generic CFAR window
for dbf in ("ffma", "mma2", "tma2", "tc"):
    os.environ["RSP_DBF"] = dbf
    run("cfg1-" + dbf, channel_num=16, beam_num=13, prtNum=32, point_PRT=4096)
os.environ.pop("RSP_DBF")
gates = [228, 723, 2453]
cfg2 = rsp.Struct(Sig_Config=rsp.Struct(fs=25e6, prtNum=16, tao=[0.16e-6, 8e-6, 28e-6], B=20e6, point_prt=[sum(gates)] + gates),
                  mtd=rsp.Struct(beam_num=2), cfar=rsp.Struct(MTD_0v_num=1))
iq = (np.random.default_rng(0).standard_normal((16, sum(gates), 2)) + 0j)
m, p = rsp.process_stage2_mtd(iq, None, cfg2)
print("stage2", m.shape, float(np.abs(m).max()))
