#!/usr/bin/env python
"""Minimal driver for ncu: a few 64-target (or 3-target) frames through the fused S4 + S5 frame path.
    ncu --set full -k regex:dbf_synth -c 2 python tools/profile_frames.py [--targets 64]"""
import argparse, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import rsp_b200 as rsp

ap = argparse.ArgumentParser()
ap.add_argument("--config", default="cfg2")
ap.add_argument("--targets", type=int, default=64)
ap.add_argument("--frames", type=int, default=3)
a = ap.parse_args()
config, cfar_params, cluster_params = rsp.named_config(a.config)
pd = rsp.build_precomputed_data(config)
chain = rsp.RadarChain(config, cfar_params, pd, max_detections=32768)
chain.set_waveform(config, pd)
v_max = config.Sig_Config.wavelength / (2 * config.Sig_Config.prt)
rng = np.random.default_rng(1)
dR = float(pd.deltaR)
vb = (chain.P / 2 - 16) / chain.P * v_max
tl = [dict(Range=float(rng.uniform(700 * dR, (chain.G - 16) * dR)), Velocity=float(rng.uniform(-vb, vb)),
           ElevationAngle=float(rng.uniform(-15.0, 60.0)), SNR_dB=float(rng.uniform(-10.0, 20.0))) for _ in range(a.targets)]
res = chain.process_targets_batch([tl] * a.frames, cluster_params, 1.0, list(range(a.frames)), return_detections=False)
print("frames", len(res), "final targets", [len(f) for f, _ in res])
