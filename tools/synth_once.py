#!/usr/bin/env python
"""A few frames through the targets -> final_targets path (for ncu captures of the synthesis kernel)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import rsp_b200 as rsp
name = sys.argv[1] if len(sys.argv) > 1 else "cfg2"
config, cfar_params, cluster_params = rsp.named_config(name)
pd = rsp.build_precomputed_data(config)
chain = rsp.RadarChain(config, cfar_params, pd)
chain.set_waveform(config, pd)
v_max = config.Sig_Config.wavelength / (2 * config.Sig_Config.prt)
tl = [dict(Range=900.0, Velocity=0.15 * v_max, ElevationAngle=-5.0, SNR_dB=20.0),
      dict(Range=3000.0, Velocity=-0.10 * v_max, ElevationAngle=8.2, SNR_dB=10.0),
      dict(Range=8000.0, Velocity=0.05 * v_max, ElevationAngle=15.0, SNR_dB=10.0)]
for i in range(4):
    fin, dets = chain.process_targets(tl, cluster_params, 1.0, i)
print(len(fin), len(dets))
