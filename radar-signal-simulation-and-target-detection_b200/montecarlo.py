"""SNR-vs-angle-error Monte-Carlo sweep (main_plot_snr_vs_angle_error.m:15-24, 157-309) on the device chain.

For every SNR point the reference runs ``num_trials`` independent frames of one target (10000 m, 20 m/s,
10 deg) through S4..S11 inside a ``parfor`` and records ``final_targets(1).Angle - 10`` (mc:270-276), then
reports ``std(errors, 'omitnan')`` and the detection probability (mc:283-284) next to the theoretical line
``|k| * sqrt(2) / sqrt(SNR)`` (mc:306-309).  Here every trial is one ``rsp_process_targets`` call (echo
synthesis, noise, chain and clustering on the GPU; the complex-ratio monopulse of mc:454-461 when the chain
was created with ``monopulse_complex=True``); trials are independent, so ranks take disjoint trial blocks
and exchange only four sums per SNR point (one all_reduce).
"""
from __future__ import annotations

import math
from typing import Optional, Sequence

import numpy as np

from .frame import RadarChain, _field
from .stream import shard_range

DEFAULT_SNR_DB = tuple(range(-10, 31, 2))                     # mc:15
TRUE_TARGET = dict(Range=10000.0, Velocity=20.0, ElevationAngle=10.0, pair_idx=5)   # mc:21-27


def snr_vs_angle_error(config, cfar_params, cluster_params, precomputed_data, snr_db_vector: Sequence[float] = DEFAULT_SNR_DB,
                       num_trials: int = 100, true_target: Optional[dict] = None, seed: int = 0, device: int = 0,
                       chain: Optional[RadarChain] = None, rank: int = 0, world: int = 1, batch: int = 0) -> dict:
    """Returns {'snr_db', 'angle_error_std', 'detection_probability', 'theoretical_error_std', 'trials'}.
    With ``world > 1`` (torch.distributed initialised) every rank returns the reduced result."""
    tt = dict(TRUE_TARGET if true_target is None else true_target)
    own = chain is None
    if own:
        chain = RadarChain(config, cfar_params, precomputed_data, device=device, monopulse_complex=True)
    if not getattr(chain, "_has_waveform", False):
        chain.set_waveform(config, precomputed_data)
    lo, hi = shard_range(num_trials, rank, world)
    n_snr = len(snr_db_vector)
    sums = np.zeros((n_snr, 4), dtype=np.float64)               # sum err, sum err^2, detections, trials
    # every (SNR point, trial) frame is independent: one pipelined pass over all of them
    frames, point, seeds = [], [], []
    for i, snr in enumerate(snr_db_vector):
        tgt = [dict(Range=tt["Range"], Velocity=tt["Velocity"], ElevationAngle=tt["ElevationAngle"], SNR_dB=float(snr))]
        for trial in range(lo, hi):
            frames.append(tgt)
            point.append(i)
            seeds.append((seed << 32) ^ (i << 20) ^ trial)
    step = max(batch, 1) if batch else max(len(frames), 1)
    for f0 in range(0, len(frames), step):
        res = chain.process_targets_batch(frames[f0:f0 + step], cluster_params, 1.0, seeds[f0:f0 + step])
        for i, (final, _) in zip(point[f0:f0 + step], res):
            sums[i, 3] += 1
            if len(final):                                       # mc:270-276: the first final target
                err = float(final[0]["angle"]) - tt["ElevationAngle"]
                sums[i, 0] += err
                sums[i, 1] += err * err
                sums[i, 2] += 1
    if world > 1:
        import torch
        import torch.distributed as dist
        t = torch.from_numpy(sums)
        if dist.get_backend() == "nccl":
            t = t.cuda(device)
        dist.all_reduce(t)
        sums = t.cpu().numpy()
    if own:
        chain.close()
    n = sums[:, 2]
    mean = np.divide(sums[:, 0], n, out=np.full(n_snr, np.nan), where=n > 0)
    var = np.divide(sums[:, 1] - n * mean ** 2, n - 1, out=np.full(n_snr, np.nan), where=n > 1)     # std(...,'omitnan')
    k = abs(float(np.asarray(_field(precomputed_data, "k_slopes_LUT"))[tt.get("pair_idx", 5) - 1]))
    snr_lin = 10.0 ** (np.asarray(snr_db_vector, dtype=np.float64) / 10.0)
    return dict(snr_db=list(map(float, snr_db_vector)), angle_error_std=np.sqrt(np.maximum(var, 0.0)),
                angle_error_mean=mean, detection_probability=sums[:, 2] / np.maximum(sums[:, 3], 1),
                theoretical_error_std=k * math.sqrt(2.0) / np.sqrt(snr_lin), trials=sums[:, 3])
