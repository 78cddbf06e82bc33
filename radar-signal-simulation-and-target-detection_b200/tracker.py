"""The multi-frame driver around the hot path (main_simulate_echoes_with_array_v8_3.m): target
kinematics at constant height and ground speed (v8_3:100-117, 209-228), the azimuth scan (v8_3:93-95,
207), the frame loop calling fun_process_single_frame (v8_3:200-248) and the inter-frame 5-D BFS track
association (v8_3:255-352).  Host-side logic over a handful of detections per frame; the per-frame work
is the device chain.
"""
from __future__ import annotations

import math
from typing import Callable, List, Optional, Sequence

import numpy as np

from .frame import fun_process_single_frame, fun_process_frames, _field
from .precompute import Struct


def default_scan_and_track_config(cluster_params) -> Struct:
    """config.scan / config.inter_frame_cluster literals of v8_3:24-25, 57-65."""
    K = 1
    return Struct(scan=Struct(rpm=6, start_azimuth=0.0),
                  inter_frame_cluster=Struct(enable=True, K=K, Gate_R=_field(cluster_params, "max_range_sep") * K,
                                             Gate_V=_field(cluster_params, "max_vel_sep") * K,
                                             Gate_El=_field(cluster_params, "max_angle_sep") * K, Gate_Az=10.0, Max_Frame_Gap=3))


def init_tracks(targets: Sequence[dict]) -> List[dict]:
    """v8_3:103-117: constant height H and ground speed from the initial R, El, V_rad."""
    out = []
    for t in targets:
        R0, El0, V0 = float(t["Range"]), float(t["ElevationAngle"]), float(t["Velocity"])
        d = dict(t)
        d["const_H"] = R0 * math.sin(math.radians(El0))
        d["const_V_ground"] = V0 / math.cos(math.radians(El0))
        d["current_R_ground"] = R0 * math.cos(math.radians(El0))
        out.append(d)
    return out


def evolve(tracks: List[dict], T_frame: float) -> List[dict]:
    """v8_3:210-228: one frame of motion; returns the targets to process (Range, ElevationAngle, Velocity
    updated, SNR_dB unchanged) and advances ``current_R_ground`` in place."""
    out = []
    for t in tracks:
        Rg = t["current_R_ground"] - t["const_V_ground"] * T_frame
        t["current_R_ground"] = Rg
        R = math.sqrt(Rg * Rg + t["const_H"] ** 2)
        El = math.degrees(math.asin(t["const_H"] / R))
        d = dict(t)
        d["Range"], d["ElevationAngle"], d["Velocity"] = R, El, t["const_V_ground"] * math.cos(math.radians(El))
        out.append(d)
    return out


def inter_frame_cluster(detection_log: Sequence[dict], cfg) -> List[dict]:
    """v8_3:255-352.  detection_log entries carry Range, Velocity, Angle, Power, iFrame, iAntAngle.
    BFS (FIFO, index order) under the 5-D gate; merge = strongest detection's R/V/Angle, power-weighted
    azimuth, max power, first/last frame, point count."""
    ifc = _field(cfg, "inter_frame_cluster")
    n = len(detection_log)
    if n == 0:
        return []
    if not _field(ifc, "enable"):                               # v8_3:337-352
        return [dict(Range=d["Range"], Velocity=d["Velocity"], Angle=d["Angle"], Azimuth=d["iAntAngle"], Power=d["Power"],
                     FirstFrame=d["iFrame"], LastFrame=d["iFrame"], NumPoints=1) for d in detection_log]
    gR, gV, gAz, gEl, gap = (_field(ifc, k) for k in ("Gate_R", "Gate_V", "Gate_Az", "Gate_El", "Max_Frame_Gap"))
    ids = [0] * n
    cur = 0
    for i in range(n):
        if ids[i]:
            continue
        cur += 1
        queue = [i]
        while queue:
            k = queue.pop(0)
            if ids[k]:
                continue
            ids[k] = cur
            a = detection_log[k]
            for j in range(n):
                if ids[j]:
                    continue
                b = detection_log[j]
                if (abs(a["Range"] - b["Range"]) <= gR and abs(a["Velocity"] - b["Velocity"]) <= gV and
                        abs(a["iAntAngle"] - b["iAntAngle"]) <= gAz and abs(a["Angle"] - b["Angle"]) <= gEl and
                        abs(a["iFrame"] - b["iFrame"]) <= gap):
                    queue.append(j)
    tracks = []
    for c in range(1, cur + 1):
        members = [detection_log[i] for i in range(n) if ids[i] == c]
        powers = np.array([m["Power"] for m in members], dtype=np.float64)
        w = members[int(np.argmax(powers))]                      # first maximum, like MATLAB max
        frames = [m["iFrame"] for m in members]
        tracks.append(dict(Range=w["Range"], Velocity=w["Velocity"], Angle=w["Angle"],
                           Azimuth=float(np.sum(np.array([m["iAntAngle"] for m in members]) * powers) / powers.sum()),
                           Power=float(powers.max()), FirstFrame=min(frames), LastFrame=max(frames), NumPoints=len(frames)))
    return tracks


def run_multiframe_simulation(targets, config, cfar_params, cluster_params, precomputed_data, total_frames: int = 50,
                              scan_cfg: Optional[Struct] = None, rng: Optional[np.random.Generator] = None,
                              process_frame: Callable = fun_process_single_frame, **kw):
    """The body of main_simulate_echoes_with_array_v8_3.m:192-352 -> (cumulative_final_log, final_tracks_log)."""
    sc = _field(config, "Sig_Config")
    scan_cfg = scan_cfg or default_scan_and_track_config(cluster_params)
    T_frame = _field(sc, "prtNum") * _field(sc, "prt")                      # v8_3:93
    deg_per_frame = _field(scan_cfg.scan, "rpm") * (360.0 / 60.0) * T_frame   # v8_3:94-95
    azimuth = float(_field(scan_cfg.scan, "start_azimuth"))
    tracks = init_tracks(targets)
    log: List[dict] = []
    rng = rng if rng is not None else np.random.default_rng()
    # the kinematics do not depend on the detections, so every frame's truth is known up front (v8_3:207-228)
    azimuths, scenes = [], []
    for frame_idx in range(1, total_frames + 1):
        azimuth = (azimuth + deg_per_frame) % 360.0                            # v8_3:207
        azimuths.append(azimuth)
        scenes.append(evolve(tracks, T_frame))                                 # v8_3:210-228
    if process_frame is fun_process_single_frame and not kw.get("host_synthesis"):
        kw.pop("host_synthesis", None)
        per_frame = fun_process_frames(scenes, config, cfar_params, cluster_params, precomputed_data, 1, rng=rng, **kw)
    else:
        per_frame = [process_frame(sc_, config, cfar_params, cluster_params, precomputed_data, i + 1, rng=rng, **kw)
                     for i, sc_ in enumerate(scenes)]
    for i, final_targets in enumerate(per_frame):                              # v8_3:236-246
        for t in final_targets:
            d = dict(t)
            d["iFrame"], d["iAntAngle"] = i + 1, azimuths[i]
            log.append(d)
    return log, inter_frame_cluster(log, scan_cfg)
