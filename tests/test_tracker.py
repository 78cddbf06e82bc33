"""Host-side driver logic around the hot path: kinematics, scan and inter-frame track association
(main_simulate_echoes_with_array_v8_3.m:100-117, 203-228, 255-352).  CPU only: the per-frame kernel is
replaced by a stub so the loop and the association are checked against hand-computed answers."""
import math

import numpy as np
import pytest

import rsp_b200 as rsp


def test_kinematics_constant_height_and_ground_speed():
    tr = rsp.init_tracks([dict(Range=3000.0, Velocity=20.0, ElevationAngle=10.0, SNR_dB=10.0)])
    H, Vg = 3000 * math.sin(math.radians(10)), 20 / math.cos(math.radians(10))
    assert tr[0]["const_H"] == pytest.approx(H) and tr[0]["const_V_ground"] == pytest.approx(Vg)
    T = 332 * 232.76e-6
    cur = rsp.evolve(tr, T)[0]
    Rg = 3000 * math.cos(math.radians(10)) - Vg * T
    assert cur["Range"] == pytest.approx(math.hypot(Rg, H)) and cur["SNR_dB"] == 10.0
    assert cur["ElevationAngle"] == pytest.approx(math.degrees(math.asin(H / cur["Range"])))
    assert cur["Velocity"] == pytest.approx(Vg * math.cos(math.radians(cur["ElevationAngle"])))
    assert tr[0]["current_R_ground"] == pytest.approx(Rg)            # state advanced in place (v8_3:213)


def _det(R, V, A, P, f, az):
    return dict(Range=R, Velocity=V, Angle=A, Power=P, iFrame=f, iAntAngle=az)


def test_inter_frame_association_gates_and_merge():
    _, _, cluster_params = rsp.default_config()
    cfg = rsp.default_scan_and_track_config(cluster_params)
    log = [_det(3000, 20.0, 10.0, 5.0, 1, 2.8), _det(2998, 20.1, 10.1, 9.0, 2, 5.6), _det(2996, 20.2, 10.0, 7.0, 3, 8.3),
           _det(10000, 25.0, 10.0, 4.0, 1, 2.8), _det(9998, 25.0, 10.0, 6.0, 6, 16.7),        # 5 frames apart: new track
           _det(2990, 20.0, 10.0, 3.0, 5, 13.9)]                                               # chains to frame 3 via the gap of 2
    tracks = rsp.inter_frame_cluster(log, cfg)
    assert len(tracks) == 3
    t0 = tracks[0]
    assert (t0["NumPoints"], t0["FirstFrame"], t0["LastFrame"]) == (4, 1, 5)
    assert (t0["Range"], t0["Power"]) == (2998, 9.0)                                           # winner takes R/V/Angle
    w = np.array([5.0, 9.0, 7.0, 3.0])
    assert t0["Azimuth"] == pytest.approx(np.sum(np.array([2.8, 5.6, 8.3, 13.9]) * w) / w.sum())
    assert tracks[1]["NumPoints"] == 1 and tracks[2]["FirstFrame"] == 6
    cfg.inter_frame_cluster.enable = False
    assert len(rsp.inter_frame_cluster(log, cfg)) == len(log)
    assert rsp.inter_frame_cluster([], cfg) == []


def test_frame_loop_tags_frames_and_azimuth():
    config, cfar_params, cluster_params = rsp.default_config()
    seen = []

    def stub(targets, config, cfar_params, cluster_params, pd, frame_idx, rng=None):
        seen.append((frame_idx, targets[0]["Range"]))
        return [] if frame_idx == 2 else [dict(Range=targets[0]["Range"], Velocity=targets[0]["Velocity"],
                                               Angle=targets[0]["ElevationAngle"], Power=1.0 + frame_idx)]
    log, tracks = rsp.run_multiframe_simulation([dict(Range=3000.0, Velocity=20.0, ElevationAngle=10.0, SNR_dB=10.0)],
                                                config, cfar_params, cluster_params, None, total_frames=4, process_frame=stub)
    assert [f for f, _ in seen] == [1, 2, 3, 4] and seen[0][1] > seen[1][1] > seen[2][1]        # closing target
    assert [d["iFrame"] for d in log] == [1, 3, 4]
    dpf = 6 * 6.0 * 332 * 232.76e-6                                                              # v8_3:94-95
    assert log[0]["iAntAngle"] == pytest.approx(dpf) and log[1]["iAntAngle"] == pytest.approx(3 * dpf)
    assert len(tracks) == 1 and tracks[0]["NumPoints"] == 3 and tracks[0]["Power"] == 5.0


def test_tracker_against_the_oracle_statement():
    """f-4: the product's FIFO search (tracker.inter_frame_cluster, v8_3:272-336) against the oracle's independent
    statement (connected components of the 5-D gate graph) on random logs with chains, ties and isolated points; the
    kinematics (init_tracks / evolve, v8_3:103-117, :210-224) against the oracle's closed form."""
    from conftest import oracle as o
    rng = np.random.default_rng(12)
    cfg = rsp.default_scan_and_track_config(dict(max_range_sep=30.0, max_vel_sep=0.4, max_angle_sep=5.0))
    ifc = cfg["inter_frame_cluster"]
    for trial in range(6):
        log = []
        for k in range(int(rng.integers(3, 9))):                 # a few true tracks: slow drift over frames, some frames missing
            R0, V0, El0 = rng.uniform(2000, 9000), rng.uniform(-20, 20), rng.uniform(0, 40)
            for fr in range(1, 13):
                if rng.random() < 0.25:
                    continue
                log.append(dict(Range=R0 - 4.0 * fr + rng.normal(0, 3), Velocity=V0 + rng.normal(0, 0.05), Angle=El0 + rng.normal(0, 0.3),
                                Power=float(rng.uniform(1, 100)), iFrame=fr, iAntAngle=2.8 * fr))
        for _ in range(10):                                        # clutter
            log.append(dict(Range=rng.uniform(500, 12000), Velocity=rng.uniform(-30, 30), Angle=rng.uniform(-10, 60),
                            Power=float(rng.uniform(1, 100)), iFrame=int(rng.integers(1, 13)), iAntAngle=float(rng.uniform(0, 40))))
        log.append(dict(log[0]))                                   # an exact duplicate: a tie in power, first maximum wins
        rng.shuffle(log)
        got = rsp.inter_frame_cluster(log, cfg)
        want = o.inter_frame_tracks(log, ifc["Gate_R"], ifc["Gate_V"], ifc["Gate_Az"], ifc["Gate_El"], ifc["Max_Frame_Gap"])
        assert len(got) == len(want)
        for a, b in zip(got, want):
            for key in ("Range", "Velocity", "Angle", "Power"):
                assert a[key] == b[key], (trial, key)
            assert abs(a["Azimuth"] - b["Azimuth"]) <= 1e-12 * max(1.0, abs(b["Azimuth"]))
            assert (a["FirstFrame"], a["LastFrame"], a["NumPoints"]) == (b["FirstFrame"], b["LastFrame"], b["NumPoints"])
    T_frame = 332 * 232.76e-6
    tg = [dict(Range=3000.0, Velocity=20.0, ElevationAngle=10.0, SNR_dB=10.0), dict(Range=10000.0, Velocity=25.0, ElevationAngle=10.0, SNR_dB=15.0)]
    tracks = rsp.init_tracks(tg)
    for n in range(1, 41):
        cur = rsp.evolve(tracks, T_frame)
        for t, c in zip(tg, cur):
            R, El, Vr = o.track_state_after(o.Target(t["Range"], t["Velocity"], t["ElevationAngle"], t["SNR_dB"]), n, T_frame)
            assert abs(c["Range"] - R) <= 1e-9 * R and abs(c["ElevationAngle"] - El) <= 1e-10 and abs(c["Velocity"] - Vr) <= 1e-10
