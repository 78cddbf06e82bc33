"""B200-native per-frame phased-array chain (DBF -> PC -> MTD -> CFAR -> monopulse).

The directory name carries a hyphen (it is the graft package name), so import it through the
alias package ``rsp_b200`` at the repository root::

    import rsp_b200 as rsp
    config, cfar_params, cluster_params = rsp.default_config()
    pd = rsp.build_precomputed_data(config)
    final_targets = rsp.fun_process_single_frame(targets, config, cfar_params, cluster_params, pd, 1)
"""
from .precompute import (Struct, default_config, named_config, build_precomputed_data, read_dbf_csv,
                         dbf_tables, NAMED_SHAPES)
from .frame import (RadarChain, fun_process_single_frame, fun_process_frames, synthesize_echo, add_noise, cluster, sort_detections)
from .stage2 import Stage2Chain, process_stage2_mtd, reference_pulses, local_execute_cfar
from . import stream
from .montecarlo import snr_vs_angle_error
from .tracker import run_multiframe_simulation, inter_frame_cluster, init_tracks, evolve, default_scan_and_track_config
from ._abi import DETECTION_DTYPE, TARGET_DTYPE, RspError, LIB_PATH

__all__ = ["Struct", "default_config", "named_config", "build_precomputed_data", "read_dbf_csv", "dbf_tables",
           "NAMED_SHAPES", "RadarChain", "fun_process_single_frame", "fun_process_frames", "synthesize_echo", "add_noise", "cluster",
           "sort_detections", "Stage2Chain", "process_stage2_mtd", "reference_pulses", "local_execute_cfar", "stream", "snr_vs_angle_error", "run_multiframe_simulation", "inter_frame_cluster", "init_tracks", "evolve",
           "default_scan_and_track_config", "DETECTION_DTYPE", "TARGET_DTYPE", "RspError", "LIB_PATH"]
