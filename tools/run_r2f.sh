python -m pytest tests -m gpu -x -q 2>&1 | tail -3
CONFIG=cfg3 STEPS=2 tools/ab_bench.sh "RSP_PULSE_BLOCK=0" "RSP_PULSE_BLOCK=8" "RSP_PULSE_BLOCK=16" "RSP_PULSE_BLOCK=8 RSP_LANES=2" "RSP_PULSE_BLOCK=16 RSP_LANES=2" "RSP_PULSE_BLOCK=32 RSP_LANES=2" "RSP_PULSE_BLOCK=4" 2>&1 | tee gpurun_out/r2f_pulse_block_cfg3.txt
CONFIG=native STEPS=2 tools/ab_bench.sh "RSP_PULSE_BLOCK=0" "RSP_LANES=3" 2>&1 | tee -a gpurun_out/r2f_pulse_block_cfg3.txt
