python -m pytest tests -m gpu -x -q 2>&1 | tail -3
for g in 0 1; do echo "RSP_GRAPH=$g"; RSP_GRAPH=$g python tools/enqueue_cost.py 2>&1 | tail -3; done | tee gpurun_out/r2e_enqueue_cost.txt
STEPS=4 tools/ab_bench.sh "RSP_GRAPH=0" "RSP_GRAPH=1" "RSP_LANES=4" "RSP_LANES=5" "RSP_LANES=6" "RSP_LANES=4 RSP_GRAPH=0" 2>&1 | tee gpurun_out/r2e_graph_lanes_ab.txt
CONFIG=cfg1 STEPS=4 tools/ab_bench.sh "RSP_GRAPH=0" "RSP_GRAPH=1" "RSP_LANES=4" "RSP_LANES=6" 2>&1 | tee -a gpurun_out/r2e_graph_lanes_ab.txt
CONFIG=cfg3 STEPS=2 tools/ab_bench.sh "RSP_GRAPH=1" "RSP_LANES=4" "RSP_LANES=2" 2>&1 | tee -a gpurun_out/r2e_graph_lanes_ab.txt
CONFIG=native STEPS=2 tools/ab_bench.sh "RSP_GRAPH=1" "RSP_LANES=4" 2>&1 | tee -a gpurun_out/r2e_graph_lanes_ab.txt
