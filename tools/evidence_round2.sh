# final-build evidence of round 2 on one B200: bench lines, ncu launch list, ncu --set full of one CPI
python bench.py > gpurun_out/r2w_bench_cfg2.json 2> gpurun_out/r2w_bench_cfg2.err
tail -c 600 gpurun_out/r2w_bench_cfg2.json
for cfg in cfg1 cfg3 native; do python bench.py --config $cfg --steps 6 --no-cpu-baseline --no-extras > gpurun_out/r2w_bench_$cfg.json 2>/dev/null; done
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2w_bench_reference.json 2> gpurun_out/r2w_bench_reference.err
RSP_GRAPH=0 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2w_launches_cfg2.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/r2w_ncu_list.log 2>&1
ncu --set full --clock-control none --import-source on -f -o gpurun_out/r2w_full_cfg2 python tools/profile_chain.py --cpis 1 > gpurun_out/r2w_ncu_full.log 2>&1
ncu --set full --clock-control none --import-source on -f -o gpurun_out/r2w_full_cfg3 python tools/profile_chain.py --config cfg3 --cpis 1 --pool 1 > gpurun_out/r2w_ncu_full_cfg3.log 2>&1
ls -la gpurun_out/r2w_*
# frame paths, other shapes, multi-GPU (each on its own gpurun call in the round: see profiles/README.md r2u*/r2w*/r2y*)
python tools/frames_probe.py --targets 64 > gpurun_out/r2u3_frames_probe_k64.json 2>/dev/null
python tools/frames_probe.py > gpurun_out/r2u3_frames_probe_t3.json 2>/dev/null
python tools/mc_sweep.py --trials 476 > gpurun_out/r2y_mc_sweep_native_1gpu.json 2>/dev/null
ncu --set full --clock-control none --import-source on -f -o gpurun_out/r2w_full_native python tools/profile_chain.py --config native --cpis 1 --pool 1 > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:dbf_synth -c 2 -f -o gpurun_out/r2w_full_frames_k64 python tools/profile_frames.py > /dev/null 2>&1
# python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 8 --steps 10 --warmup 3 > gpurun_out/r2w_bench_cfg2_8gpu.json
