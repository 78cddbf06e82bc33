# final-build evidence of round 2 on one B200: bench lines, ncu launch list, ncu --set full of one CPI
python bench.py > gpurun_out/r2w_bench_cfg2.json 2> gpurun_out/r2w_bench_cfg2.err
tail -c 600 gpurun_out/r2w_bench_cfg2.json
for cfg in cfg1 cfg3 native; do python bench.py --config $cfg --steps 6 --no-cpu-baseline --no-extras > gpurun_out/r2w_bench_$cfg.json 2>/dev/null; done
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2w_bench_reference.json 2> gpurun_out/r2w_bench_reference.err
RSP_GRAPH=0 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2w_launches_cfg2.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/r2w_ncu_list.log 2>&1
ncu --set full --clock-control none --import-source on -f -o gpurun_out/r2w_full_cfg2 python tools/profile_chain.py --cpis 1 > gpurun_out/r2w_ncu_full.log 2>&1
ncu --set full --clock-control none --import-source on -f -o gpurun_out/r2w_full_cfg3 python tools/profile_chain.py --config cfg3 --cpis 1 --pool 1 > gpurun_out/r2w_ncu_full_cfg3.log 2>&1
ls -la gpurun_out/r2w_*
