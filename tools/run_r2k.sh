V=radar-signal-simulation-and-target-detection_b200/lib/variants
for cfg in cfg2 cfg3; do
for lib in "" $V/librsp_cfaru5.so $V/librsp_cfaru8.so; do
echo "lib=$lib"; CONFIG=$cfg STEPS=$([ $cfg = cfg2 ] && echo 4 || echo 2) tools/ab_bench.sh "RSP_LIBRARY=$lib"
done; done 2>&1 | tee gpurun_out/r2k_cfar5_u.txt
