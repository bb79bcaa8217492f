mkdir -p gpurun_out
for c in zern c5pol; do
ncu --set full --clock-control none --import-source on -k regex:trace -s 8 -c 1 -o gpurun_out/r2b_$c -f python scripts/bench_configs.py $c > gpurun_out/ncu_$c.log 2>&1
ncu -i gpurun_out/r2b_$c.ncu-rep --page raw --csv > gpurun_out/r2b_$c.raw.csv 2>/dev/null
done
ls -la gpurun_out/r2b_*.ncu-rep
