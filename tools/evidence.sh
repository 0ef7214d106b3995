# Evidence pass A (GPU box): bench lines, ncu launch list and a full capture of the Monte-Carlo
# kernel at the bench configuration.  Outputs go to gpurun_out/ (<= 64 MiB per call).
python bench.py --steps 20 --warmup 5 2>gpurun_out/bench.err > gpurun_out/bench_n1.json; cut -c1-300 gpurun_out/bench_n1.json
python bench.py --impl reference --steps 3 --warmup 1 2>/dev/null > gpurun_out/bench_ref.json; cut -c1-200 gpurun_out/bench_ref.json
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 1 --quick > /dev/null 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:mc_kernel -s 4 -c 1 -f -o gpurun_out/prof_mc_cfg2 python bench.py --steps 3 --warmup 3 --quick > /dev/null 2>&1
ls -la gpurun_out | tail -6
