# Evidence pass A (GPU box): bench lines, per-kernel rooflines, ncu launch list and full captures of
# the Monte-Carlo kernel.  Outputs go to gpurun_out/ (<= 64 MiB per call) and are copied to profiles/.
python bench.py --steps 20 --warmup 5 2>gpurun_out/bench_v5.err > gpurun_out/bench_v5_n1.json; cut -c1-300 gpurun_out/bench_v5_n1.json
python bench.py --impl reference --steps 3 --warmup 1 2>/dev/null > gpurun_out/bench_v5_ref.json; cut -c1-200 gpurun_out/bench_v5_ref.json
python tools/kernel_bench.py > gpurun_out/kernel_bench_v5.jsonl 2>gpurun_out/kb.err; wc -l gpurun_out/kernel_bench_v5.jsonl
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_v5.csv python bench.py --steps 2 --warmup 1 --quick > /dev/null 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:mc_kernel -s 4 -c 1 -f -o gpurun_out/prof_mc_r01_v5_cfg2 python bench.py --steps 3 --warmup 3 --quick > /dev/null 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:mc_kernel -s 1 -c 1 -f -o gpurun_out/prof_mc_r01_v5_1e6 python tools/probe_mc.py 1000000 1 1 1 > /dev/null 2>&1
ls -la gpurun_out | tail -9
