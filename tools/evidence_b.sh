# Evidence pass B (GPU box): K4 captures and the config-4 wall time.
timeout 300 ncu --set full --clock-control none --import-source on -k regex:allan_stream_kernel -c 1 -f -o gpurun_out/prof_allan_r01_v13 python tools/kernel_bench.py K4 > /dev/null 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 16 --csv --log-file gpurun_out/launches_allan_v13.csv python tools/kernel_bench.py K4 > /dev/null 2>&1
timeout 600 python tools/config4.py 2>gpurun_out/config4.err | tee gpurun_out/config4_v2.json
tail -3 gpurun_out/config4.err
ls -la gpurun_out | tail -5
