#!/bin/bash
# ncu captures of every kernel family (round 2), summarised ON THE BOX (tools/ncu_summary.py) into
# gpurun_out/ncu_r02_<mode>.json -- the .ncu-rep files are deleted (together they exceed what travels
# back) -- plus the launch list of one bench step.  GPU box only (one GPU).
cd "$(dirname "$0")/.." || exit 1
for spec in "k1:imu_noise_kernel:6.5536e7" "k3:err_stage1_kernel:1e6" "k5:psd_fft_kernel:513" "k6:gps_noise_kernel:1e7" \
            "k7:ekf_kernel:4.096e6" "allan_gen:allan_gen_kernel:7.68e8" "allan_stream:allan_stream_kernel:1.92e8" \
            "mc_c3:mc_spec_kernel:1.25e7" "mc_plain:mc_kernel:1e9"; do
  IFS=: read -r mode kern units <<< "$spec"
  timeout 300 ncu --set full --import-source on --clock-control none -k regex:$kern -c 1 -f \
      -o gpurun_out/prof_r02_$mode python tools/evidence_r02.py $mode > gpurun_out/ncu_r02_$mode.log 2>&1
  python tools/ncu_summary.py gpurun_out/prof_r02_$mode.ncu-rep --units $units > gpurun_out/ncu_r02_$mode.json 2>> gpurun_out/ncu_r02_$mode.log
  rm -f gpurun_out/prof_r02_$mode.ncu-rep
  tail -1 gpurun_out/ncu_r02_$mode.log
done
# the dominant kernel at bench.py's own workload (config 2) -> the inputs of bench.py's roofline fields
timeout 300 ncu --set full --import-source on --clock-control none -k 'regex:mc_(spec|av)_kernel' -s 4 -c 1 -f \
    -o gpurun_out/prof_r02_mc_c2 python bench.py --steps 3 --warmup 3 --quick > gpurun_out/ncu_r02_mc_c2.log 2>&1
python tools/ncu_summary.py gpurun_out/prof_r02_mc_c2.ncu-rep --units 1e6 > gpurun_out/ncu_mc_r02_cfg2.json 2>> gpurun_out/ncu_r02_mc_c2.log
rm -f gpurun_out/prof_r02_mc_c2.ncu-rep
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r02.csv \
    python bench.py --steps 2 --warmup 3 --no-config4 --c3-runs 2000 --c5-runs 512 > gpurun_out/launches_r02_bench.log 2>&1
