#!/bin/bash
# One GPU-box call that refreshes every measured artefact of the round: GPU tests, the ncu summaries of all
# kernel families (tools/evidence_r02.sh), the roofline inputs of bench.py from the capture at its own
# workload, the per-kernel rooflines and a full bench line.  Everything lands in gpurun_out/.
cd "$(dirname "$0")/.." || exit 1
timeout 500 python -m pytest tests -m gpu -q 2>&1 | tail -6 > gpurun_out/pytest_final.log; tail -2 gpurun_out/pytest_final.log
bash tools/evidence_r02.sh
cp gpurun_out/ncu_mc_r02_cfg2.json gpurun_out/ncu_r02_mc_plain.json profiles/
python tools/make_roofline_inputs.py profiles/ncu_mc_r02_cfg2.json 4 6,2,0 profiles/ncu_r02_mc_plain.json > /dev/null
cp profiles/roofline_inputs_r02.json gpurun_out/roofline_inputs_r02.json
timeout 500 python tools/kernel_bench.py > gpurun_out/kernel_bench_final.jsonl 2> gpurun_out/kernel_bench_final.err
timeout 400 python bench.py > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err; tail -2 gpurun_out/bench_final.err | cut -c1-200
du -sh gpurun_out
