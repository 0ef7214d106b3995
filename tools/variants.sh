#!/bin/bash
# Build libb2ins variants with different register caps for the G = 1 (throughput) kernel and
# time 10^6 runs with each.  GPU box only.   bash tools/variants.sh > gpurun_out/variants.jsonl
cd "$(dirname "$0")/../gnss_ins_sim_b200/csrc" || exit 1
for mb in 3 4 5; do
  nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -shared -Xcompiler -fPIC \
       -DB2INS_G1_MINBLOCKS=$mb -o ../../tools/libb2ins_mb$mb.so b2ins_api.cu || exit 1
done
cd ../..
for mb in 3 4 5; do
  for rf in 1 0; do
    echo -n "{\"minblocks\": $mb, \"result\": "
    B2INS_LIB=$PWD/tools/libb2ins_mb$mb.so python tools/probe_mc.py 1000000 1 $rf 2 | tr -d '\n'
    echo "}"
  done
done
