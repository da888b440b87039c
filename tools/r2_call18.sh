#!/bin/bash
# 2-GPU call: event timeline of slab steps (131 072 and 256 000 atoms per GPU) next to the single-GPU timeline; ncu of the list builder
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
rm -f gpurun_out/tl_*
for nc in 32 40; do
  MDG_TIMELINE=gpurun_out/tl_n1_${nc}_ timeout 300 python bench.py --steps 60 --warmup 40 --ncell $nc --no-c4 --no-e2e > gpurun_out/c18_n1_$nc.json 2> gpurun_out/c18_n1_$nc.err
  MDG_TIMELINE=gpurun_out/tl_n2_${nc}_ timeout 300 $TR --master-port 2953$((nc/10)) bench.py --gpus 2 --steps 60 --warmup 40 --ncell $nc --no-c4 --no-e2e --no-dist-parity > gpurun_out/c18_n2_$nc.json 2> gpurun_out/c18_n2_$nc.err
done
timeout 300 ncu --set full --clock-control none --import-source on -k k_build_fast --launch-skip 4 -c 1 -o gpurun_out/c18_prof_build -f python bench.py --steps 30 --warmup 5 --no-c4 --no-e2e > gpurun_out/c18_ncu.log 2>&1
ls -la gpurun_out/tl_* gpurun_out/c18_*
