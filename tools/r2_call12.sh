#!/bin/bash
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_sn_gemm_tc2|k_sn_edge_bwd|k_sn_edge_fwd|k_build_cells" -s 8 -c 8 \
    -o gpurun_out/c12_prof_schnet python tools/schnet_md_bench.py --config si --steps 1 > gpurun_out/c12.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_build_fast" -s 3 -c 1 \
    -o gpurun_out/c12_prof_build python bench.py --steps 30 --warmup 10 --no-cpu-baseline --no-e2e --no-c4 > /dev/null 2>&1
ls -la gpurun_out | grep c12
