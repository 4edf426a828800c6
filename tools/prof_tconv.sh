#!/bin/bash
# ncu --set full of the fused TransformerConv kernels of ONE training step of bench.py (first 6 tconv_* launches:
# forward of layers 0 and 1, backward dst/src passes of layers 1 and 0), on one GPU.
#   tools/prof_tconv.sh <tag>      (under gpurun; writes gpurun_out/<tag>_tconv.ncu-rep, *_tconv_ncu_full.csv and
#                                   appends the DRAM bytes of layer 0 to profiles/tconv_traffic.json)
TAG=${1:-r02}
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:tconv_ -c 6 -f -o gpurun_out/${TAG}_tconv \
    python bench.py --steps 1 --warmup 1 --step-only > gpurun_out/${TAG}_tconv_ncu.log 2>&1
ncu -i gpurun_out/${TAG}_tconv.ncu-rep --page raw --csv > gpurun_out/${TAG}_tconv_raw.csv 2>/dev/null
python tools/tconv_traffic.py gpurun_out/${TAG}_tconv_raw.csv gpurun_out/${TAG}_tconv_ncu.log ${TAG}
