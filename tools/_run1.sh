timeout 300 python bench.py --steps 2 --warmup 1 --step-only > gpurun_out/r02as_plain.log 2>&1 &&
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_bf16x3 -s 6 -c 6 -f -o gpurun_out/r02as_gemm python bench.py --steps 2 --warmup 1 --step-only > gpurun_out/r02as_ncu.log 2>&1
tail -n 2 gpurun_out/r02as_ncu.log | cut -c1-200
