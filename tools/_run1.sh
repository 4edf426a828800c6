timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -n 12 > gpurun_out/r02r_gpu_tests.txt
timeout 600 python bench.py --skip-cpu-baseline > gpurun_out/r02r_bench.json 2> gpurun_out/r02r_bench.err
tail -n 4 gpurun_out/r02r_gpu_tests.txt; tail -n 3 gpurun_out/r02r_bench.err
