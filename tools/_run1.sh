timeout 600 python -m pytest tests/test_gpu_models.py tests/test_gpu_kernels.py -q > gpurun_out/r02an_tests.txt 2>&1
tail -n 3 gpurun_out/r02an_tests.txt
timeout 300 python tools/prof_baselines.py gat 32768 2>&1 | tail -n 1
