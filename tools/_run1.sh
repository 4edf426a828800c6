timeout 600 python -m pytest tests/test_gpu_training.py -q > gpurun_out/r02ar_tests.txt 2>&1
tail -n 25 gpurun_out/r02ar_tests.txt
