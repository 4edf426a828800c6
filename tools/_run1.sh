timeout 900 python -m pytest tests/test_gpu_score_tc.py tests/test_gpu_step.py tests/test_gpu_peer.py tests/test_gpu_training.py -q 2>&1 | tail -n 15 > gpurun_out/r02t_gpu_tests.txt
tail -n 6 gpurun_out/r02t_gpu_tests.txt
