timeout 600 python tools/exp_tconv.py > gpurun_out/r02q_exp_tconv.txt 2>&1
timeout 600 python -m pytest tests/test_gpu_hub.py -q 2>&1 | tail -n 3 > gpurun_out/r02q_hub_tests.txt
cat gpurun_out/r02q_exp_tconv.txt gpurun_out/r02q_hub_tests.txt
