timeout 600 python -m pytest tests/test_gpu_hub.py tests/test_gpu_optim.py tests/test_gpu_step.py -q 2>&1 | tail -5 > gpurun_out/r02i_gpu_tests.txt
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02i_small_launches.csv python tools/prof_small.py 1024 --ncu > gpurun_out/r02i_ncu.log 2>&1
tail -3 gpurun_out/r02i_gpu_tests.txt
