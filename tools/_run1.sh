timeout 600 python -m pytest tests/test_gpu_models.py tests/test_gpu_kernels.py tests/test_gpu_training.py -q > gpurun_out/r02al_tests.txt 2>&1
tail -n 12 gpurun_out/r02al_tests.txt
timeout 300 python tools/prof_baselines.py gat 32768 2>&1 | tail -n 1
timeout 300 python tools/prof_baselines.py sage 32768 2>&1 | tail -n 1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/r02al_gat_launches.csv python tools/prof_baselines.py gat 32768 > gpurun_out/r02al_gat.log 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/r02al_sage_launches.csv python tools/prof_baselines.py sage 32768 > gpurun_out/r02al_sage.log 2>&1
