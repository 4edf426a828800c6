timeout 600 python -m pytest tests/test_gpu_models.py tests/test_gpu_training.py -q > gpurun_out/r02am_tests.txt 2>&1
tail -n 3 gpurun_out/r02am_tests.txt
timeout 300 python tools/prof_baselines.py sage 32768 2>&1 | tail -n 1
for sp in 3 4 5 7 8; do echo "splits $sp"; ETPGT_SCORE_SPLITS=$sp ETPGT_SCORE_STATS=1 timeout 120 python tools/prof_scoring.py 2>&1 | tail -n 3; done
