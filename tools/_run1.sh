timeout 300 python -m pytest tests/test_gpu_score_tc.py -x -q > gpurun_out/r02ag_score_tests.txt 2>&1
tail -n 3 gpurun_out/r02ag_score_tests.txt
timeout 600 python -m pytest tests/test_gpu_gemm_tc.py tests/test_gpu_models.py -q > gpurun_out/r02ag_ffn_tests.txt 2>&1
tail -n 15 gpurun_out/r02ag_ffn_tests.txt
timeout 120 python tools/prof_scoring.py > gpurun_out/r02ag_score.txt 2>&1 || exit 1
timeout 120 python tools/prof_scoring.py 23861 popular >> gpurun_out/r02ag_score.txt 2>&1
timeout 120 python tools/prof_scoring.py 100000 >> gpurun_out/r02ag_score.txt 2>&1
for pm in 3 4 5 6 8; do echo "pending $pm" >> gpurun_out/r02ag_score.txt; ETPGT_SCORE_PENDING=$pm ETPGT_SCORE_STATS=1 timeout 120 python tools/prof_scoring.py 2>&1 | tail -n 3 >> gpurun_out/r02ag_score.txt; done
cat gpurun_out/r02ag_score.txt
