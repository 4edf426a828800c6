timeout 300 python -m pytest tests/test_gpu_score_tc.py -x -q > gpurun_out/r02ai_score_tests.txt 2>&1
tail -n 3 gpurun_out/r02ai_score_tests.txt
ETPGT_SCORE_EPI=2 timeout 300 python -m pytest tests/test_gpu_score_tc.py -x -q > gpurun_out/r02ai_score_tests_epi2.txt 2>&1
tail -n 2 gpurun_out/r02ai_score_tests_epi2.txt
timeout 600 python -m pytest tests/test_gpu_laplacian.py -q > gpurun_out/r02ai_lap_tests.txt 2>&1
tail -n 15 gpurun_out/r02ai_lap_tests.txt
timeout 120 python tools/prof_scoring.py > gpurun_out/r02ai_score.txt 2>&1 || exit 1
timeout 120 python tools/prof_scoring.py 23861 popular >> gpurun_out/r02ai_score.txt 2>&1
timeout 120 python tools/prof_scoring.py 100000 >> gpurun_out/r02ai_score.txt 2>&1
ETPGT_SCORE_STATS=1 timeout 120 python tools/prof_scoring.py 2>&1 | tail -n 3 >> gpurun_out/r02ai_score.txt
ETPGT_SCORE_EPI=2 ETPGT_SCORE_STATS=1 timeout 120 python tools/prof_scoring.py 2>&1 | tail -n 3 >> gpurun_out/r02ai_score.txt
cat gpurun_out/r02ai_score.txt
