timeout 300 python -m pytest tests/test_gpu_score_tc.py -x -q > gpurun_out/r02ac_score_tests.txt 2>&1
tail -n 3 gpurun_out/r02ac_score_tests.txt
timeout 120 python tools/prof_scoring.py > gpurun_out/r02ac_score.txt 2>&1 || exit 1
for pm in 4 6; do ETPGT_SCORE_PENDING=$pm ETPGT_SCORE_STATS=1 timeout 120 python tools/prof_scoring.py 2>&1 | tail -n 3 >> gpurun_out/r02ac_score.txt; done
ETPGT_SCORE_STATS=1 timeout 120 python tools/prof_scoring.py 2>&1 | tail -n 3 >> gpurun_out/r02ac_score.txt
cat gpurun_out/r02ac_score.txt
timeout 500 ncu --set full --clock-control none --import-source on -k regex:score_ -s 3 -c 2 -o gpurun_out/r02ac_score python tools/prof_scoring.py > gpurun_out/r02ac_ncu.log 2>&1
tail -n 2 gpurun_out/r02ac_ncu.log
