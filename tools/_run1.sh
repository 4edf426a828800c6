timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -25 > gpurun_out/r02f_gpu_tests.txt
timeout 900 python bench.py > gpurun_out/r02f_bench.json 2> gpurun_out/r02f_bench.err
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r02f_bench_ref.json 2> gpurun_out/r02f_bench_ref.err
timeout 600 bash tools/prof_tconv.sh r02f > gpurun_out/r02f_prof.log 2>&1
tail -5 gpurun_out/r02f_gpu_tests.txt; tail -3 gpurun_out/r02f_bench.err; tail -2 gpurun_out/r02f_prof.log
