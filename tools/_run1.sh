timeout 600 python -m pytest tests/test_gpu_hub.py -q 2>&1 | tail -15 > gpurun_out/r02j_gpu_tests.txt
( time timeout 900 python bench.py > gpurun_out/r02j_bench.json 2> gpurun_out/r02j_bench.err ) 2> gpurun_out/r02j_bench.time
( time timeout 600 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/r02j_bench_ref.json 2> gpurun_out/r02j_bench_ref.err ) 2> gpurun_out/r02j_bench_ref.time
tail -3 gpurun_out/r02j_gpu_tests.txt; cat gpurun_out/r02j_bench.time gpurun_out/r02j_bench_ref.time; tail -3 gpurun_out/r02j_bench.err
