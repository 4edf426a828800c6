TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
timeout 400 python -m pytest tests/test_gpu_peer.py -q -k two_gpus 2>&1 | tail -15 > gpurun_out/r02k_two_gpu_test.txt
timeout 400 $TR bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r02k_bench_2gpu.json 2> gpurun_out/r02k_bench_2gpu.err
timeout 600 $TR bench.py --gpus 2 --steps 20 --warmup 5 --workload scaled > gpurun_out/r02k_scaled_2gpu.json 2> gpurun_out/r02k_scaled_2gpu.err
timeout 600 python bench.py --steps 20 --warmup 5 --workload scaled > gpurun_out/r02k_scaled_1gpu.json 2> gpurun_out/r02k_scaled_1gpu.err
tail -3 gpurun_out/r02k_two_gpu_test.txt; tail -3 gpurun_out/r02k_bench_2gpu.err gpurun_out/r02k_scaled_2gpu.err gpurun_out/r02k_scaled_1gpu.err
