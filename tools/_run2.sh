TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
timeout 300 python -m pytest tests/test_gpu_peer.py -q -k two_gpus 2>&1 | tail -15 > gpurun_out/r02e_two_gpu_test.txt
timeout 400 $TR bench.py --gpus 2 --steps 20 --warmup 5 --step-only > gpurun_out/r02e_bench_2gpu_peer.json 2> gpurun_out/r02e_bench_2gpu_peer.err
timeout 400 $TR bench.py --gpus 2 --steps 20 --warmup 5 --step-only --exchange nccl > gpurun_out/r02e_bench_2gpu_nccl.json 2> gpurun_out/r02e_bench_2gpu_nccl.err
timeout 400 $TR bench.py --gpus 2 --steps 40 --warmup 5 --step-only > gpurun_out/r02e_bench_2gpu_peer40.json 2> gpurun_out/r02e_bench_2gpu_peer40.err
