N=${1:-8}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
timeout 300 $TR tools/check_dp.py peer > gpurun_out/r02o_dp${N}_check_peer.txt 2>&1
timeout 300 $TR tools/dp_timeline.py peer rr > gpurun_out/r02o_timeline_peer${N}.json 2> gpurun_out/r02o_timeline_peer${N}.err
timeout 400 $TR bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/r02o_bench_${N}gpu.json 2> gpurun_out/r02o_bench_${N}gpu.err
timeout 400 $TR bench.py --gpus $N --steps 20 --warmup 5 --step-only --exchange nccl > gpurun_out/r02o_bench_${N}gpu_nccl.json 2> gpurun_out/r02o_bench_${N}gpu_nccl.err
timeout 600 $TR bench.py --gpus $N --steps 20 --warmup 5 --workload scaled > gpurun_out/r02o_scaled_${N}gpu.json 2> gpurun_out/r02o_scaled_${N}gpu.err
tail -n 4 gpurun_out/r02o_dp${N}_check_peer.txt; cat gpurun_out/r02o_timeline_peer${N}.json | cut -c1-900
