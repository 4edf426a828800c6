# N-GPU validation (under gpurun --gpus N): the two-process test of the suite, the data-parallel parity check, the bench.
#   gpurun --gpus 2 --timeout 1200 -- bash tools/_run2.sh 2 [tag]
N=${1:-2}
TAG=${2:-r02final}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
timeout 400 python -m pytest tests/test_gpu_peer.py -x -q > gpurun_out/${TAG}_peer_tests_${N}gpu.txt 2>&1
tail -n 3 gpurun_out/${TAG}_peer_tests_${N}gpu.txt
timeout 300 $TR tools/check_dp.py peer > gpurun_out/${TAG}_dp${N}_check_peer.txt 2>&1
tail -n 4 gpurun_out/${TAG}_dp${N}_check_peer.txt
timeout 400 $TR bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/${TAG}_bench_${N}gpu.json 2> gpurun_out/${TAG}_bench_${N}gpu.err
tail -n 1 gpurun_out/${TAG}_bench_${N}gpu.json | cut -c1-400
