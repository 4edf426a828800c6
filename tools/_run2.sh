N=${1:-2}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
timeout 400 python -m pytest tests/test_gpu_peer.py -x -q > gpurun_out/r02ap_peer_tests_${N}gpu.txt 2>&1
tail -n 3 gpurun_out/r02ap_peer_tests_${N}gpu.txt
timeout 300 $TR tools/check_dp.py peer > gpurun_out/r02ap_dp${N}_check_peer.txt 2>&1
tail -n 4 gpurun_out/r02ap_dp${N}_check_peer.txt
timeout 400 $TR bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/r02ap_bench_${N}gpu.json 2> gpurun_out/r02ap_bench_${N}gpu.err
tail -c 300 gpurun_out/r02ap_bench_${N}gpu.err
python -c "
import json,sys
d=json.load(open('gpurun_out/r02ap_bench_${N}gpu.json'))
print({k:d[k] for k in ('value','ms_per_step','n_gpus')}, d['e2e']['ms_per_step'], d.get('sharded_eval'), d.get('sustained'))
"
