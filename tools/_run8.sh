TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511"
timeout 100 $TR bench.py --gpus 8 --steps 20 --warmup 5 > gpurun_out/r02au_bench_8gpu.json 2> gpurun_out/r02au_bench_8gpu.err
echo "rc=$?"
tail -n 1 gpurun_out/r02au_bench_8gpu.json | cut -c1-300
