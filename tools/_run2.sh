TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
timeout 300 python -m pytest tests/test_gpu_peer.py -q 2>&1 | tail -15 > gpurun_out/r02n_peer_tests.txt
timeout 300 $TR tools/dp_timeline.py peer rr > gpurun_out/r02n_timeline_peer2.json 2> gpurun_out/r02n_timeline_peer2.err
timeout 300 $TR tools/check_dp.py peer > gpurun_out/r02n_dp2_check_peer.txt 2>&1
tail -n 3 gpurun_out/r02n_peer_tests.txt; cat gpurun_out/r02n_timeline_*.json; tail -n 4 gpurun_out/r02n_dp2_check_peer.txt
