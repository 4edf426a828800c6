timeout 280 python -m pytest tests/test_gpu_peer.py -x -q -k two_gpus > gpurun_out/r02aq_peer_tests_2gpu.txt 2>&1
tail -n 25 gpurun_out/r02aq_peer_tests_2gpu.txt
