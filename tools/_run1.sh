# One-GPU validation of the final state (under gpurun): the GPU parity suite, the smoke check, the bench line.
#   gpurun --timeout 2400 -- bash tools/_run1.sh [tag]
TAG=${1:-r02final}
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/${TAG}_gpu_tests.txt 2>&1
tail -n 4 gpurun_out/${TAG}_gpu_tests.txt
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -n 2
timeout 900 python bench.py > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
tail -c 300 gpurun_out/${TAG}_bench.err
