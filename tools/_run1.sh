timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/r02at_gpu_tests.txt 2>&1
tail -n 4 gpurun_out/r02at_gpu_tests.txt
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -n 3
timeout 900 python bench.py > gpurun_out/r02at_bench.json 2> gpurun_out/r02at_bench.err
tail -c 300 gpurun_out/r02at_bench.err
python -c "
import json
d=json.load(open('gpurun_out/r02at_bench.json'))
print({k:d[k] for k in ('value','ms_per_step','gpu_launches')}, d['e2e']['ms_per_step'], d['scoring']['ms'], d['scoring']['frac'], d['roofline']['frac'])
print({k:v['ms_per_step'] for k,v in d['baseline_models'].items() if 'ms_per_step' in v}, d['laplacian_pe_device']['seconds'])
"
