timeout 900 python bench.py > gpurun_out/r02ao_bench.json 2> gpurun_out/r02ao_bench.err
tail -c 400 gpurun_out/r02ao_bench.err
python -c "
import json
d=json.load(open('gpurun_out/r02ao_bench.json'))
print({k:d[k] for k in ('value','ms_per_step','gpu_launches')}, d['e2e']['ms_per_step'])
print(d['scoring'])
print(d['baseline_models']['gat_l3_h4'], d['baseline_models']['graphsage_l3_mean'], d['baseline_models']['graph_transformer_ffn_l3_h4'])
print(d['laplacian_pe_device'])
"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/r02ao_launches.csv python bench.py --steps 2 --warmup 1 --step-only > gpurun_out/r02ao_ncu.log 2>&1
tail -n 2 gpurun_out/r02ao_ncu.log | cut -c1-300
