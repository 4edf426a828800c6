TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
for v in 148 74 37 296; do
ETPGT_DP_TABLE_CTAS=$v timeout 300 $TR tools/dp_timeline.py peer rr 2>/dev/null | cut -c1-700 > gpurun_out/r02u_tl_$v.json
done
ETPGT_DP_NO_OVERLAP=1 timeout 300 $TR tools/dp_timeline.py peer rr 2>/dev/null | cut -c1-700 > gpurun_out/r02u_tl_nooverlap.json
head -c 700 gpurun_out/r02u_tl_*.json
