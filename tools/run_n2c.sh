# N=2: NCCL merge tests, then weak / strong bench lines with the exchange under the own rows (default) and after the row pass
timeout 300 python -m pytest tests/test_distributed.py -m gpu -x -q -k "slab_partition_merge" > gpurun_out/s22_tests.log 2>&1; tail -n 3 gpurun_out/s22_tests.log
for ov in 0 1; do
for sc in weak strong; do
XGPU_NO_OVERLAP=$ov timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 3 --scaling $sc > gpurun_out/s22_n2_${sc}_noov$ov.json 2> gpurun_out/s22_n2_${sc}_noov$ov.err
tail -n 1 gpurun_out/s22_n2_${sc}_noov$ov.json | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('$sc noov=$ov', d['n_gpus'], round(d['ms_per_step'],3), round(d['value']/1e6,1), 'e2e', round(d['e2e']['ms_per_step'],2), d.get('check',{}).get('ok')); print(d['config'].get('phases_ms_max_over_ranks'))"
tail -n 3 gpurun_out/s22_n2_${sc}_noov$ov.err
done
done
