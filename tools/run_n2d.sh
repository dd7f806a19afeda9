# N=2 weak bench line twice (default path: exchange after the row pass)
for i in 1 2; do
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 3 --scaling weak > gpurun_out/s30_n2_weak_$i.json 2> gpurun_out/s30_n2_weak_$i.err
tail -n 1 gpurun_out/s30_n2_weak_$i.json | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('weak', d['n_gpus'], round(d['ms_per_step'],3), round(d['value']/1e6,1), 'e2e', round(d['e2e']['ms_per_step'],2), d.get('check',{}).get('ok')); print(d['config'].get('phases_ms_max_over_ranks'))"
done
