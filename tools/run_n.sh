# usage: run_n.sh N  -> weak and strong bench lines at N GPUs
N=$1
for sc in weak strong; do
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 10 --warmup 3 --scaling $sc > gpurun_out/s14_n${N}_$sc.json 2> gpurun_out/s14_n${N}_$sc.err
tail -n 1 gpurun_out/s14_n${N}_$sc.json | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('$sc', d['n_gpus'], round(d['ms_per_step'],3), round(d['value']/1e6,1), 'e2e', round(d['e2e']['ms_per_step'],2), d.get('check',{}).get('ok')); print(d['config'].get('phases_ms_max_over_ranks'))"
done
