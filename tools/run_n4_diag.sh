mkdir -p gpurun_out
for sc in weak strong; do
XGPU_DUMP_PHASES=1 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 4 --steps 10 --warmup 3 --scaling $sc > gpurun_out/s17_n4_$sc.json 2> gpurun_out/s17_n4_$sc.err
tail -n 1 gpurun_out/s17_n4_$sc.json | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('$sc', d['n_gpus'], round(d['ms_per_step'],3), round(d['value']/1e6,1), 'e2e', round(d['e2e']['ms_per_step'],2), d.get('check',{}).get('ok'))"
for r in 0 1 2 3; do echo "rank $r: $(cat gpurun_out/phases_rank${r}_of4_$sc.json)"; done
done
