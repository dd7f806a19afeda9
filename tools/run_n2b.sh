python -m pytest tests/test_distributed.py tests/test_pipeline.py -m gpu -x -q > gpurun_out/s21_tests.log 2>&1; tail -n 2 gpurun_out/s21_tests.log
for sc in weak strong; do
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 3 --scaling $sc > gpurun_out/s21_n2_$sc.json 2> gpurun_out/s21_n2_$sc.err
tail -n 1 gpurun_out/s21_n2_$sc.json | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('$sc', d['n_gpus'], round(d['ms_per_step'],3), round(d['value']/1e6,1), 'e2e', round(d['e2e']['ms_per_step'],2), d.get('check',{}).get('ok')); print(d['config'].get('phases_ms_max_over_ranks'))"
done
python bench.py --no-cpu > gpurun_out/s21_n1.json 2>/dev/null; python -c "
import json; d=json.load(open('gpurun_out/s21_n1.json')); print('n1', round(d['ms_per_step'],3), d['config']['phases_ms'])"
