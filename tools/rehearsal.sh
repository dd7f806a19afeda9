# the driver's round-end sequence on a fresh box (1 GPU part): smoke, reference arm, bench
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -n 2
python bench.py --impl reference --gpus 1 --steps 2 --warmup 1 > gpurun_out/s16_ref.json 2> gpurun_out/s16_ref.err; tail -c 700 gpurun_out/s16_ref.json; echo
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --impl reference --gpus 2 --steps 1 --warmup 1 > gpurun_out/s16_ref_n2.json 2> gpurun_out/s16_ref_n2.err; echo "rc=$?"; tail -c 400 gpurun_out/s16_ref_n2.json; echo
python bench.py > gpurun_out/s16_bench.json 2> gpurun_out/s16_bench.err; python -c "
import json; d=json.load(open('gpurun_out/s16_bench.json')); print(round(d['ms_per_step'],3), round(d['e2e']['ms_per_step'],1), d['roofline']['frac'], d['roofline']['traffic'], d['cpu_baseline']['value'], d['gpu_launches'], d['clocks'])"
