for cfg in "default" "NCCL_MIN_P2P_NCHANNELS=16 NCCL_MAX_P2P_NCHANNELS=32" "NCCL_MIN_P2P_NCHANNELS=32 NCCL_MAX_P2P_NCHANNELS=32" "NCCL_MIN_P2P_NCHANNELS=8 NCCL_MAX_P2P_NCHANNELS=16"; do
  if [ "$cfg" = "default" ]; then E=""; else E="$cfg"; fi
  env $E python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 --scaling weak > gpurun_out/s18.json 2> gpurun_out/s18.err
  tail -n 1 gpurun_out/s18.json | python -c "
import json,sys; d=json.loads(sys.stdin.read()); p=d['config']['phases_ms_max_over_ranks']; print('$cfg', round(d['ms_per_step'],3), 'exchange', p['merge_exchange'], 'pack', p['merge_pack'])"
done
