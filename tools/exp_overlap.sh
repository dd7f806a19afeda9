python -m pytest tests/test_gpu_kernels.py -m gpu -x -q -k "elmat_cut or lane_per" > gpurun_out/s10_tests.log 2>&1; tail -2 gpurun_out/s10_tests.log
for cfg in "low 1 2" "low 1 1" "high 0 2" "high 0 1" "low 0 2" "high 1 1"; do
  set -- $cfg
  XGPU_SIDE_PRIO=$1 XGPU_SIDE_SHORT=$2 XGPU_CUT_BPSM=$3 python bench.py --no-cpu --steps 20 > gpurun_out/s10_$1_$2_$3.json 2> gpurun_out/s10_$1_$2_$3.err
  python -c "
import json,sys; d=json.load(open('gpurun_out/s10_$1_$2_$3.json')); print('$cfg', round(d['ms_per_step'],3), d['config']['phases_ms']['elmat_cut'])"
done
