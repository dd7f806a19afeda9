for i in 1 2 3; do python bench.py --no-cpu > gpurun_out/s12_c$i.json 2>/dev/null; python -c "
import json; d=json.load(open('gpurun_out/s12_c$i.json')); print('concurrent', round(d['ms_per_step'],3), round(d['e2e']['ms_per_step'],2))"; done
for i in 1 2 3; do XGPU_SERIAL=1 python bench.py --no-cpu > gpurun_out/s12_s$i.json 2>/dev/null; python -c "
import json; d=json.load(open('gpurun_out/s12_s$i.json')); print('serial', round(d['ms_per_step'],3), round(d['e2e']['ms_per_step'],2))"; done
