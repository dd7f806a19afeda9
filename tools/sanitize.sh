#!/bin/bash
# compute-sanitizer over a slice of the GPU suite (memcheck: every kernel family; racecheck: the shared-memory kernels).
# logs -> gpurun_out/r02_sanitizer_*.log (summaries copied to profiles/)
set -x
export PYTHONUNBUFFERED=1
S=/usr/local/cuda/bin/compute-sanitizer
timeout 900 $S --tool memcheck --error-exitcode 9 --log-file gpurun_out/r02_sanitizer_memcheck.log \
  python -m pytest tests/test_gpu_kernels.py tests/test_gpu_lsetcurv.py tests/test_gpu_csr_fused.py -m gpu -q -x \
  -k "not intcurved and not tensor_tiles and not facetpatch_tensor" > gpurun_out/r02_sanitizer_memcheck_pytest.log 2>&1
echo "memcheck rc=$?" >> gpurun_out/r02_sanitizer_memcheck_pytest.log
timeout 600 $S --tool racecheck --error-exitcode 9 --log-file gpurun_out/r02_sanitizer_racecheck.log \
  python -m pytest tests/test_gpu_kernels.py tests/test_gpu_lsetcurv.py -m gpu -q -x \
  -k "elmat_cut_pair or lane_per_point or calc_deformation or csr_pattern" > gpurun_out/r02_sanitizer_racecheck_pytest.log 2>&1
echo "racecheck rc=$?" >> gpurun_out/r02_sanitizer_racecheck_pytest.log
tail -3 gpurun_out/r02_sanitizer_memcheck_pytest.log gpurun_out/r02_sanitizer_racecheck_pytest.log
tail -5 gpurun_out/r02_sanitizer_memcheck.log gpurun_out/r02_sanitizer_racecheck.log
