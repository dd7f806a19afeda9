#!/bin/bash
# compute-sanitizer memcheck + racecheck over the kernels touched in session 4 (facet patch with downscale, row-range values
# pass, half-width sorting network, codim-2 boundary rule) -> gpurun_out/r02_sanitizer_s4_*.log
export PYTHONUNBUFFERED=1
S=/usr/local/cuda/bin/compute-sanitizer
SEL="downscale or row_ranges or half_width or if_bnd or periodic"
timeout 500 $S --tool memcheck --error-exitcode 9 --log-file gpurun_out/r02_sanitizer_s4_memcheck.log \
  python -m pytest tests/test_gpu_kernels.py tests/test_gpu_csr_fused.py tests/test_gpu_bnd.py tests/test_periodic_facets.py -m gpu -q -x \
  -k "$SEL" > gpurun_out/r02_sanitizer_s4_memcheck_pytest.log 2>&1
echo "memcheck rc=$?" >> gpurun_out/r02_sanitizer_s4_memcheck_pytest.log
timeout 400 $S --tool racecheck --error-exitcode 9 --log-file gpurun_out/r02_sanitizer_s4_racecheck.log \
  python -m pytest tests/test_gpu_kernels.py tests/test_gpu_csr_fused.py -m gpu -q -x \
  -k "(downscale and 0.5) or half_width" > gpurun_out/r02_sanitizer_s4_racecheck_pytest.log 2>&1
echo "racecheck rc=$?" >> gpurun_out/r02_sanitizer_s4_racecheck_pytest.log
tail -n 3 gpurun_out/r02_sanitizer_s4_memcheck_pytest.log gpurun_out/r02_sanitizer_s4_racecheck_pytest.log
tail -n 4 gpurun_out/r02_sanitizer_s4_memcheck.log gpurun_out/r02_sanitizer_s4_racecheck.log
