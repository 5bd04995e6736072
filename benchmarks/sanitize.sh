#!/bin/bash
# Run under gpurun: compute-sanitizer racecheck / synccheck / memcheck over the
# GPU tests that exercise the shared-memory kernels (mbarrier rings of the TMA
# kernels, the in-place shared-memory FFT, the regrid staging).  Full-size cases
# are deselected (the tools slow kernels down 10-100x).
#   gpurun --timeout 2400 -- 'bash benchmarks/sanitize.sh'
mkdir -p gpurun_out
SEL='not fullsize and not full_size and not headline and not 721 and not m50 and not correlated'
TESTS="tests/test_spectrum_pfa_gpu.py tests/test_regrid_tma_gpu.py tests/test_det_metrics_gpu.py tests/test_host_streaming_gpu.py tests/test_ens_metrics_gpu.py"
for tool in racecheck synccheck memcheck; do
  timeout 900 compute-sanitizer --tool $tool --print-limit 20 \
    python -m pytest $TESTS -x -q -m gpu -k "$SEL" -p no:cacheprovider \
    > gpurun_out/r2_sanitizer_$tool.txt 2>&1
  echo "== $tool rc=$?"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|passed|failed|hazard" gpurun_out/r2_sanitizer_$tool.txt | tail -5
done
