mkdir -p gpurun_out
timeout 180 python -m pytest tests -m gpu -q -x --timeout 120 --timeout-method=thread -p no:cacheprovider -k "single_op and (Invert or Rotate or Sharpness or Equalize or Color)" > gpurun_out/pytest_quick.log 2>&1; rc=$?; echo "quick exit $rc" >> gpurun_out/pytest_quick.log; tail -5 gpurun_out/pytest_quick.log
if [ $rc -eq 0 ]; then bash tools/gpu_check.sh 4096; fi
