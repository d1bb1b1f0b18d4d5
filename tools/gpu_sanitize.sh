# compute-sanitizer over the small golden configurations, one tool per call:  bash tools/gpu_sanitize.sh memcheck
TOOL=${1:-memcheck}
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
SEL="tests/test_gpu_parity.py::test_cabi_against_reference_goldens tests/test_gpu_parity.py::test_parameter_gradients_against_goldens tests/test_gpu_parity.py::test_value_and_grad_single_pass tests/test_gpu_parity.py::test_big_frame_preprocess_kernel_sets tests/test_gpu_wide.py::test_prepared_forward_against_reference_goldens tests/test_gpu_wide.py::test_wide_kernel_variants tests/test_gpu_jacobian.py::test_jacobian_small_batches_and_unaligned"
timeout -k 20 1500 compute-sanitizer --tool $TOOL --error-exitcode 7 --log-file gpurun_out/sanitizer_$TOOL.log \
  python -m pytest $SEL -x -q -p no:cacheprovider > gpurun_out/sanitizer_${TOOL}_pytest.log 2>&1
echo "sanitizer $TOOL rc=$?" | tee -a gpurun_out/sanitizer_${TOOL}_pytest.log
tail -5 gpurun_out/sanitizer_${TOOL}_pytest.log
grep -c "ERROR SUMMARY" gpurun_out/sanitizer_$TOOL.log; grep "ERROR SUMMARY" gpurun_out/sanitizer_$TOOL.log | sort | uniq -c | head
grep -B2 -A12 "Invalid\|Race\|Uninitialized\|hazard" gpurun_out/sanitizer_$TOOL.log | head -80
