#!/bin/bash
# round-end validation, part 1: every -m gpu test, smoke(), memcheck (allocator caching off) over the kernels added this session
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/ -q -m gpu -p no:cacheprovider -rf > gpurun_out/final_test_all.log 2>&1; echo "== pytest -m gpu exit $?"; grep -v PASSED gpurun_out/final_test_all.log | tail -8
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/final_smoke.log 2>&1; echo "== smoke exit $?"; tail -n 2 gpurun_out/final_smoke.log
PYTORCH_NO_CUDA_MEMORY_CACHING=1 timeout 420 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_ops_gpu.py -q -m gpu -p no:cacheprovider -x "--deselect=tests/test_ops_gpu.py::test_conv3x3_streamk[0]" \
  -k "(streamk and 2048 and (plain or bias_res)) or (streamk and 1000) or (up2x and 1-8-8) or (up2x and 2-16-16) or (folded and 300-64) or (folded and 520-96) or conv3x3_streamk" \
  > gpurun_out/final_sanitizer.log 2>&1; echo "== memcheck exit $?"; grep -E "ERROR SUMMARY|passed|failed|Invalid|Error" gpurun_out/final_sanitizer.log | tail -6
