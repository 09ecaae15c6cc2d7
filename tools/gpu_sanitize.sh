#!/bin/bash
# compute-sanitizer over a cross-section of the paths: memcheck (global/shared OOB, misaligned) and racecheck (shared-memory hazards)
mkdir -p gpurun_out
export PYTHONDONTWRITEBYTECODE=1
SEL_MEM='tests/test_gpu_fastpath.py::test_fast_path_matches_oracle_and_generic[cells2-boundary-0.2] tests/test_gpu_fastpath.py::test_fast_path_matches_oracle_and_generic[cells1-boundary-0.2] tests/test_gpu_fastpath.py::test_fast_path_matches_oracle_and_generic[cells4-bc4-0.1] tests/test_gpu_fastpath.py::test_structured_symbolic_equals_sort_based tests/test_gpu_dmma.py tests/test_gpu_neumann.py tests/test_gpu_linear_problem.py tests/test_golden.py'
timeout 500 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest -m gpu -x -q -p no:cacheprovider $SEL_MEM > gpurun_out/sanitize_memcheck.log 2>&1
echo "memcheck rc=$?"; grep "ERROR SUMMARY\|passed\|failed" gpurun_out/sanitize_memcheck.log | tail -3
SEL_RACE='tests/test_gpu_fastpath.py::test_fast_path_matches_oracle_and_generic[cells1-boundary-0.2] tests/test_gpu_fastpath.py::test_fast_path_matches_oracle_and_generic[cells0-boundary-0.0] tests/test_gpu_dmma.py::test_dmma_matches_oracle_and_generic[cells0-3-False-boundary-0.15]'
timeout 500 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest -m gpu -x -q -p no:cacheprovider $SEL_RACE > gpurun_out/sanitize_racecheck.log 2>&1
echo "racecheck rc=$?"; grep "RACECHECK SUMMARY\|passed\|failed" gpurun_out/sanitize_racecheck.log | tail -3
