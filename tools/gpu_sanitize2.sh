#!/bin/bash
# round 2: compute-sanitizer over the kernels added this round — block kernels (product spaces, skeleton, interior penalty), matrix
# sums, K4, the fused unstructured cell kernel, field forms, the final affine sweep (orthogonal and six-coefficient bodies)
mkdir -p gpurun_out
export PYTHONDONTWRITEBYTECODE=1
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
SEL_MEM='tests/test_multifield.py tests/test_gpu_k4.py tests/test_gpu_unstructured.py tests/test_gpu_field.py::test_plaplacian_reference_golden_l2_norm tests/test_gpu_fastpath.py::test_affine_kernel_on_sheared_meshes tests/test_gpu_cartesian.py'
timeout 700 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest -m gpu -x -q -p no:cacheprovider $SEL_MEM > gpurun_out/sanitize2_memcheck.log 2>&1
echo "memcheck rc=$?"; grep "ERROR SUMMARY\|passed\|failed" gpurun_out/sanitize2_memcheck.log | tail -3
SEL_RACE='tests/test_multifield.py::test_gpu_interior_penalty_blocks_parity tests/test_multifield.py::test_gpu_stokes_blocks_parity tests/test_gpu_k4.py tests/test_gpu_fastpath.py::test_affine_kernel_on_sheared_meshes'
timeout 500 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest -m gpu -x -q -p no:cacheprovider $SEL_RACE > gpurun_out/sanitize2_racecheck.log 2>&1
echo "racecheck rc=$?"; grep "RACECHECK SUMMARY\|passed\|failed" gpurun_out/sanitize2_racecheck.log | tail -3
