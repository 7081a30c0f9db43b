set -x
# the round-2 kernels (molecule-symmetrised 3B evaluator with its two g_k accumulator arrays, adapted-basis 2B evaluator,
# two-stream chains, sliced reduce) under compute-sanitizer
SEL="tests/test_gpu_parity.py::test_bare_polynomial tests/test_gpu_parity.py::test_bare_polynomial_ragged_batches_and_relabelling tests/test_gpu_parity.py::test_golden_fixture_host_arrays tests/test_gpu_parity.py::test_distributed_entry_points_on_a_single_rank_communicator"
timeout 900 compute-sanitizer --tool memcheck --log-file gpurun_out/r02_sanitizer2_memcheck.log python -m pytest $SEL -x -q 2>&1 | tail -3 > gpurun_out/r02_sanitizer2_memcheck_pytest.log
timeout 900 compute-sanitizer --tool racecheck --log-file gpurun_out/r02_sanitizer2_racecheck.log python -m pytest tests/test_gpu_parity.py::test_bare_polynomial tests/test_gpu_parity.py::test_per_tuple_parity -x -q 2>&1 | tail -3 > gpurun_out/r02_sanitizer2_racecheck_pytest.log
tail -3 gpurun_out/r02_sanitizer2_*.log
