set -x
nvidia-smi -L
python -m pytest tests -m gpu -x -q 2>&1 | tail -5 > gpurun_out/r02_base_gputests.log
python bench.py --steps 10 --warmup 3 > gpurun_out/r02_base_bench.json 2> gpurun_out/r02_base_bench.err
SEL='tests/test_gpu_parity.py::test_sliced_list_build_matches_full tests/test_gpu_parity.py::test_lists_edge_cases tests/test_gpu_parity.py::test_row_builder_fast_path_and_fallbacks tests/test_gpu_parity.py::test_run_to_run_determinism_is_bitwise'
timeout 900 compute-sanitizer --tool memcheck --log-file gpurun_out/r02_sanitizer_memcheck.log python -m pytest $SEL tests/test_gpu_parity.py::test_synthetic_4096_against_oracle -x -q 2>&1 | tail -3 > gpurun_out/r02_sanitizer_memcheck_pytest.log
timeout 1200 compute-sanitizer --tool racecheck --log-file gpurun_out/r02_sanitizer_racecheck.log python -m pytest $SEL -x -q 2>&1 | tail -3 > gpurun_out/r02_sanitizer_racecheck_pytest.log
timeout 600 compute-sanitizer --tool synccheck --log-file gpurun_out/r02_sanitizer_synccheck.log python -m pytest tests/test_gpu_parity.py::test_lists_edge_cases tests/test_gpu_parity.py::test_run_to_run_determinism_is_bitwise -x -q 2>&1 | tail -3 > gpurun_out/r02_sanitizer_synccheck_pytest.log
lscpu | head -20 > gpurun_out/r02_lscpu.txt
ls /usr/include/nccl.h /usr/lib/x86_64-linux-gnu/libnccl* 2>&1 > gpurun_out/r02_nccl_files.txt
