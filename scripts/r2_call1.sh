set -x
mkdir -p gpurun_out/r2c1
python -m pytest tests -m gpu -x -q > gpurun_out/r2c1/pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2c1/pytest.log
python bench.py > gpurun_out/r2c1/bench.json 2> gpurun_out/r2c1/bench.err
timeout 600 compute-sanitizer --tool memcheck --log-file gpurun_out/r2c1/memcheck_smoke.log python __graft_entry__.py smoke > gpurun_out/r2c1/memcheck_smoke.out 2>&1; echo "rc=$?" >> gpurun_out/r2c1/memcheck_smoke.out
timeout 900 compute-sanitizer --tool racecheck --log-file gpurun_out/r2c1/racecheck_smoke.log python __graft_entry__.py smoke > gpurun_out/r2c1/racecheck_smoke.out 2>&1; echo "rc=$?" >> gpurun_out/r2c1/racecheck_smoke.out
timeout 900 compute-sanitizer --tool memcheck --log-file gpurun_out/r2c1/memcheck_kernels.log python -m pytest tests/test_kernels_gpu.py -m gpu -x -q -k "attention or gemm_inplace or conv_as or posconv or layernorm" > gpurun_out/r2c1/memcheck_kernels.out 2>&1; echo "rc=$?" >> gpurun_out/r2c1/memcheck_kernels.out
timeout 900 compute-sanitizer --tool racecheck --log-file gpurun_out/r2c1/racecheck_attn.log python -m pytest tests/test_kernels_gpu.py -m gpu -x -q -k "attention_random_ragged" > gpurun_out/r2c1/racecheck_attn.out 2>&1; echo "rc=$?" >> gpurun_out/r2c1/racecheck_attn.out
tail -3 gpurun_out/r2c1/*.out gpurun_out/r2c1/pytest.log
