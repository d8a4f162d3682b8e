python -m pytest tests/test_gpu_parity.py -x -q -k "op_gemm or forward_matches or batch_invariance" > gpurun_out/pytest_g1.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_g1.log
python tools/bench_gemm.py --ts 0 --variants 2 > gpurun_out/bench_gemm15.log 2>&1; tail -8 gpurun_out/bench_gemm15.log
timeout 300 python tools/ab_options.py > gpurun_out/ab2.log 2>&1; tail -3 gpurun_out/ab2.log
