set -x
python -m pytest tests/test_gpu_autoencoder.py -x -q > gpurun_out/pytest_ae2.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/pytest_ae2.log
timeout 300 python tools/check_ae.py bench 64 32 > gpurun_out/ae_bench4.log 2>&1; tail -9 gpurun_out/ae_bench4.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gn_apply --launch-skip 25 --launch-count 1 -o gpurun_out/prof_ae_gn2 -f python tools/check_ae.py bench 16 16 > gpurun_out/ae_ncu_c.log 2>&1; echo "ncu c rc=$?"
