# end-of-round validation on one B200: tests, smoke, bench (both arms), ncu launch list of the bench command
set -x
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu11.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu11.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke5.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/smoke5.log
python bench.py > gpurun_out/bench12.json 2> gpurun_out/bench12.err; echo "bench rc=$?"
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench12_ref.json 2> gpurun_out/bench12_ref.err; echo "ref rc=$?"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k 'regex:gemm|attention|token|patch|conv3x3|ddpm|ln_stats|finalize|fill_t|dec_t' -c 700 --csv --log-file gpurun_out/launches_r01c.csv python bench.py --steps 1 --warmup 3 --e2e-steps 1 --no-cpu > gpurun_out/ncu_launches_c.log 2>&1; echo "ncu list rc=$?"
