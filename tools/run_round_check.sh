# end-of-round validation on one B200: tests, smoke, bench (both arms)
set -x
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu12.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu12.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke6.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/smoke6.log
python bench.py > gpurun_out/bench13.json 2> gpurun_out/bench13.err; echo "bench rc=$?"
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench13_ref.json 2> gpurun_out/bench13_ref.err; echo "ref rc=$?"
