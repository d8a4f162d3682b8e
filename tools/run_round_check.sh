set -x
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu8.log 2>&1; echo "pytest rc=$?"
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke2.log 2>&1; echo "smoke rc=$?"
python bench.py > gpurun_out/bench9.json 2> gpurun_out/bench9.err; echo "bench rc=$?"
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench9_ref.json 2> gpurun_out/bench9_ref.err; echo "ref rc=$?"
tail -3 gpurun_out/pytest_gpu8.log
