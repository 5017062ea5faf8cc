set -x
python -m pytest tests -m gpu -x -q > gpurun_out/t1_pytest.log 2>&1; echo "pytest rc=$?" 
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/t1_smoke.log 2>&1; echo "smoke rc=$?"
python bench.py > gpurun_out/t1_bench.json 2> gpurun_out/t1_bench.err; echo "bench rc=$?"
python bench.py --no-graphs --steps 1 --warmup 3 --no-e2e --no-cpu-baseline --no-extras > gpurun_out/t1_plain.log 2>&1; echo "plain rc=$?"
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/t1_launches.csv python bench.py --no-graphs --steps 1 --warmup 1 --no-e2e --no-cpu-baseline --no-extras > gpurun_out/t1_ncu.log 2>&1; echo "ncu rc=$?"
tail -3 gpurun_out/t1_pytest.log; tail -2 gpurun_out/t1_smoke.log
