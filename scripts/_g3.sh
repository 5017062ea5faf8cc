python -m pytest tests -m gpu -q > gpurun_out/t3_pytest.log 2>&1; echo "pytest rc=$?"; tail -8 gpurun_out/t3_pytest.log
python bench.py --no-cpu-baseline --no-extras > gpurun_out/t3_bench.json 2> gpurun_out/t3_bench.err; echo "bench rc=$?"
SAM2B200_NO_GEMM=1 python bench.py --no-cpu-baseline --no-extras --no-e2e > gpurun_out/t3_bench_nogemm.json 2> gpurun_out/t3_bench_nogemm.err; echo "bench rc=$?"
python - <<'PY'
import json
for f in ("t3_bench","t3_bench_nogemm"):
    d=json.load(open(f"gpurun_out/{f}.json")); print(f, d["ms_per_step"], d["kernel_families_ms_per_step"], d["roofline"]["frac"], d.get("e2e",{}) and d["e2e"].get("ms_per_step"))
PY
