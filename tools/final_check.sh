python -m pytest tests -x -q -m gpu > gpurun_out/tests_final.log 2>&1; tail -2 gpurun_out/tests_final.log
python __graft_entry__.py smoke
python bench.py > gpurun_out/r2d_bench_n1.json 2> gpurun_out/r2d_bench_n1.err; tail -c 200 gpurun_out/r2d_bench_n1.json
python bench.py --impl reference > gpurun_out/r2d_bench_reference_arm.json 2>/dev/null
python - <<'EOF'
import json
r=json.load(open("gpurun_out/r2d_bench_reference_arm.json")); d=json.load(open("gpurun_out/r2d_bench_n1.json"))
print("ref", r["value"], "cpu_baseline", d["cpu_baseline"]["value"], "ours", d["value"], d["e2e"]["value"], "C3", d["configs"]["C3_full"]["value"], d["configs"]["C3_full"]["e2e"]["value"], r["configs"])
print(json.dumps(d["configs"]["per_frame"])[:1200])
print(d["configs"]["C4"])
EOF
