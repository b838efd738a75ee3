mkdir -p gpurun_out/r2
timeout 900 python -m pytest tests -q -m gpu 2>&1 | tail -n 4
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -n 3
timeout 600 python bench.py --steps 20 --warmup 3 > gpurun_out/r2/bench_final_c2.json 2> gpurun_out/r2/bench_final_c2.err; tail -c 300 gpurun_out/r2/bench_final_c2.err
timeout 600 python bench.py --impl reference --steps 2 --warmup 0 > gpurun_out/r2/bench_final_ref.json 2> gpurun_out/r2/bench_final_ref.err; tail -c 300 gpurun_out/r2/bench_final_ref.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2/bench_final_c2.json').read().strip().splitlines()[-1])
print('native', d['value'], d['e2e'], d['ms_per_step'], d['aux']['serial_step_ms'], d['roofline']['frac'], d['roofline']['traffic'], d['roofline']['fp64'], d['clocks'])
r=json.loads(open('gpurun_out/r2/bench_final_ref.json').read().strip().splitlines()[-1])
print('ref', r['value'], r['steps'], r['cpu_baseline'], [ (o['seed'],o['status'],o['scp_iterations'],o['minsep_pass'],o['dyn_pass']) for o in r['outcomes']][:20])
PY
