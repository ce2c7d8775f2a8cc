timeout 600 python -m pytest tests -m gpu -q > gpurun_out/pytest_final3.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_final3.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
timeout 300 python tools/config_bench.py > gpurun_out/configs_r2c.log 2>&1; echo "cfg rc=$?"
timeout 400 python bench.py > gpurun_out/bench_final3.json 2> gpurun_out/bench_final3.err; echo "bench rc=$?"; python -c "
import json;d=json.loads(open('gpurun_out/bench_final3.json').read().strip().splitlines()[-1]);print(d['value'],d['ms_per_step'],d['roofline']['frac'],d['e2e']['value'],d['parity']['ok'],json.dumps(d['extra'].get('config4')))"
