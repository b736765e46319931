set -x
python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke(); print('SMOKE_OK')" 2>&1 | tail -2
python bench.py > gpurun_out/bench_final.log 2>&1; echo "bench rc=$?"; tail -1 gpurun_out/bench_final.log | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['achieved'], d['roofline']['frac'], d['clocks'], d['gpu_launches'])"
python bench.py --layers 2 --steps 2 --warmup 3 --no-alt --no-cpu > gpurun_out/plain2.log 2>&1; echo "plain rc=$?"
ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/launches.csv python bench.py --layers 2 --steps 2 --warmup 3 --no-alt --no-cpu > gpurun_out/ncu1.log 2>&1; echo "ncu rc=$?"
