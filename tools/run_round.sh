python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu.log
python bench.py --steps 5 --warmup 3 > gpurun_out/bench_elect.log 2>&1; echo "bench rc=$?"; tail -1 gpurun_out/bench_elect.log | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['clocks'], d['roofline']['achieved'], d['roofline']['frac'], d['e2e']['value'], d['alt'])"
python tools/conv_bench.py 2>&1 | tail -8
