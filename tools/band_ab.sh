for b in 0 8 4 16; do
NPM_GEMM_BAND=$b python bench.py --steps 10 --warmup 3 --no-cpu --no-alt 2>/dev/null | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('band $b', d['ms_per_step'], d['value'], d['clocks']['sm_mhz'], d['roofline']['achieved'])"
done
