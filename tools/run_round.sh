for i in 1 2; do
NPM_NO_COLSUM_RIDE=1 python bench.py --steps 10 --warmup 3 --no-cpu --no-alt 2>/dev/null | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('off', d['ms_per_step'], d['value'], d['clocks']['sm_mhz'], d['gpu_launches'])"
python bench.py --steps 10 --warmup 3 --no-cpu --no-alt 2>/dev/null | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('on ', d['ms_per_step'], d['value'], d['clocks']['sm_mhz'], d['gpu_launches'])"
done
