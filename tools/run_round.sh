timeout 600 python -m pytest tests -m gpu -q -x 2>&1 | tail -3
for i in 1 2; do
NPM_NO_PDL=1 python bench.py --steps 10 --warmup 3 --no-cpu --no-alt 2>/dev/null | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('nopdl', d['ms_per_step'], d['value'], d['clocks']['sm_mhz'], d['final_loss'])"
python bench.py --steps 10 --warmup 3 --no-cpu --no-alt 2>/dev/null | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('pdl  ', d['ms_per_step'], d['value'], d['clocks']['sm_mhz'], d['final_loss'])"
done
