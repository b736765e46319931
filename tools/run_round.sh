for i in 1 2; do
NPM_GEMM_NO_CHUNKED_MN=1 python bench.py --steps 5 --warmup 3 --no-cpu --no-alt 2>/dev/null | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('nochunk', d['ms_per_step'], d['clocks']['sm_mhz'], d['roofline']['achieved'])"
python bench.py --steps 5 --warmup 3 --no-cpu --no-alt 2>/dev/null | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('chunked', d['ms_per_step'], d['clocks']['sm_mhz'], d['roofline']['achieved'])"
done
