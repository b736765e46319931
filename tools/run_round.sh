python -m pytest tests -m gpu -q -x -k "conv" 2>&1 | tail -3
python tools/conv_bench.py 2>&1 | grep "32,32,3"
python tools/config_bench.py tf32 2>&1 | tail -4
