timeout 120 python tools/attn_probe.py 8 16 1024 1024 --bwd --time --causal 2>&1 | grep -E "fwd|bwd"
timeout 600 python -m pytest tests -m gpu -q -x -k "causal" 2>&1 | tail -2
