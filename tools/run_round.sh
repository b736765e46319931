timeout 120 python tools/attn_probe.py 2 4 256 384 --bwd 2>&1 | grep -E "err|rror"
timeout 120 python tools/attn_probe.py 1 3 200 77 --bwd 2>&1 | grep -E "err|rror"
timeout 120 python tools/attn_probe.py 1 2 130 300 --bwd 2>&1 | grep -E "err|rror"
timeout 120 python tools/attn_probe.py 8 16 1024 1024 --bwd --time 2>&1 | grep -E "fwd|bwd|err"
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:attn_ -c 4 python tools/attn_probe.py 8 16 1024 1024 --bwd 2>&1 | grep -E "attn_|gpu__time" | sed 's/(.*//'
