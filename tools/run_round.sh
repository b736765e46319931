timeout 120 python tools/attn_probe.py 2 4 256 384 --bwd 2>&1 | grep -E "err|rror"
timeout 120 python tools/attn_probe.py 1 3 200 77 --bwd 2>&1 | grep -E "err|rror"
for i in 1 2; do
NPM_B200_LIB=$PWD/tools/tmp/libnpm_prev.so timeout 120 python tools/attn_probe.py 8 16 1024 1024 --bwd --time 2>&1 | grep -E "fwd|bwd " | tr '\n' ' '; echo " <- before"
timeout 120 python tools/attn_probe.py 8 16 1024 1024 --bwd --time 2>&1 | grep -E "fwd|bwd " | tr '\n' ' '; echo " <- now"
done
timeout 600 python -m pytest tests -m gpu -q -x -k "causal or mha or decoder or attention or encoder" 2>&1 | tail -3
