for b in 0 8 4; do echo "band=$b"; NPM_GEMM_BAND=$b python tools/gemm_bench.py 2>&1 | sed -n 4,11p | cut -c1-150; done
python tools/gemm_probe.py km:tf32:4096x2048x512; python tools/gemm_probe.py mm:tf32:3072x1024x512; python tools/gemm_probe.py kk:tf32:2304x768x256
