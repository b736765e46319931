python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu.log
python tools/gemm_one.py km 8192 4096 1024 3 > gpurun_out/gemm_one_plain.log 2>&1; echo "gemm_one rc=$?"; cat gpurun_out/gemm_one_plain.log
ncu --set full --clock-control none --import-source on -k regex:gemm_tc2 -c 1 -f -o gpurun_out/prof_gemm3 python tools/gemm_one.py km 8192 4096 1024 1 > gpurun_out/ncu2.log 2>&1; echo "ncu gemm rc=$?"
python tools/attn_probe.py 8 16 1024 1024 --bwd --time > gpurun_out/attn_plain.log 2>&1; echo "attn rc=$?"; grep -E "fwd|bwd" gpurun_out/attn_plain.log
ncu --set full --clock-control none --import-source on -k regex:attn_ -c 4 -f -o gpurun_out/prof_attn4 python tools/attn_probe.py 8 16 1024 1024 --bwd > gpurun_out/ncu3.log 2>&1; echo "ncu attn rc=$?"
