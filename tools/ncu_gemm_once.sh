python tools/gemm_one.py km 8192 4096 1024 3 > gpurun_out/gemm_one_plain.log 2>&1; echo "plain rc=$?"; cat gpurun_out/gemm_one_plain.log
ncu --set full --clock-control none --import-source on -k regex:gemm_tc2 -c 1 -f -o gpurun_out/prof_gemm4 python tools/gemm_one.py km 8192 4096 1024 1 > gpurun_out/ncu2.log 2>&1; echo "ncu rc=$?"
