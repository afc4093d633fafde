ncu --set full --clock-control none --import-source on -k regex:pgtt_gemm_kernel -s 10 -c 3 -f -o gpurun_out/r02d_gemm python tools/learner_gemm_time.py > gpurun_out/prof4.log 2>&1
ls -la gpurun_out/r02d_gemm.ncu-rep
