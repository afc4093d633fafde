python tools/mlp_time.py
ncu --set full --clock-control none --import-source on -k regex:pgtt_bgemm_kernel -s 20 -c 12 -f -o gpurun_out/r02f_bgemm python tools/mlp_time.py > /dev/null 2>&1
ls -la gpurun_out/r02f_bgemm.ncu-rep
