set -x
python -m pytest tests -m gpu -q 2>&1 | tail -3 | cut -c1-300
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
ncu --set full --clock-control none --import-source on -k regex:pgtt_env_kernel -s 6 -c 1 -f -o gpurun_out/r02i_warp python bench.py --steps 4 --warmup 3 --no-cpu-baseline --no-extras > /dev/null 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -s 60 -c 200 --csv --log-file gpurun_out/r02i_launches.csv python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-extras > /dev/null 2>&1
ls -la gpurun_out/r02i_*
