ncu --metrics gpu__time_duration.sum --clock-control none -s 4000 -c 600 --csv --log-file gpurun_out/r02d_learner_launches.csv python tools/learner_time.py highest > gpurun_out/prof5.log 2>&1
tail -2 gpurun_out/prof5.log
