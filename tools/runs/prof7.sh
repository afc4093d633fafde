ncu --metrics gpu__time_duration.sum --clock-control none -s 4000 -c 700 --csv --log-file gpurun_out/r02g_learner_launches.csv python tools/learner_time.py highest > gpurun_out/prof7_learner.log 2>&1
tail -2 gpurun_out/prof7_learner.log
