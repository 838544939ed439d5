set -x
timeout -k 10 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -3
timeout -k 10 300 python scripts/step_once.py > gpurun_out/step_once.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/r2j_launches.csv timeout -k 10 600 python scripts/step_once.py > gpurun_out/ncu_j1.log 2>&1
timeout -k 10 300 python scripts/profile_step.py > gpurun_out/r2j_warm_step_profile.txt 2>&1
timeout -k 10 300 python scripts/timeline_step.py > gpurun_out/r2j_timeline.txt 2>&1
PROBE_H=300 timeout -k 10 300 python scripts/mincut_small_probe.py 2>&1 | tail -1
PROBE_H=300 ncu --set full --clock-control none --import-source on -k regex:mincut_pool_x -c 2 --launch-skip 150 -o gpurun_out/r2j_poolx timeout -k 10 300 python scripts/mincut_small_probe.py > gpurun_out/ncu_j2.log 2>&1
PROBE_H=300 PROBE_SKIP_FWD=1 ncu --set full --clock-control none --import-source on -k regex:mincut_pool_x_bwd -c 1 --launch-skip 4 -o gpurun_out/r2j_poolx_bwd timeout -k 10 300 python scripts/mincut_small_probe.py > gpurun_out/ncu_j3.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:jacobi_eig -c 1 --launch-skip 1 -o gpurun_out/r2j_jacobi timeout -k 10 600 python scripts/posenc_probe.py > gpurun_out/ncu_j4.log 2>&1
timeout -k 10 300 python scripts/posenc_probe.py 2>&1 | tail -2
