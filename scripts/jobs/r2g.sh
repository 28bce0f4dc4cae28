mkdir -p gpurun_out
timeout 200 python scripts/factor_sections.py 400 > gpurun_out/r2g_sections.log 2>&1; echo "sections rc=$?"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/r2g_launches_prof.csv python scripts/prof_one.py 400 1 > gpurun_out/r2g_ncu.log 2>&1; echo "ncu launches rc=$?"
python scripts/agg_launches.py gpurun_out/r2g_launches_prof.csv
timeout 200 python -m pytest tests/test_device_callers.py -x -q -m gpu 2>&1 | tail -3
