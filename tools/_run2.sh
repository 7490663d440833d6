out=gpurun_out
timeout 400 python -m pytest tests/test_resample.py -x -q -m gpu > $out/pytest_r02_rs.log 2>&1; echo rc=$? >> $out/pytest_r02_rs.log
tail -3 $out/pytest_r02_rs.log
timeout 200 python tools/time_resample.py 2>&1 | tail -1 | tee $out/time_resample_r02.log
timeout 400 ncu --set full --clock-control none --import-source on -k regex:resample_tiled -s 2 -c 1 -o $out/prof_r02_resample_tiled -f python tools/time_resample.py > $out/ncu_r02_rs.log 2>&1
