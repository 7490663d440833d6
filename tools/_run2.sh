out=gpurun_out
timeout 500 python -m pytest tests/test_extraction_gpu.py tests/test_pipeline_gpu.py -x -q -m gpu > $out/pytest_r02_p1.log 2>&1; echo rc=$? >> $out/pytest_r02_p1.log
tail -3 $out/pytest_r02_p1.log
for rep in 1 2; do for v in variants/old.so speech_emotion_privacy_trust_b200/libsept_b200.so; do
  for nf in 800 1600; do
    ms=$(SEPT_LIB_PATH=$v timeout 200 python bench.py --steps 30 --warmup 3 --no-extras --n-fft $nf 2>/dev/null | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('%.4f' % d['ms_per_step'])")
    echo "$rep n_fft=$nf $v $ms"
  done
  echo "$rep mfcc $v $(SEPT_LIB_PATH=$v timeout 200 python tools/time_features.py 2>/dev/null | tail -1)"
done; done | tee $out/ab_pass1_merge.log
