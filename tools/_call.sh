timeout 400 python -m pytest tests/test_gpu_conv_rg.py tests/test_gpu_conv.py tests/test_gpu_backward.py -m gpu -q -x > gpurun_out/pytest_narrow.log 2>&1; echo "rc=$?"; tail -4 gpurun_out/pytest_narrow.log
rm -f gpurun_out/conv_real_narrow_ab.jsonl
for nr in 1 0 1 0; do
 CPLXK_CONV_REAL_NARROW=$nr timeout 200 python tools/conv_real_ab.py >> gpurun_out/conv_real_narrow_ab.jsonl 2>gpurun_out/conv_real_narrow_ab.err
done
cat gpurun_out/conv_real_narrow_ab.jsonl | cut -c60-
timeout 200 python tools/convvd_probe.py 2>&1 | head -4
