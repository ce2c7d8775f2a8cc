timeout 400 python -m pytest tests/test_gpu_conv_rg.py tests/test_gpu_conv.py tests/test_gpu_backward.py tests/test_gpu_graph.py -m gpu -q -x > gpurun_out/pytest_abs2.log 2>&1; echo "rc=$?"; tail -4 gpurun_out/pytest_abs2.log
timeout 200 python tools/convvd_probe.py 2>&1 | head -4 > gpurun_out/convvd_probe3.jsonl; cat gpurun_out/convvd_probe3.jsonl
