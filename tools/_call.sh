timeout 200 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_final4.log 2>&1; echo "rc=$?"; tail -2 gpurun_out/pytest_final4.log
