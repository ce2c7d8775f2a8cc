T="tests/test_gpu_conv.py -k row_mode_conv_shapes"
timeout 150 python -m pytest $T -m gpu -q -x > gpurun_out/pytest_row_bo.log 2>&1; echo "baseoff rc=$?"; tail -4 gpurun_out/pytest_row_bo.log
CPLXK_LIB=$PWD/cplxmodule_b200/csrc/libcplxk_nobo.so timeout 150 python -m pytest $T -m gpu -q -x > gpurun_out/pytest_row_nobo.log 2>&1; echo "nobo rc=$?"; tail -4 gpurun_out/pytest_row_nobo.log
rm -f gpurun_out/conv_row_ab.jsonl
for i in 1 2; do
 timeout 120 python tools/conv_bench.py --plain >> gpurun_out/conv_row_ab.jsonl 2>gpurun_out/conv_row_ab.err
 CPLXK_CONV_ROW=0 timeout 120 python tools/conv_bench.py --plain >> gpurun_out/conv_row_ab.jsonl 2>>gpurun_out/conv_row_ab.err
done
cat gpurun_out/conv_row_ab.jsonl
