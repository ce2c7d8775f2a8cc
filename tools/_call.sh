python -m pytest tests/test_gpu_conv.py tests/test_gpu_conv_rg.py -m gpu -x -q > gpurun_out/pytest_conv_opt.log 2>&1; tail -3 gpurun_out/pytest_conv_opt.log
rm -f gpurun_out/conv_opt_ab.jsonl
for i in 1 2; do
 python tools/conv_bench.py --plain >> gpurun_out/conv_opt_ab.jsonl 2>gpurun_out/conv_opt_ab.err
 CPLXK_CONV_AMAX_PASS=1 python tools/conv_bench.py --plain >> gpurun_out/conv_opt_ab.jsonl 2>>gpurun_out/conv_opt_ab.err
 CPLXK_LIB=$PWD/cplxmodule_b200/csrc/libcplxk_epi16.so python tools/conv_bench.py --plain >> gpurun_out/conv_opt_ab.jsonl 2>>gpurun_out/conv_opt_ab.err
done
cat gpurun_out/conv_opt_ab.jsonl
