CPLXK_CONV_CVT_SEQ=1 timeout 300 python -m pytest tests/test_gpu_conv.py -m gpu -q -x -k "scaled_fp16 or chunked or config4 or row_mode" > gpurun_out/pytest_seq.log 2>&1; echo "rc=$?"; tail -3 gpurun_out/pytest_seq.log
L=$PWD/cplxmodule_b200/csrc/libcplxk_ctrace.so
for ov in 4 8; do
 echo "== seq=1 overlap $ov"
 CPLXK_LIB=$L CPLXK_CONV_CVT_SEQ=1 CPLXK_CONV_TRACE=1 CPLXK_CONV_OVERLAP=$ov timeout 120 python tools/prof_conv.py 3 f32 nchw 2>&1 | tail -$((ov+1)) | head -$ov
done
rm -f gpurun_out/conv_seq_ab.jsonl
for cfg in "0 4" "1 4" "1 8" "0 8" "1 4" "0 4" "1 0" "0 0"; do
 set -- $cfg
 echo "{\"CPLXK_CONV_CVT_SEQ\": $1}" >> gpurun_out/conv_seq_ab.jsonl
 CPLXK_CONV_CVT_SEQ=$1 CPLXK_CONV_OVERLAP=$2 timeout 120 python tools/conv_bench.py --plain --fp32-nchw >> gpurun_out/conv_seq_ab.jsonl 2>gpurun_out/conv_seq_ab.err
done
cat gpurun_out/conv_seq_ab.jsonl
