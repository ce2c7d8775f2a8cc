L=$PWD/cplxmodule_b200/csrc/libcplxk_ctrace.so
for ov in 4 8; do
 echo "== overlap $ov"
 CPLXK_LIB=$L CPLXK_CONV_TRACE=1 CPLXK_CONV_OVERLAP=$ov timeout 120 python tools/prof_conv.py 3 f32 nchw 2>&1 | tail -$((ov+1))
done
rm -f gpurun_out/conv_overlap_ab2.jsonl
for ov in 4 0 8 0; do
 CPLXK_CONV_OVERLAP=$ov timeout 120 python tools/conv_bench.py --plain --fp32-nchw >> gpurun_out/conv_overlap_ab2.jsonl 2>gpurun_out/conv_overlap_ab2.err
done
cat gpurun_out/conv_overlap_ab2.jsonl
