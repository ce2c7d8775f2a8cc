timeout 400 python -m pytest tests/test_gpu_conv_rg.py tests/test_gpu_conv.py tests/test_gpu_backward.py tests/test_gpu_rows_f.py tests/test_gpu_graph.py tests/test_gpu_edge_cases.py -m gpu -q -x > gpurun_out/pytest_flat.log 2>&1; echo "rc=$?"; tail -4 gpurun_out/pytest_flat.log
rm -f gpurun_out/convvd_flat_ab.jsonl
for f in 1 0 1 0; do
 echo "{\"CPLXK_COMBINE_FLAT\": $f}" >> gpurun_out/convvd_flat_ab.jsonl
 CPLXK_COMBINE_FLAT=$f timeout 200 python tools/convvd_probe.py 2>&1 | head -2 >> gpurun_out/convvd_flat_ab.jsonl
done
cat gpurun_out/convvd_flat_ab.jsonl
