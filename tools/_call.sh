timeout 600 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_final2.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_final2.log
timeout 300 python tools/config_bench.py > gpurun_out/configs_r2b.log 2>&1; echo "cfg rc=$?"; grep -c config gpurun_out/configs_r2b.log
timeout 400 python bench.py > gpurun_out/bench_final2.json 2> gpurun_out/bench_final2.err; echo "bench rc=$?"; cut -c1-400 gpurun_out/bench_final2.json
for v in "f32 nchw" "bf16 nhwc"; do
 set -- $v
 timeout 200 ncu --set full --clock-control none -k regex:conv_tc_pair -s 2 -c 1 -o /tmp/p_$1 -f python tools/prof_conv.py 4 $1 $2 > /tmp/p.log 2>&1
 ncu -i /tmp/p_$1.ncu-rep --page raw --csv > gpurun_out/prof_convpair_row_$1_r2.raw.csv 2>/dev/null
done
ls -la gpurun_out/*.raw.csv | tail -3
