timeout 200 ncu --set full --clock-control none -k regex:conv_tc_pair -s 2 -c 1 -o /tmp/p_real -f python tools/prof_real_conv.py 4 bf16 > /tmp/p.log 2>&1
ncu -i /tmp/p_real.ncu-rep --page raw --csv > gpurun_out/prof_convpair_real_bf16_r2.raw.csv 2>/dev/null
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file gpurun_out/conv_real_launches_r2.csv python tools/prof_real_conv.py 4 bf16 > /dev/null 2>&1
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/conv_launches_r2b.csv python tools/prof_conv.py 3 f32 nchw > /dev/null 2>&1
ls -la gpurun_out/*.csv | tail -4
