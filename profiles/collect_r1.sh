set -x
K='regex:^(gs|llr|mvs|validate|void validate|merge|cnn_|void cnn_|start_peak)'
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -k "$K" -c 400 --csv --log-file gpurun_out/launches_final_rna002.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --profile-steps-only > gpurun_out/ncu_lf2.log 2>&1; echo rc=$?
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -k "$K" -c 600 --csv --log-file gpurun_out/launches_final_rna004.csv python bench.py --chemistry rna004 --steps 1 --warmup 3 --no-cpu-baseline --profile-steps-only > gpurun_out/ncu_lf4.log 2>&1; echo rc=$?
timeout 200 ncu --set full --clock-control none --import-source on -k regex:^mvs_series_kernel -s 3 -c 1 -o gpurun_out/full_mvs_series_kernel_v2 -f python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_f1.log 2>&1; echo rc=$?
timeout 250 ncu --set full --clock-control none --import-source on -k regex:cnn_conv64_tc_kernel -s 6 -c 2 -o gpurun_out/full_cnn_conv64_tc_final -f python bench.py --chemistry rna004 --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_f2.log 2>&1; echo rc=$?
timeout 250 ncu --set full --clock-control none --import-source on -k regex:^validate_kernel -s 3 -c 1 -o gpurun_out/full_validate_handover_final -f python bench.py --chemistry rna004 --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_f3.log 2>&1; echo rc=$?
