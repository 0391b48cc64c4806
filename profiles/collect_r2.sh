# Round-2 evidence, run on the GPU box from the repo root (gpurun): ncu --set full captures of the kernels changed this
# round at the bench configuration, per-launch duration lists of a whole step of both workloads.
# Each ncu command runs only after the same program has exited 0 without ncu.
set -x
mkdir -p gpurun_out
python tools/vfab.py adapted_b200/csrc/libadapted_b200.so rna002 100000 > gpurun_out/vfab_rna002.log 2>&1 || exit 1
python tools/vfab.py adapted_b200/csrc/libadapted_b200.so rna004 250000 > gpurun_out/vfab_rna004.log 2>&1 || exit 1
timeout 600 ncu --set full --clock-control none --import-source on -k "regex:validate_hist|series_median" -c 2 -o gpurun_out/r2_final_rna002 python tools/vfab.py adapted_b200/csrc/libadapted_b200.so rna002 100000 > gpurun_out/ncu_final_rna002.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k "regex:validate_hist|series_median|cnn_prep_warp|validate_cand" -c 4 -o gpurun_out/r2_final_rna004 python tools/vfab.py adapted_b200/csrc/libadapted_b200.so rna004 250000 > gpurun_out/ncu_final_rna004.log 2>&1
K='regex:^(gs|llr|mvs|series|svb|validate|void validate|merge|cnn_|void cnn_|start_peak)'
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k "$K" -c 400 --csv --log-file gpurun_out/r2_launches_rna002.csv python bench.py --chemistry rna002 --steps 2 --warmup 3 --no-cpu-baseline --no-secondary --file-reads 0 --profile-steps-only > gpurun_out/ncu_r2_lf2.log 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k "$K" -c 900 --csv --log-file gpurun_out/r2_launches_rna004.csv python bench.py --chemistry rna004 --reads 100000 --steps 1 --warmup 3 --no-cpu-baseline --no-secondary --file-reads 0 --profile-steps-only > gpurun_out/ncu_r2_lf4.log 2>&1
echo done
