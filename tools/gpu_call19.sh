set -x
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
timeout 600 python bench.py --steps 200 --warmup 3 > gpurun_out/bench_r1_final4.json 2> gpurun_out/bench_r1_final4.err; cut -c1-200 gpurun_out/bench_r1_final4.json; tail -2 gpurun_out/bench_r1_final4.err
timeout 300 python tools/train_step_profile.py 50 64 fused > gpurun_out/train_profile_r1_final2.txt 2>&1; head -12 gpurun_out/train_profile_r1_final2.txt
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file gpurun_out/step_launches_r1b.csv python tools/step_ncu_target.py 64 > gpurun_out/ncu_step_b.log 2>&1; tail -2 gpurun_out/ncu_step_b.log
timeout 600 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:"conv_gemm_kernel|gn_fused|dense_rows|time_embedding" -s 30 -c 24 -o gpurun_out/tap_kernels_full_r1b -f python tools/step_ncu_target.py 64 > gpurun_out/ncu_tap_b.log 2>&1; tail -2 gpurun_out/ncu_tap_b.log
timeout 600 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:conv_halo_tp -s 4 -c 2 -o gpurun_out/conv_halo_tp_full_r1b -f python tools/step_ncu_target.py 64 > gpurun_out/ncu_tp_b.log 2>&1; tail -2 gpurun_out/ncu_tp_b.log
