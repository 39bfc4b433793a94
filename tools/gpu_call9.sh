set -x
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
timeout 900 ncu --profile-from-start off --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file gpurun_out/step_launches_r1.csv python tools/step_ncu_target.py 64 > gpurun_out/ncu_step.log 2>&1; tail -2 gpurun_out/ncu_step.log
timeout 900 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:conv_halo_tp -s 4 -c 2 -o gpurun_out/conv_halo_tp_full_r1 -f python tools/step_ncu_target.py 64 > gpurun_out/ncu_tp.log 2>&1; tail -2 gpurun_out/ncu_tp.log
timeout 900 ncu --profile-from-start off --set full --clock-control none -k regex:"gn_apply|fir_tma|gn_chan_stats|conv_gemm_kernel" -c 30 -o gpurun_out/hbm_kernels_full_r1 -f python tools/step_ncu_target.py 64 > gpurun_out/ncu_hbm.log 2>&1; tail -2 gpurun_out/ncu_hbm.log
timeout 600 python bench.py --steps 200 --warmup 3 > gpurun_out/bench_r1_final.json 2> gpurun_out/bench_r1_final.err; cut -c1-200 gpurun_out/bench_r1_final.json; tail -3 gpurun_out/bench_r1_final.err
