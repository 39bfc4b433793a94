set -x
# full captures: transposed conv (160px/96ch and 80px/192ch), per-tap conv, GroupNorm apply, FIR
timeout 900 ncu --set full --clock-control none --import-source on -k regex:conv_halo_t -s 2 -c 1 -o gpurun_out/conv_halo_t_full_r1 python tools/conv_ncu_target.py 64 > gpurun_out/ncu2.log 2>&1; tail -2 gpurun_out/ncu2.log
timeout 900 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:"gn_apply|fir_nhwc|gn_chan_stats|conv_gemm_kernel" -c 40 -o gpurun_out/forward_mix_full_r1 python tools/forward_ncu_target.py 64 > gpurun_out/ncu3.log 2>&1; tail -2 gpurun_out/ncu3.log
# phase timestamps (box-local rebuild with the timestamp macro)
CSD_NVCC_EXTRA=-DCSD_ENABLE_PHASE_TIMESTAMPS python -m conditional_score_diffusion_b200.build --force > /dev/null && timeout 300 python tools/conv_phase_timing.py > gpurun_out/conv_phase_timing_r1.txt 2>&1; cat gpurun_out/conv_phase_timing_r1.txt
ls -la gpurun_out
