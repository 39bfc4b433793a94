set -x
cp conditional_score_diffusion_b200/libcsd_b200.so /tmp/libcsd_keep.so
CSD_NVCC_EXTRA=-DCSD_ENABLE_PHASE_TIMESTAMPS python -m conditional_score_diffusion_b200.build > /dev/null 2>&1
CSD_NVCC_EXTRA=-DCSD_ENABLE_PHASE_TIMESTAMPS timeout 200 python tools/tap_phase_timing.py > gpurun_out/tap_phase_timing2.txt 2>&1; cat gpurun_out/tap_phase_timing2.txt
