# Round-2 ncu captures (run on the GPU box through gpurun). Summaries are produced ON the box (the .ncu-rep files of the
# multi-kernel captures exceed gpurun's 64 MiB return limit) and land in gpurun_out/*.md; copy what should be judged to
# profiles/.
set -u
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on --profile-from-start off -o gpurun_out/ops_full_r2 -f python tools/ops_ncu_target.py > gpurun_out/ncu_ops.log 2>&1
python tools/ncu_raw_summary.py gpurun_out/ops_full_r2.ncu-rep > gpurun_out/ops_full_r2.md 2>gpurun_out/ops_summary.err
rm -f gpurun_out/ops_full_r2.ncu-rep
ncu --profile-from-start off --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file gpurun_out/step_launches_r2.csv python tools/step_ncu_target.py > gpurun_out/ncu_step.log 2>&1
python tools/summarize_ncu_csv.py gpurun_out/step_launches_r2.csv > gpurun_out/step_launches_r2.md 2>gpurun_out/step_summary.err
python tools/ncu_traffic.py gpurun_out/step_launches_r2.csv > gpurun_out/ncu_traffic_r2.json 2>>gpurun_out/step_summary.err
rm -f gpurun_out/step_launches_r2.csv
ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:conv_halo_tp -c 6 -o gpurun_out/conv_halo_tp_full_r2 -f python tools/step_ncu_target.py > gpurun_out/ncu_tp.log 2>&1
python tools/ncu_raw_summary.py gpurun_out/conv_halo_tp_full_r2.ncu-rep > gpurun_out/conv_halo_tp_full_r2.md 2>gpurun_out/tp_summary.err
rm -f gpurun_out/conv_halo_tp_full_r2.ncu-rep
ls -la gpurun_out | tail -12
head -30 gpurun_out/ops_full_r2.md | cut -c1-400
