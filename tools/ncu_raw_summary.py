"""Summarise an .ncu-rep (ncu --set full) into a small markdown table: per launch duration, DRAM bytes, DRAM and
tensor-pipe utilisation, registers, shared memory.   python tools/ncu_raw_summary.py file.ncu-rep > profiles/x.md"""
import csv
import io
import subprocess
import sys

path = sys.argv[1]
out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr, units, data = rows[0], rows[1], rows[2:]
ix = {h: i for i, h in enumerate(hdr)}
want = [("gpu__time_duration.sum", "time"), ("dram__bytes_read.sum", "DRAM rd"), ("dram__bytes_write.sum", "DRAM wr"),
        ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "DRAM %"),
        ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor pipe % (active)"),
        ("sm__inst_executed_pipe_tensor.sum", "tensor inst"),
        ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM %"),
        ("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "smem wavefronts"),
        ("lts__t_bytes.sum", "L2 bytes"), ("launch__registers_per_thread", "regs"),
        ("launch__shared_mem_per_block_dynamic", "dyn smem"), ("launch__grid_size", "grid"),
        ("launch__block_size", "block"), ("sm__warps_active.avg.pct_of_peak_sustained_active", "occupancy %")]
cols = [(k, n) for k, n in want if k in ix]
tens = [h for h in hdr if "pipe_tensor" in h and "pct" in h]
print(f"source: {path}\n")
print("| # | kernel | " + " | ".join(f"{n} [{units[ix[k]]}]" for k, n in cols) + " |")
print("|---|---|" + "---|" * len(cols))
for n, r in enumerate(data):
    name = r[ix["Kernel Name"]].split("(")[0].replace("csd::", "").replace("void ", "")
    print(f"| {n} | {name} | " + " | ".join(r[ix[k]] for k, _ in cols) + " |")
if tens:
    print("\ntensor-pipe metrics present: " + ", ".join(f"{h}={data[0][ix[h]]}" for h in tens[:6]))
