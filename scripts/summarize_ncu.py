"""Summarise an ncu report (.ncu-rep) into profiles/: key raw metrics as JSON + the hottest SASS
lines of the source page. Usage: python scripts/summarize_ncu.py gpurun_out/prof.ncu-rep profiles/r01_k1 [rows nq]"""
import csv
import io
import json
import re
import subprocess
import sys

rep, out = sys.argv[1], sys.argv[2]
rows_nq = (int(sys.argv[3]), int(sys.argv[4])) if len(sys.argv) > 4 else None
KEYS = re.compile(r"^(gpu__time_duration\.sum|dram__bytes_read\.sum|dram__bytes_write\.sum|"
                  r"dram__bytes_read\.sum\.per_second|gpu__dram_throughput\.avg\.pct_of_peak_sustained_elapsed|"
                  r"sm__cycles_elapsed\.avg\.per_second|sm__cycles_elapsed\.max|launch__grid_size|launch__block_size|"
                  r"launch__registers_per_thread|launch__shared_mem_per_block_dynamic|launch__waves_per_multiprocessor|"
                  r"lts__t_sector_hit_rate\.pct|lts__throughput\.avg\.pct_of_peak_sustained_elapsed|"
                  r"l1tex__m_xbar2l1tex_read_bytes\.sum|l1tex__m_xbar2l1tex_read_bytes\.sum\.per_second|"
                  r"sm__throughput\.avg\.pct_of_peak_sustained_elapsed|smsp__issue_active\.avg\.pct_of_peak_sustained_active|"
                  r"sm__warps_active\.avg\.pct_of_peak_sustained_active|smsp__inst_executed\.sum|"
                  r"smsp__sass_inst_executed_op_utcmma\.sum|smsp__sass_inst_executed_op_tma_ld\.sum|"
                  r"smsp__sass_inst_executed_op_tmem_ldt\.sum|"
                  r"sm__ops_path_tensor_op_utchmma_src_bf16_dst_fp32_sparsity_off\.(sum|avg\.pct_of_peak_sustained_elapsed)|"
                  r"smsp__mem_tensor_reads_op_utcmma_matrix_c\.sum\.pct_of_peak_sustained_elapsed)$")
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
r = list(csv.reader(io.StringIO(raw)))
hdr, units, data = r[0], r[1], r[2:]
summary = []
for row in data:
    d = {"kernel": row[hdr.index("Kernel Name")]}
    for i, h in enumerate(hdr):
        if KEYS.match(h) or h.endswith("sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed"):
            d[h.split("TriageCompute.")[-1] + (f" [{units[i]}]" if units[i] else "")] = row[i]
    summary.append(d)
json.dump(summary, open(out + "_metrics.json", "w"), indent=1)
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
r = list(csv.reader(io.StringIO(src)))
hdr = r[1]
ia, isrc, ismp, iex = hdr.index("Address"), hdr.index("Source"), hdr.index("# Samples"), hdr.index("Instructions Executed")
body = [x for x in r[2:] if len(x) > ismp and x[ismp].isdigit()]
tot = sum(int(x[ismp]) for x in body)
with open(out + "_hot_sass.txt", "w") as f:
    f.write(f"# {r[0][1]}\n# total warp-stall samples {tot}; top 30 SASS instructions\n# addr  samples  executed  instruction\n")
    for x in sorted(body, key=lambda x: -int(x[ismp]))[:30]:
        f.write(f"{x[ia][-5:]} {x[ismp]:>8} {x[iex]:>10}  {x[isrc]}\n")
if rows_nq:
    k = summary[0]
    rd = [v for kk, v in k.items() if kk.startswith("dram__bytes_read.sum [")][0], [kk for kk in k if kk.startswith("dram__bytes_read.sum [")][0]
    wr = [v for kk, v in k.items() if kk.startswith("dram__bytes_write.sum [")][0], [kk for kk in k if kk.startswith("dram__bytes_write.sum [")][0]
    mul = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1}
    tb = float(rd[0]) * mul[rd[1].split("[")[1][:-1]] + float(wr[0]) * mul[wr[1].split("[")[1][:-1]]
    json.dump({"rows": rows_nq[0], "nq": rows_nq[1], "dram_bytes_per_launch": tb, "source": rep.split("/")[-1] + " (ncu --set full, one K1 launch)"},
              open("profiles/k1_traffic.json", "w"))
print(json.dumps(summary[0], indent=1))
