"""Summarise an .ncu-rep (read here, no GPU): key raw metrics per kernel + hottest source lines.
usage: python tools/ncu_summary.py gpurun_out/prof.ncu-rep [n_lines]"""
import csv, io, subprocess, sys
rep = sys.argv[1]; nl = int(sys.argv[2]) if len(sys.argv) > 2 else 25
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr = rows[0]
want = ['Kernel Name', 'gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'lts__t_bytes.sum', 'l1tex__t_bytes.sum', 'sm__warps_active.avg.pct_of_peak_sustained_active', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'smsp__inst_executed.sum', 'launch__registers_per_thread', 'launch__grid_size', 'launch__block_size',
        'launch__occupancy_limit_registers', 'launch__occupancy_limit_shared_mem', 'launch__waves_per_multiprocessor',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'sm__inst_executed_pipe_lsu.sum', 'sm__inst_executed_pipe_alu.sum', 'smsp__thread_inst_executed_per_inst_executed.ratio']
for r in rows[2:]:
    print("== kernel")
    for w in want:
        if w in hdr:
            print("  %-70s %s %s" % (w, r[hdr.index(w)], rows[1][hdr.index(w)]))
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
def I(v):
    try: return int(float(v))
    except: return 0
cur = None; h = None; lines = {}
for r in csv.reader(io.StringIO(src)):
    if len(r) == 2 and r[0] == "File Path": cur = r[1].split("/")[-1]; continue
    if r and r[0] == "Line No": h = r; continue
    if h is None or len(r) != len(h) or not r[0]: continue          # SASS rows have an empty line number
    key = (cur, r[0])
    si, ii = h.index('# Samples'), h.index('Instructions Executed')
    e = lines.setdefault(key, [0, 0, r[1], {}])
    e[0] += I(r[si]); e[1] += I(r[ii])
    for k, name in enumerate(h):
        if name.startswith("stall_") and "Not Issued" not in name and I(r[k]): e[3][name[6:]] = e[3].get(name[6:], 0) + I(r[k])
tot_s = sum(e[0] for e in lines.values()) or 1; tot_i = sum(e[1] for e in lines.values()) or 1
print("== source lines: samples=%d warp-instructions=%d" % (tot_s, tot_i))
for (f, ln), e in sorted(lines.items(), key=lambda kv: -kv[1][0])[:nl]:
    top = sorted(e[3].items(), key=lambda kv: -kv[1])[:2]
    print("  %s:%-4s smp %5.1f%% inst %5.1f%% %-28s %s" % (f[:14], ln, 100.0 * e[0] / tot_s, 100.0 * e[1] / tot_i, ",".join("%s=%d" % t for t in top), e[2].strip()[:90]))
