"""Summarise an ncu report: key raw metrics + instructions / stall samples per function and line."""
import collections, csv, re, subprocess, sys
rep = sys.argv[1]
src_path = sys.argv[2] if len(sys.argv) > 2 else '/root/repo/chapterhouseqe_b200/csrc/device_code.cuh'
rows_per_launch = float(sys.argv[3]) if len(sys.argv) > 3 else 4194304.0
raw = subprocess.run(['ncu','-i',rep,'--page','raw','--csv'],capture_output=True,text=True).stdout
rows = list(csv.reader(raw.splitlines()))
h = rows[0]
want = ['gpu__time_duration.sum','dram__bytes_read.sum','dram__bytes_write.sum','dram__throughput.avg.pct_of_peak_sustained_elapsed',
 'sm__warps_active.avg.pct_of_peak_sustained_active','launch__registers_per_thread','launch__occupancy_limit_registers',
 'launch__occupancy_limit_shared_mem','launch__occupancy_limit_warps','sm__throughput.avg.pct_of_peak_sustained_elapsed',
 'l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum','l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum',
 'l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum','l1tex__t_requests_pipe_lsu_mem_global_op_st.sum','smsp__inst_executed.sum',
 'smsp__issue_active.avg.pct_of_peak_sustained_active','lts__t_sector_hit_rate.pct','l1tex__t_sector_hit_rate.pct',
 'smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio','smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio',
 'smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio','smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio',
 'smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio','smsp__average_warps_issue_stalled_wait_per_issue_active.ratio',
 'smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio','smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio',
 'smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio','smsp__average_warps_issue_stalled_membar_per_issue_active.ratio',
 'smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio','smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio',
 'launch__grid_size','launch__waves_per_multiprocessor','dram__bytes.sum.per_second']
for w in want:
    if w in h:
        i = h.index(w); print(f"{w:95s}", [r[i] for r in rows[1:]])
mix = subprocess.run(['ncu','-i',rep,'--page','source','--csv','--print-source','cuda,sass'],capture_output=True,text=True).stdout
rows = list(csv.reader(mix.splitlines()))
src = open(src_path).read().split('\n')
fn_at = {}; cur = '?'
for i, l in enumerate(src, 1):
    m = re.search(r'(?:__device__|__global__)[^;]*?\b(\w+)\s*\(', l)
    if m and not l.strip().startswith('//'): cur = m.group(1)
    fn_at[i] = cur
def toi(x):
    try: return int(x)
    except: return 0
cur_file = None; hdr = None; lines = []; nlaunch = 0
for r in rows:
    if r and r[0] == 'File Path': cur_file = r[1].split('/')[-1]; continue
    if r and r[0] == 'Function Name': continue
    if r and r[0] == 'Kernel Name': nlaunch += 1; continue
    if r and r[0] == 'Line No': hdr = r; ie = hdr.index('Instructions Executed'); ss = hdr.index('# Samples'); continue
    if hdr and r and r[0] != '' and len(r) > ie:
        try: ln = int(r[0])
        except: continue
        lines.append((cur_file, ln, r[1].strip()[:100], toi(r[ie]), toi(r[ss])))
nl = max(1, len([1 for r in rows if r and r[0]=='File Path' and (r[1].endswith('kernels.cu') or r[1].endswith('device_code.cuh'))]))
agg = collections.Counter(); sm = collections.Counter()
for f, ln, _, ins, s in lines:
    key = fn_at.get(ln, '?') if f in ('kernels.cu', 'device_code.cuh') else f
    agg[key] += ins; sm[key] += s
tot = sum(agg.values()); tots = sum(sm.values())
print(f"total warp-instr (all captured launches) {tot}, launches~{nl}; instr/row = {tot/nl/rows_per_launch*32:.1f}")
for k, v in agg.most_common(25):
    print(f"{100*v/tot:5.1f}% inst  {100*sm[k]/max(tots,1):5.1f}% samples  {v/nl/rows_per_launch*32:7.1f} instr/row  {k}")
print("--- top lines by stall samples")
seen = set()
for l in sorted(lines, key=lambda x: -x[4]):
    k = (l[0], l[1])
    if k in seen: continue
    seen.add(k)
    print(f"{100*l[4]/max(tots,1)*1:5.1f}% smp {100*l[3]/tot:5.1f}% inst  {l[0]}:{l[1]}  {l[2]}")
    if len(seen) >= 30: break
