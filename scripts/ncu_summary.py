"""Summarise an .ncu-rep here (no GPU): key raw metrics per launch and the hottest source lines.
usage: python scripts/ncu_summary.py gpurun_out/prof.ncu-rep [n_lines]"""
import csv, subprocess, sys, io
rep = sys.argv[1]
nl = int(sys.argv[2]) if len(sys.argv) > 2 else 30
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
keys = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'launch__registers_per_thread', 'launch__shared_mem_per_block_dynamic', 'smsp__inst_executed.sum', 'sm__cycles_elapsed.avg',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum',
        'lts__t_sector_hit_rate.pct', 'smsp__inst_executed_op_local_ld.sum', 'smsp__inst_executed_op_local_st.sum']
stall = [h for h in hdr if h.startswith('smsp__average_warps_issue_stalled_') and h.endswith('_per_issue_active.ratio')]
for r in rows[2:]:
    print(r[hdr.index('Kernel Name')][:60], r[hdr.index('Block Size')], r[hdr.index('Grid Size')])
    for k in keys:
        if k in hdr:
            i = hdr.index(k); print('   %s = %s %s' % (k, r[i], units[i]))
    st = sorted(((float(r[hdr.index(h)]), h) for h in stall), reverse=True)[:7]
    print('   top stalls (warps per issue):', ', '.join('%s %.2f' % (h[len('smsp__average_warps_issue_stalled_'):-len('_per_issue_active.ratio')], v) for v, h in st))
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
cur = None; h = None; out = []; first_kernel = None
for r in rows:
    if len(r) >= 2 and r[0] == 'Function Name':
        if first_kernel is None: first_kernel = r[1]
        elif r[1] != first_kernel: pass
    if len(r) >= 2 and r[0] == 'File Path': cur = r[1].split('/')[-1]; continue
    if len(r) > 3 and r[0] == 'Line No': h = r; continue
    if h and len(r) == len(h) and r[0] != '':
        d = dict(zip(h, r))
        try: out.append((int(d['Instructions Executed']), int(d['# Samples']), cur, r[0], r[1].strip()[:100]))
        except Exception: pass
ti = sum(o[0] for o in out) or 1; ts = sum(o[1] for o in out) or 1
print('source lines by stall samples (inst%% / samples%%), total inst %d samples %d' % (ti, ts))
for o in sorted(out, key=lambda o: -o[1])[:nl]:
    print('  %5.1f%% %5.1f%%  %s:%s  %s' % (100 * o[0] / ti, 100 * o[1] / ts, o[2], o[3], o[4]))
