#!/usr/bin/env python
"""Summaries of ncu exports (read on the CPU box):  launches <csv>  |  raw <ncu-rep>  |  sass <ncu-rep> <kernel> [n]"""
import collections, csv, subprocess, sys

def launches(path):
    lines = [l for l in open(path) if not l.startswith('==')]
    agg = collections.OrderedDict()
    for x in csv.DictReader(lines):
        k = x['Kernel Name'].split('(')[0][-40:]
        a = agg.setdefault(k, [0, 0.0]); a[0] += 1; a[1] += float(x['Metric Value'].replace(',', ''))
    tot = sum(a[1] for a in agg.values())
    print('total %.1f us over %d launches' % (tot / 1e3, sum(a[0] for a in agg.values())))
    for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print('%-42s n=%4d total=%9.1f us share=%.3f avg=%8.1f us' % (k, a[0], a[1] / 1e3, a[1] / tot, a[1] / a[0] / 1e3))

WANT = ['gpu__time_duration.sum', 'sm__cycles_elapsed.max', 'launch__registers_per_thread',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed', 'sm__inst_executed_pipe_tensor.sum',
        'dram__bytes_read.sum', 'dram__bytes_write.sum', 'lts__t_bytes.sum', 'lts__t_sectors_srcunit_tex_op_read.sum',
        'smsp__inst_executed.sum', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum',
        'smsp__inst_executed_op_local_ld.sum', 'smsp__inst_executed_op_local_st.sum', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed']

def raw(rep):
    out = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    stall = [h for h in hdr if 'smsp__average_warps_issue_stalled' in h and h.endswith('_per_issue_active.ratio')]
    for r in rows[2:]:
        print('---', r[idx['Kernel Name']], 'grid', r[idx['Grid Size']] if 'Grid Size' in idx else '')
        for w in WANT:
            if w in idx: print('   %-70s %s %s' % (w, r[idx[w]], units[idx[w]]))
        s = sorted(((float(r[idx[h]] or 0), h) for h in stall), reverse=True)[:6]
        print('   stalls/issue:', ', '.join('%s %.2f' % (h.replace('smsp__average_warps_issue_stalled_', '').replace('_per_issue_active.ratio', ''), v) for v, h in s))

def sass(rep, kernel, n=40):
    out = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv', '--kernel-name', kernel], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    inst, cur, hdr = [], None, None
    for r in rows:
        if r and r[0] == 'Kernel Name': cur = []; inst.append(cur); continue
        if r and r[0] == 'Address': hdr = r; continue
        if cur is not None and hdr and len(r) == len(hdr): cur.append(r)
    data = inst[0]; si = hdr.index('# Samples')
    tot = sum(int(r[si]) for r in data)
    print('total samples', tot, 'instructions', len(data))
    top = sorted(range(len(data)), key=lambda i: -int(data[i][si]))[:n]
    for i in sorted(top): print('%6d %6d %5.1f%%  %s' % (i, int(data[i][si]), 100.0 * int(data[i][si]) / tot, data[i][1][:110]))

if __name__ == '__main__':
    cmd = sys.argv[1]
    if cmd == 'launches': launches(sys.argv[2])
    elif cmd == 'raw': raw(sys.argv[2])
    elif cmd == 'sass': sass(sys.argv[2], sys.argv[3], int(sys.argv[4]) if len(sys.argv) > 4 else 40)

def lines(rep, kernel, n=40):
    """per CUDA source line: executed warp instructions and stall samples (needs -lineinfo + --import-source)"""
    out = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv', '--print-source', 'cuda,sass', '--kernel-name', kernel],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    agg = {}; cur_file = ''; hdr = None
    for r in rows:
        if not r: continue
        if r[0] == 'File Path': cur_file = r[1].split('/')[-1]; continue
        if r[0] == 'Line No': hdr = r; continue
        if hdr is None or len(r) != len(hdr): continue
        if r[0] == '' : continue          # sass rows have empty line number; cuda rows carry aggregated metrics
        try:
            ln = int(r[0]); samples = int(r[hdr.index('# Samples')] or 0); inst = int(r[hdr.index('Instructions Executed')] or 0)
        except ValueError: continue
        a = agg.setdefault((cur_file, ln), [r[1].strip()[:90], 0, 0]); a[1] += samples; a[2] += inst
    ts = sum(a[1] for a in agg.values()); ti = sum(a[2] for a in agg.values())
    print('total samples %d, executed warp-instructions %d' % (ts, ti))
    print('--- by executed instructions')
    for (f, ln), a in sorted(agg.items(), key=lambda kv: -kv[1][2])[:n]:
        print('%5.1f%% inst %5.1f%% smp  %s:%d  %s' % (100.0 * a[2] / max(ti, 1), 100.0 * a[1] / max(ts, 1), f, ln, a[0]))
    print('--- by stall samples')
    for (f, ln), a in sorted(agg.items(), key=lambda kv: -kv[1][1])[:n]:
        print('%5.1f%% smp %5.1f%% inst  %s:%d  %s' % (100.0 * a[1] / max(ts, 1), 100.0 * a[2] / max(ti, 1), f, ln, a[0]))

if __name__ == '__main__' and sys.argv[1] == 'lines':
    lines(sys.argv[2], sys.argv[3], int(sys.argv[4]) if len(sys.argv) > 4 else 40)
