"""Text summary of an ncu report (run here, no GPU needed): per kernel the headline metrics of the raw page, the instruction
mix and the hottest stall sites of the source page.   python scripts/ncu_summary.py gpurun_out/x.ncu-rep > profiles/x.txt"""
import collections, csv, io, re, subprocess, sys

rep = sys.argv[1]
raw = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr = rows[0]
KEYS = ('gpu__time_duration.sum', 'launch__grid_size', 'launch__block_size', 'launch__registers_per_thread', 'launch__occupancy_limit_registers',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'smsp__issue_active.avg.pct_of_peak_sustained_active', 'smsp__inst_executed.sum',
        'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum',
        'smsp__sass_inst_executed_op_tmem_ldt.sum', 'smsp__sass_inst_executed_op_tmem_stt.sum')
names = []
for r in rows[2:]:
    d = dict(zip(hdr, r))
    names.append(d['Kernel Name'])
    print('=' * 100)
    print(d['Kernel Name'])
    for k in KEYS:
        if k in d and d[k] not in ('', 'n/a'):
            print(f'  {k:78s} {d[k]}')
    st = sorted(((float(v), k) for k, v in d.items() if k.startswith('smsp__average_warps_issue_stalled') and k.endswith('per_issue_active.ratio') and v not in ('', 'n/a')), reverse=True)
    print('  stall reasons (warps stalled per issue-active cycle): ' + ', '.join(f"{k.split('stalled_')[1].split('_per_')[0]} {v:.2f}" for v, k in st[:7]))
for name in dict.fromkeys(names):
    short = re.sub(r'\(.*', '', name).split('::')[-1]
    src = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv', '--kernel-name', 'regex:' + re.escape(short.split('<')[0])], capture_output=True, text=True).stdout
    rs = list(csv.reader(io.StringIO(src)))
    if len(rs) < 3:
        continue
    h = rs[1]; ix = {c: i for i, c in enumerate(h)}
    seen, data = set(), []
    for r in rs[2:]:
        if len(r) < len(h) - 2 or not r[ix['# Samples']].isdigit() or r[ix['Address']] in seen:
            continue
        seen.add(r[ix['Address']]); data.append(r)
    tot = sum(int(r[ix['# Samples']]) for r in data) or 1
    ti = sum(int(r[ix['Instructions Executed']]) for r in data) or 1
    print('=' * 100)
    print(f'{short}: source page (first launch), {ti} warp instructions, {tot} stall samples')
    cnt, smp = collections.Counter(), collections.Counter()
    for r in data:
        m = re.match(r'(@!?U?P\d+\s+)?([A-Z0-9_.]+)', r[ix['Source']].strip())
        op = m.group(2).split('.')[0] if m else '?'
        cnt[op] += int(r[ix['Instructions Executed']]); smp[op] += int(r[ix['# Samples']])
    print('  instruction mix: ' + ', '.join(f'{op} {100 * c / ti:.1f}%' for op, c in cnt.most_common(14)))
    print('  hottest stall sites:')
    for r in sorted(data, key=lambda r: -int(r[ix['# Samples']]))[:12]:
        s2 = sorted(((int(r[ix[k]]), k) for k in h if k.startswith('stall_') and 'Not' not in k and r[ix[k]].isdigit() and int(r[ix[k]]) > 0), reverse=True)[:2]
        print(f"    {100 * int(r[ix['# Samples']]) / tot:5.1f}%  {r[ix['Source']].strip()[:64]:64s} {', '.join(f'{k[6:]} {v}' for v, k in s2)}")
