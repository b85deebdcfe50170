#!/usr/bin/env python
"""Summarise ncu outputs brought back in gpurun_out/ into profiles/ (tracked).

  python tools/ncu_summary.py launches gpurun_out/r01_launches.csv profiles/r01_launches_summary.md
  python tools/ncu_summary.py rep gpurun_out/r01_k_shade.ncu-rep profiles/r01_k_shade_ncu.md [profiles/k_shade_traffic.json]
"""
import csv
import io
import json
import re
import os
import subprocess
import sys
from collections import defaultdict

EXTRA = {'lts__t_sectors_srcunit_tex_op_read.sum.per_second', 'l1tex__m_xbar2l1tex_read_bytes.sum', 'l1tex__m_xbar2l1tex_read_bytes.sum.per_second',
         'lts__t_sector_hit_rate.pct', 'sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active', 'smsp__inst_executed.sum',
         'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active'}
KEYS = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active', 'sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread', 'launch__grid_size', 'launch__block_size',
        'launch__shared_mem_per_block_dynamic', 'smsp__cycles_active.avg', 'sm__cycles_elapsed.max', 'l1tex__data_bank_conflicts_pipe_lsu.sum',
        'lts__t_bytes.sum', 'sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active', 'smsp__inst_executed.sum',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_uniform.sum', 'sm__pipe_tensor_op_hmma_cycles_active.avg.pct_of_peak_sustained_active']


def launches(src, dst):
    rows = []
    with open(src) as f:
        txt = f.read()
    start = txt.find('"ID"')
    rd = csv.DictReader(io.StringIO(txt[start:]))
    for r in rd:
        if r.get('Metric Name') == 'gpu__time_duration.sum':
            v = float(r['Metric Value'].replace(',', ''))
            unit = r.get('Metric Unit', 'ns')
            scale = {'ns': 1e-6, 'us': 1e-3, 'usecond': 1e-3, 'nsecond': 1e-6, 'ms': 1.0, 'msecond': 1.0, 's': 1e3, 'second': 1e3}.get(unit, 1e-6)
            rows.append((re.sub(r'\(.*', '', r['Kernel Name']).strip(), v * scale))
    agg = defaultdict(lambda: [0, 0.0])
    for k, ms in rows:
        agg[k][0] += 1
        agg[k][1] += ms
    tot = sum(v[1] for v in agg.values())
    with open(dst, 'w') as f:
        f.write(f'# ncu launch list summary ({src})\n\nAll launches of the command, device time per kernel (cold-cache, serialised: compare SHARES).\n\n')
        f.write(f'total launches {len(rows)}, total device time {tot:.2f} ms\n\n| kernel | launches | total ms | share |\n|---|---:|---:|---:|\n')
        for k, (n, ms) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write(f'| `{k}` | {n} | {ms:.3f} | {100 * ms / tot:.1f} % |\n')
    print(open(dst).read())


def csvpage(src, dst, traffic_json=None):
    """`ncu -i x.ncu-rep --page raw --csv` output (made on the GPU box) -> markdown table per kernel + optional traffic JSON."""
    rd = list(csv.reader(open(src)))
    hdr, units, data = rd[0], rd[1], rd[2:]
    traffic = {'_source': f'ncu --set full --clock-control none, one launch per kernel ({src})'}
    with open(dst, 'w') as f:
        f.write(f'# ncu --set full summary ({src})\n\nOne launch per kernel inside a steady-state bench.py frame; numbers under ncu are never bench values.\n\n')
        for row in data:
            d = dict(zip(hdr, row))
            u = dict(zip(hdr, units))
            name = re.sub(r'\(.*', '', d.get('Kernel Name', '?')).strip()
            f.write(f"## {name}  (ID {d.get('ID')})\n\n| metric | value | unit |\n|---|---:|---|\n")
            for k in hdr:
                if any(k == x or k.startswith(x) for x in KEYS) or ('issue_stalled' in k and k.endswith('per_issue_active.ratio')) or k in EXTRA:
                    f.write(f'| {k} | {d[k]} | {u[k]} |\n')
            f.write('\n')
            try:
                scale = {'byte': 1, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9, 'Tbyte': 1e12}
                rb = float(d['dram__bytes_read.sum'].replace(',', '')) * scale.get(u['dram__bytes_read.sum'], 1)
                wb = float(d['dram__bytes_write.sum'].replace(',', '')) * scale.get(u['dram__bytes_write.sum'], 1)
                key = name.replace('void ', '').replace('arah::', '').split('<')[0]
                if key not in traffic:
                    tscale = {'ns': 1e-6, 'nsecond': 1e-6, 'us': 1e-3, 'usecond': 1e-3, 'ms': 1.0, 'msecond': 1.0, 's': 1e3, 'second': 1e3}
                    ms = float(d['gpu__time_duration.sum'].replace(',', '')) * tscale.get(u['gpu__time_duration.sum'], 1.0)
                    traffic[key] = {'dram_bytes_per_launch': rb + wb, 'dram_read': rb, 'dram_write': wb, 'ms': ms,
                                    'tensor_pipe_pct': d.get('sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active')}
            except Exception:
                pass
    if traffic_json:
        old = {}
        if os.path.exists(traffic_json):
            old = json.load(open(traffic_json))
        old.update(traffic)
        json.dump(old, open(traffic_json, 'w'), indent=1)
    print(open(dst).read()[:2500])


def table(srcs, dst):
    """Several raw-page CSVs -> one compact markdown table, one row per distinct kernel (its first captured launch)."""
    cols = [('ms', 'gpu__time_duration.sum'), ('grid', 'launch__grid_size'), ('block', 'launch__block_size'), ('regs', 'launch__registers_per_thread'),
            ('smem KB', 'launch__shared_mem_per_block_dynamic'), ('tensor %', 'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active'),
            ('XU %', 'sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active'), ('FMA %', 'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active'),
            ('issue %', 'smsp__issue_active.avg.pct_of_peak_sustained_active'), ('DRAM %', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed'),
            ('DRAM rd', 'dram__bytes_read.sum'), ('DRAM wr', 'dram__bytes_write.sum'), ('L2->SM', 'l1tex__m_xbar2l1tex_read_bytes.sum.per_second')]
    seen = set()
    with open(dst, 'w') as f:
        f.write('# ncu --set full --clock-control none: one launch per kernel (' + ', '.join(srcs) + ')\n\n'
                'Numbers under ncu are never bench values; times are cold-cache and serialised.  Units as ncu prints them.\n\n')
        f.write('| kernel | ' + ' | '.join(c for c, _ in cols) + ' |\n|---|' + '---:|' * len(cols) + '\n')
        for src in srcs:
            rd = list(csv.reader(open(src)))
            hdr, units, data = rd[0], rd[1], rd[2:]
            u = dict(zip(hdr, units))
            for row in data:
                d = dict(zip(hdr, row))
                name = re.sub(r'\(.*', '', d.get('Kernel Name', '?')).strip().replace('void ', '')
                name = re.sub(r'^(arah\w*::)+', '', name)
                if name in seen:
                    continue
                seen.add(name)
                cells = []
                for c, k in cols:
                    v = d.get(k, '')
                    try:
                        x = float(v.replace(',', ''))
                        v = f'{x:.3g}' if abs(x) < 1000 else f'{x:.0f}'
                    except Exception:
                        pass
                    unit = u.get(k, '')
                    if c in ('ms', 'DRAM rd', 'DRAM wr', 'L2->SM') and unit not in ('', '%'):
                        v += ' ' + unit.replace('second', 's').replace('byte', 'B')
                    cells.append(v)
                f.write(f'| `{name}` | ' + ' | '.join(cells) + ' |\n')
    print(open(dst).read())


def rep(src, dst, traffic_json=None):
    out = subprocess.run(['ncu', '-i', src, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rd = list(csv.reader(io.StringIO(out)))
    hdr, units, data = rd[0], rd[1], rd[2:]
    with open(dst, 'w') as f:
        f.write(f'# ncu --set full summary ({src})\n\n')
        for row in data:
            d = dict(zip(hdr, row))
            u = dict(zip(hdr, units))
            f.write(f"## {d.get('Kernel Name', '?')}  (ID {d.get('ID')})\n\n| metric | value | unit |\n|---|---:|---|\n")
            for k in hdr:
                if any(k == x or k.startswith(x) for x in KEYS) or 'stall' in k.lower() and 'pct' in k.lower():
                    f.write(f'| {k} | {d[k]} | {u[k]} |\n')
            f.write('\n')
            if traffic_json:
                try:
                    rd_b = float(d['dram__bytes_read.sum'].replace(',', ''))
                    wr_b = float(d['dram__bytes_write.sum'].replace(',', ''))
                    scale = {'byte': 1, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9}
                    tot = rd_b * scale.get(u['dram__bytes_read.sum'], 1) + wr_b * scale.get(u['dram__bytes_write.sum'], 1)
                    json.dump({'dram_bytes_per_launch': tot, 'kernel': d.get('Kernel Name'), 'source': src}, open(traffic_json, 'w'))
                    traffic_json = None
                except Exception as e:
                    print('traffic parse failed', e)
    print(open(dst).read()[:6000])


if __name__ == '__main__':
    if sys.argv[1] == 'launches':
        launches(sys.argv[2], sys.argv[3])
    elif sys.argv[1] == 'table':
        table(sys.argv[2:-1], sys.argv[-1])
    elif sys.argv[1] == 'csv':
        csvpage(sys.argv[2], sys.argv[3], sys.argv[4] if len(sys.argv) > 4 else None)
    else:
        rep(sys.argv[2], sys.argv[3], sys.argv[4] if len(sys.argv) > 4 else None)
