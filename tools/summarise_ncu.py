#!/usr/bin/env python
"""Summarises ncu output into the small text files kept under profiles/.

    python tools/summarise_ncu.py launches <launches.csv> <steps_in_run> > profiles/rNN_launches_summary.md
    python tools/summarise_ncu.py full <prof.ncu-rep>                     > profiles/rNN_attn_full_summary.md

`launches`: per-kernel totals of a `--metrics gpu__time_duration.sum` launch list (cold-cache, serialised times:
compare SHARES, not absolutes).  `full`: the handful of counters the roofline block cites, per captured launch.
"""
import csv
import io
import re
import subprocess
import sys
from collections import defaultdict


def short(name: str) -> str:
    name = re.sub(r'\(.*', '', name)
    name = re.sub(r'<.*', '', name)
    name = name.replace('void ', '').strip()
    return name[-70:]


def launches(path: str, steps: float):
    rows = []
    with open(path, newline='') as f:
        text = f.read()
    start = text.index('"ID"')
    for r in csv.DictReader(io.StringIO(text[start:])):
        if r['Metric Name'] != 'gpu__time_duration.sum':
            continue
        v = float(r['Metric Value'].replace(',', ''))
        unit = r['Metric Unit']
        us = v / 1e3 if unit in ('ns', 'nsecond') else v if unit in ('us', 'usecond') else v * 1e3 if unit in ('ms', 'msecond') else v
        rows.append((r['Kernel Name'], us))
    tot = defaultdict(lambda: [0, 0.0])
    for n, us in rows:
        k = short(n)
        tot[k][0] += 1
        tot[k][1] += us
    total = sum(v[1] for v in tot.values())
    print(f'# ncu launch list summary: {len(rows)} launches, {total / 1e3:.1f} ms total device time '
          f'({steps:g} steps incl. warm-up and e2e leg -> {total / 1e3 / steps:.1f} ms/step serialised, cold cache)\n')
    print('| kernel | launches | total ms | share | us/launch |')
    print('|---|---:|---:|---:|---:|')
    for k, (n, us) in sorted(tot.items(), key=lambda kv: -kv[1][1])[:40]:
        print(f'| `{k}` | {n} | {us / 1e3:.2f} | {100 * us / total:.1f}% | {us / n:.1f} |')
    mine = {k: v for k, v in tot.items() if k.startswith('svae::')}
    ms = sum(v[1] for v in mine.values())
    print(f'\nLibrary kernels (svae::*): {sum(v[0] for v in mine.values())} launches, {ms / 1e3:.2f} ms = {100 * ms / total:.1f}% of device time.')


METRICS = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'dram__throughput.avg.pct_of_peak_sustained_elapsed',
           'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
           'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_tensor.sum',
           'sm__warps_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread', 'launch__grid_size',
           'launch__block_size', 'launch__occupancy_limit_registers', 'launch__occupancy_limit_shared_mem',
           'smsp__inst_executed.sum', 'sm__inst_executed_pipe_xu.sum', 'smsp__cycles_active.avg',
           'l1tex__t_bytes.sum', 'lts__t_bytes.sum', 'lts__t_sector_hit_rate.pct', 'sm__cycles_elapsed.max',
           'smsp__average_warp_latency_issue_stalled_long_scoreboard.ratio', 'smsp__issue_active.avg.pct_of_peak_sustained_active']


def full(path: str):
    out = subprocess.run(['ncu', '-i', path, '--page', 'raw', '--csv'], capture_output=True, text=True, check=True).stdout
    start = out.index('"ID"')
    rd = list(csv.reader(io.StringIO(out[start:])))
    hdr, units, data = rd[0], rd[1], rd[2:]
    idx = {h: i for i, h in enumerate(hdr)}
    print(f'# ncu --set full summary of {path.split("/")[-1]} ({len(data)} launches; cold-cache, clocks uncontrolled)\n')
    names = [short(r[idx['Kernel Name']]) for r in data]
    print('| metric | unit | ' + ' | '.join(f'`{n}`' for n in names) + ' |')
    print('|---|---|' + '---:|' * len(names))
    for m in METRICS:
        if m not in idx:
            continue
        print(f'| {m} | {units[idx[m]]} | ' + ' | '.join(r[idx[m]] for r in data) + ' |')
    for r in data:
        rd_b, wr_b = r[idx['dram__bytes_read.sum']], r[idx['dram__bytes_write.sum']]
        print(f'\n{short(r[idx["Kernel Name"]])}: DRAM read {rd_b} {units[idx["dram__bytes_read.sum"]]}, write {wr_b} {units[idx["dram__bytes_write.sum"]]}')


if __name__ == '__main__':
    if sys.argv[1] == 'launches':
        launches(sys.argv[2], float(sys.argv[3]) if len(sys.argv) > 3 else 1)
    else:
        full(sys.argv[2])
