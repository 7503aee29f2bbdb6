"""Turns the raw ncu outputs brought back in gpurun_out/ into the tracked summaries in profiles/.

    python profiles/summarize.py <round-tag> <launches.csv> <full.ncu-rep> "<bench command>"
"""
import collections
import csv
import json
import os
import shutil
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
WANT = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'launch__registers_per_thread', 'sm__warps_active.avg.pct_of_peak_sustained_active', 'smsp__inst_executed.sum',
        'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_tensor.sum', 'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active',
        'l1tex__t_sector_hit_rate.pct', 'lts__t_sector_hit_rate.pct',
        'smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_wait_per_issue_active.ratio',
        'sm__inst_executed_pipe_tensor_subpipe_dmma.avg.pct_of_peak_sustained_active',
        'launch__grid_size', 'launch__block_size']


def to_bytes(s, u):
    return float(s) * {'Gbyte': 1e9, 'Mbyte': 1e6, 'Kbyte': 1e3, 'byte': 1}[u]


def launches(tag, path, cmd):
    lines = [l for l in open(path) if not l.startswith('==')]
    agg = collections.OrderedDict()
    n = 0
    for row in csv.DictReader(lines):
        v = float(row['Metric Value'].replace(',', ''))
        unit = row['Metric Unit']
        v = v / 1e6 if unit == 'ns' else (v / 1e3 if unit == 'us' else v)
        a = agg.setdefault(row['Kernel Name'], [0, 0.0])
        a[0] += 1
        a[1] += v
        n += 1
    tot = sum(v[1] for v in agg.values())
    with open(os.path.join(ROOT, 'profiles', f'{tag}_launches_summary.md'), 'w') as f:
        f.write(f"# {tag} -- ncu launch list\n\nCommand: `ncu --metrics gpu__time_duration.sum --clock-control none ... {cmd}`\n\n")
        f.write("Per-launch times under ncu are cold-cache and serialised: compare SHARES, not absolutes.\n\n")
        f.write(f"Total device time in the list: {tot:.1f} ms over {n} launches. Raw list: `profiles/{tag}_launches.csv`.\n\n")
        f.write("| share | total ms | launches | ms/launch | kernel |\n|---:|---:|---:|---:|---|\n")
        for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            if v[1] / tot >= 0.0002:
                f.write(f"| {100 * v[1] / tot:.1f}% | {v[1]:.3f} | {v[0]} | {v[1] / v[0]:.3f} | `{k[:120]}` |\n")
    shutil.copy(path, os.path.join(ROOT, 'profiles', f'{tag}_launches.csv'))


def full(tag, rep, cmd):
    raw = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    latest = None
    with open(os.path.join(ROOT, 'profiles', f'{tag}_ncu_full.md'), 'w') as f:
        f.write(f"# {tag} -- `ncu --set full` capture\n\nCommand: `ncu --set full --clock-control none --import-source on ... {cmd}`\n\n")
        f.write("Launches in capture order.\n\n")
        for r in rows[2:]:
            name = r[idx['Kernel Name']]
            f.write(f"## {name[:100]}\n\n| metric | value | unit |\n|---|---:|---|\n")
            for w in WANT:
                if w in idx:
                    f.write(f"| `{w}` | {r[idx[w]]} | {units[idx[w]]} |\n")
            f.write("\n")
            if 'residual_jacobian_kernel<' in name:
                # template arguments <MODEL, JAC, MINB, STRAGGLER>: the main pass with Jacobians
                args = [a.strip().replace('(int)', '').replace('(bool)', '') for a in
                        name.split('residual_jacobian_kernel<')[1].split('>')[0].split(',')]
                if len(args) >= 4 and args[1] in ('1', 'true') and args[3] in ('0', 'false'):
                    latest = r
    if latest is not None:
        tr = to_bytes(latest[idx['dram__bytes_read.sum']], units[idx['dram__bytes_read.sum']]) + \
            to_bytes(latest[idx['dram__bytes_write.sum']], units[idx['dram__bytes_write.sum']])
        json.dump({"tag": tag, "kernel": latest[idx['Kernel Name']][:80], "dram_bytes_per_launch": tr,
                   "source": f"profiles/{tag}_ncu_full.md"},
                  open(os.path.join(ROOT, 'profiles', 'jacobian_kernel_latest.json'), 'w'))


if __name__ == '__main__':
    tag, lcsv, rep = sys.argv[1:4]
    cmd = sys.argv[4] if len(sys.argv) > 4 else 'python bench.py'
    if os.path.exists(lcsv):
        launches(tag, lcsv, cmd)
    if os.path.exists(rep):
        full(tag, rep, cmd)
