"""Turns gpurun_out/ ncu artefacts into the tracked summaries under profiles/.

    python tools/summarize_profiles.py launches gpurun_out/launches_X.csv profiles/NAME   -> NAME.csv + NAME.md
    python tools/summarize_profiles.py full gpurun_out/prof_X.ncu-rep profiles/NAME.json [workload_key]
"""
import collections
import csv
import json
import os
import re
import subprocess
import sys

KEEP = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed',
        'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'lts__throughput.avg.pct_of_peak_sustained_elapsed', 'launch__registers_per_thread',
        'sm__warps_active.avg.per_cycle_active',
        'l1tex__data_pipe_tc_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed', 'smsp__inst_executed.sum',
        'sm__cycles_elapsed.max', 'sm__cycles_elapsed.max.per_second', 'lts__t_sectors_srcunit_tex.sum',
        'lts__t_sector_hit_rate.pct', 'launch__grid_size', 'launch__block_size',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'smsp__sass_inst_executed_op_tmem_ldt.sum',
        'smsp__average_warps_issue_stalled_wait_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio']


def short_name(name):
    name = re.sub(r'\(.*$', '', name)
    return name.replace('void ', '').strip()


def launches(src, dst):
    rows = []
    with open(src) as f:
        lines = [ln for ln in f if ln.startswith('"')]
    for rec in csv.DictReader(lines):
        if rec.get('Metric Name') == 'gpu__time_duration.sum':
            rows.append((short_name(rec['Kernel Name']), float(rec['Metric Value'].replace(',', '')) / 1e6))
    with open(dst + '.csv', 'w') as f:
        f.write('launch,kernel,ms\n')
        for i, (name, ms) in enumerate(rows):
            f.write('%d,"%s",%.6f\n' % (i, name, ms))
    agg = collections.OrderedDict()
    for name, ms in rows:
        cnt, tot = agg.get(name, (0, 0.0))
        agg[name] = (cnt + 1, tot + ms)
    total = sum(t for _, t in agg.values())
    with open(dst + '.md', 'w') as f:
        f.write('# ncu launch list (%s)\n\n`ncu --metrics gpu__time_duration.sum --clock-control none`; times are '
                'cold-cache and serialised: compare shares.\n\n| kernel | launches | total ms | share |\n|---|---|---|---|\n'
                % os.path.basename(src))
        for name, (cnt, tot) in agg.items():
            f.write('| %s | %d | %.3f | %.1f %% |\n' % (name, cnt, tot, 100.0 * tot / total))
    print(open(dst + '.md').read())


def full(src, dst, key=None):
    out = subprocess.run(['ncu', '-i', src, '--page', 'raw', '--csv'], stdout=subprocess.PIPE, text=True).stdout
    rows = list(csv.reader([ln for ln in out.splitlines() if ln.startswith('"')]))
    hdr, units, vals = rows[0], rows[1], rows[2]
    rec = {'kernel': vals[hdr.index('Kernel Name')], 'source': os.path.basename(src)}
    for h, u, v in zip(hdr, units, vals):
        if h in KEEP:
            rec[h] = {'unit': u, 'value': v}
    with open(dst, 'w') as f:
        json.dump(rec, f, indent=1)
    print(json.dumps(rec, indent=1))
    if key:
        def gb(name):
            e = rec[name]
            scale = {'Gbyte': 1e9, 'Mbyte': 1e6, 'Kbyte': 1e3, 'byte': 1.0, 'Tbyte': 1e12}[e['unit']]
            return float(e['value'].replace(',', '')) * scale
        path = os.path.join(os.path.dirname(dst), 'k1_dram_traffic.json')
        traffic = json.load(open(path)) if os.path.exists(path) else {}
        rd, wr = gb('dram__bytes_read.sum'), gb('dram__bytes_write.sum')
        traffic[key] = {'dram_bytes_per_launch': int(rd + wr), 'read': int(rd), 'write': int(wr),
                        'source': os.path.basename(dst)}
        with open(path, 'w') as f:
            json.dump(traffic, f, indent=1)


if __name__ == '__main__':
    if sys.argv[1] == 'launches':
        launches(sys.argv[2], sys.argv[3])
    else:
        full(sys.argv[2], sys.argv[3], sys.argv[4] if len(sys.argv) > 4 else None)
