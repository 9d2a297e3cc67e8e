"""Summarises ncu outputs into profiles/ (launch list CSV -> table; .ncu-rep -> key metrics)."""
import collections
import csv
import json
import subprocess
import sys


def launches(path):
    lines = [l for l in open(path) if not l.startswith('==')]
    d = collections.OrderedDict()
    for row in csv.DictReader(lines):
        e = d.setdefault(int(row['ID']), {'name': row['Kernel Name'].split('(')[0]})
        e[row['Metric Name']] = float(row['Metric Value'])
    return list(d.values())


def rep_metrics(path, want):
    out = subprocess.check_output(['ncu', '-i', path, '--page', 'raw', '--csv'], stderr=subprocess.DEVNULL).decode()
    r = list(csv.reader(out.splitlines()))
    hdr, units = r[0], r[1]
    res = []
    for row in r[2:]:
        m = {}
        for w in want:
            if w in hdr:
                i = hdr.index(w)
                m[w] = (row[i], units[i])
        res.append(m)
    return res


if __name__ == '__main__':
    tag = sys.argv[1]
    ls = launches('gpurun_out/launches_%s.csv' % tag)
    tot = sum(l['gpu__time_duration.sum'] for l in ls)
    agg = collections.OrderedDict()
    for l in ls:
        a = agg.setdefault(l['name'], [0, 0.0, 0.0])
        a[0] += 1
        a[1] += l['gpu__time_duration.sum']
        a[2] += l.get('dram__bytes_read.sum', 0) + l.get('dram__bytes_write.sum', 0)
    with open('profiles/%s_launches.txt' % tag, 'w') as f:
        f.write('# ncu --metrics gpu__time_duration.sum,dram__bytes_{read,write}.sum --clock-control none '
                '(python bench.py --steps 6 --warmup 3, 512^3 D3Q19 BGK AA); cold-cache, serialised: compare shares\n')
        f.write('%-58s %6s %12s %8s %14s\n' % ('kernel', 'count', 'avg us', 'share', 'avg DRAM MB'))
        for k, (n, t, b) in agg.items():
            f.write('%-58s %6d %12.1f %7.1f%% %14.1f\n' % (k, n, t / n / 1e3, 100 * t / tot, b / n / 1e6))
    print(open('profiles/%s_launches.txt' % tag).read())
    want = ['Kernel Name', 'gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
            'dram__bytes.sum.per_second', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
            'sm__warps_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread',
            'launch__occupancy_limit_registers', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
            'smsp__issue_active.avg.pct_of_peak_sustained_active', 'l1tex__t_sector_hit_rate.pct',
            'lts__t_sector_hit_rate.pct', 'l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum',
            'l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum', 'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active',
            'smsp__inst_executed.sum']
    ms = rep_metrics('gpurun_out/prof_bulk_%s.ncu-rep' % tag, want)
    with open('profiles/%s_bulk_full.txt' % tag, 'w') as f:
        f.write('# ncu --set full --clock-control none --import-source on -k regex:collide_bulk -s 8 -c 2 '
                '(python bench.py, 512^3 D3Q19 BGK AA)\n')
        for m in ms:
            for k, (v, u) in m.items():
                f.write('%-70s %s %s\n' % (k, v, u))
            f.write('\n')
    print(open('profiles/%s_bulk_full.txt' % tag).read())
    # traffic for bench.py: DRAM bytes per launch of the dominant kernel (mean of even / odd)
    tr = [float(m['dram__bytes_read.sum'][0]) * (1e9 if m['dram__bytes_read.sum'][1] == 'Gbyte' else 1e6) +
          float(m['dram__bytes_write.sum'][0]) * (1e9 if m['dram__bytes_write.sum'][1] == 'Gbyte' else 1e6) for m in ms]
    json.dump({'dram_bytes_per_launch': sum(tr) / len(tr), 'source': 'profiles/%s_bulk_full.txt' % tag,
               'workload': 'ldc_3d D3Q19 BGK AA 512^3'}, open('profiles/traffic.json', 'w'))
