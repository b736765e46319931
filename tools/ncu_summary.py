"""tools/ncu_summary.py — condense ncu output into the small text summaries committed under profiles/.
  python tools/ncu_summary.py launches gpurun_out/launches.csv   # per-kernel share of one training step
  python tools/ncu_summary.py full gpurun_out/prof.ncu-rep       # key metrics of a --set full capture
"""
import collections
import csv
import re
import subprocess
import sys


def launches(path):
    with open(path) as f:
        lines = [l for l in f if not l.startswith('==')]
    rows = [(int(r['ID']), r['Kernel Name'], float(r['Metric Value'].replace(',', '')))
            for r in csv.DictReader(lines) if r.get('Metric Name') == 'gpu__time_duration.sum']
    adam = [i for i, (_, k, _) in enumerate(rows) if 'opt_multi' in k]
    print(f'# {len(rows)} launches captured; optimizer launches at {adam}')
    if len(adam) < 4:
        a, b = 0, len(rows)
    else:
        a, b = adam[2] + 1, adam[3] + 1        # the step after 3 warm-up steps
    step = rows[a:b]
    tot = sum(v for _, _, v in step)
    agg = collections.defaultdict(lambda: [0, 0.0])
    for _, k, v in step:
        name = re.sub(r'\(.*', '', k).replace('void ', '').replace('npm::<unnamed>::', '')
        agg[name][0] += 1
        agg[name][1] += v
    print(f'# one training step: {len(step)} launches, sum of kernel durations {tot / 1e6:.3f} ms '
          f'(ncu: cold-cache, serialised — compare SHARES)')
    for k, (n, v) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f'{v / 1e3:10.1f} us {100 * v / tot:5.1f}%  n={n:4d}  {k[:100]}')


KEYS = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed',
        'sm__pipe_tensor_subpipe_hmma_cycles_active_realtime.avg', 'sm__mem_tensor_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'lts__t_bytes.sum', 'lts__t_sector_hit_rate.pct',
        'launch__registers_per_thread', 'launch__grid_size', 'launch__block_size', 'launch__shared_mem_per_block_dynamic',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'sm__cycles_elapsed.avg', 'smsp__inst_executed.sum',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'sm__cycles_active.avg']


def full(path):
    out = subprocess.run(['ncu', '-i', path, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    r = list(csv.reader(out.splitlines()))
    hdr, units = r[0], r[1]
    for row in r[2:]:
        d = dict(zip(hdr, row))
        u = dict(zip(hdr, units))
        print('kernel:', d.get('Kernel Name', '')[:120], ' grid', d.get('launch__grid_size'), 'block', d.get('launch__block_size'))
        for k in KEYS:
            if k in d:
                print(f'   {k:80s} {d[k]:>16s} {u[k]}')


if __name__ == '__main__':
    {'launches': launches, 'full': full}[sys.argv[1]](sys.argv[2])
