"""tools/ncu_bandwidth.py <ncu csv> — per-kernel duration, DRAM bytes and achieved GB/s of the bandwidth kernels from
  NPM_EW_ONCE=1 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none \
      --csv --log-file gpurun_out/ew_ncu.csv python tools/ew_bench.py
(one launch per kernel at the cfg5 sizes, L2 flushed before each).  Algorithmic bytes per kernel as in DESIGN.md §4.3."""
import collections
import csv
import re
import sys

PEAK = 6543.7      # MEASURED_PEAKS.json hbm_gbs
ALGO = [   # launch order of tools/ew_bench.py: (label, algorithmic bytes)
    ('layernorm_fwd [8192,1024]', 8 * 8192 * 1024), ('layernorm_bwd', 12 * 8192 * 1024), ('dropout_fwd', 8 * 8192 * 1024),
    ('dropout+layernorm fwd', 8 * 8192 * 1024), ('dropout+layernorm bwd + residual', 16 * 8192 * 1024),
    ('add_inplace', 12 * 8192 * 1024), ('add3', 16 * 8192 * 1024), ('colsum [8192,1024]', 4 * 8192 * 1024),
    ('colsum [8192,4096]', 4 * 8192 * 4096), ('relu_bwd+colsum [8192,4096]', 12 * 8192 * 4096), ('softmax_fwd [8192,1024]', 8 * 8192 * 1024),
    ('adam_multi 16.8M params', 28 * 16798720),
    ('adam_multi + weight planes', 28 * 16798720 + 4 * 16777216), ('weight_split of y (set-up, not a result)', 8 * 8192 * 4096),
    ('relu_bwd+colsum planes [8192,4096]', 10 * 8192 * 4096), ('dropout+layernorm fwd -> planes', 8 * 8192 * 1024),
]


def main(path):
    with open(path) as f:
        lines = [l for l in f if not l.startswith('==')]
    rows = collections.OrderedDict()
    for r in csv.DictReader(lines):
        k = (int(r['ID']), re.sub(r'\(.*', '', r['Kernel Name']).replace('void ', '').replace('npm::<unnamed>::', ''))
        rows.setdefault(k, {})[r['Metric Name']] = (float(r['Metric Value'].replace(',', '')), r['Metric Unit'])
    def to_bytes(v):
        val, unit = v
        return val * {'byte': 1, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9}[unit]
    def to_us(v):
        val, unit = v
        return val * {'ns': 1e-3, 'us': 1, 'ms': 1e3, 'nsecond': 1e-3, 'usecond': 1, 'msecond': 1e3}[unit]
    mine = [(k, m) for k, m in rows.items() if 'at::' not in k[1] and 'elementwise_kernel' not in k[1] and 'distribution' not in k[1]]
    print(f'# {len(mine)} launches of this library; peak = {PEAK} GB/s (MEASURED_PEAKS.json); one launch each, L2 flushed before it')
    print(f'# {"kernel":44s} {"us":>8s} {"dram MB":>9s} {"algo MB":>9s} {"traffic/algo":>12s} {"algo GB/s":>10s} {"frac":>6s}')
    groups, i = [], 0
    # kernels that run as two launches (first stage + reduce_partials) are summed into the label they belong to
    for (kid, name), m in mine:
        if name.startswith('reduce_partials') or name.startswith('fill'):
            if groups:
                groups[-1][1].append(m); groups[-1][2].append(name)
            continue
        groups.append([name, [m], [name]])
    for gi, (name, ms, names) in enumerate(groups):
        us = sum(to_us(m['gpu__time_duration.sum']) for m in ms)
        dram = sum(to_bytes(m['dram__bytes_read.sum']) + to_bytes(m['dram__bytes_write.sum']) for m in ms)
        label, algo = ALGO[gi] if gi < len(ALGO) else (name, 0)
        gbs = algo / us / 1e3 if us else 0
        print(f'{label:34s} {"+".join(n[:18] for n in names)[:40]:40s} {us:8.1f} {dram / 1e6:9.1f} {algo / 1e6:9.1f} {dram / algo if algo else 0:12.2f} {gbs:10.1f} {gbs / PEAK:6.2f}')


if __name__ == '__main__':
    main(sys.argv[1])
