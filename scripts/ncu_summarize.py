"""Turns ncu output into the markdown summaries kept under profiles/.

    python scripts/ncu_summarize.py launches <launches.csv>          # per-kernel totals of a gpu__time_duration pass
    python scripts/ncu_summarize.py full <raw.csv>                    # selected metrics of `ncu -i rep --page raw --csv`
"""
import collections
import csv
import re
import sys


def short(name):
    name = re.sub(r'\(.*', '', name)
    return name.replace('b200pose::', '').replace('void ', '')[:60]


def launches(path):
    rows = [r for r in csv.reader(open(path)) if len(r) > 10]
    hdr = rows[0]
    ki, vi = hdr.index('Kernel Name'), hdr.index('Metric Value')
    tot, cnt = collections.OrderedDict(), collections.Counter()
    for r in rows[1:]:
        k = short(r[ki])
        tot[k] = tot.get(k, 0.0) + float(r[vi].replace(',', '')) / 1e3
        cnt[k] += 1
    s = sum(tot.values())
    print('| kernel | launches | total us | share |\n|---|---|---|---|')
    for k, v in sorted(tot.items(), key=lambda kv: -kv[1]):
        print('| %s | %d | %.1f | %.1f%% |' % (k, cnt[k], v, 100 * v / s))
    print('\nTotal %.1f us over %d launches (cold-cache, serialised).' % (s, sum(cnt.values())))


COLS = [('gpu__time_duration.sum', 'us', 1e-3), ('dram__bytes_read.sum', 'dram rd MB', 1e-6), ('dram__bytes_write.sum', 'dram wr MB', 1e-6),
        ('gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'dram %', 1), ('sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active', 'tensor active %', 1),
        ('sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active', 'fp64 pipe %', 1), ('smsp__issue_active.avg.pct_of_peak_sustained_active', 'issue active %', 1),
        ('sm__warps_active.avg.pct_of_peak_sustained_active', 'warps active %', 1), ('launch__waves_per_multiprocessor', 'waves/SM', 1),
        ('lts__t_sector_hit_rate.pct', 'L2 hit %', 1), ('sm__throughput.avg.pct_of_peak_sustained_elapsed', 'SM %', 1),
        ('launch__registers_per_thread', 'regs', 1)]


def full(path):
    rows = list(csv.reader(open(path)))
    hdr, units = rows[0], rows[1]
    idx = {}
    for name, _, _ in COLS:
        cands = [i for i, h in enumerate(hdr) if h == name or h.endswith('.' + name)]
        idx[name] = cands[0] if cands else None
    ki, gi = hdr.index('Kernel Name'), hdr.index('Grid Size')
    print('| kernel | grid | ' + ' | '.join(c[1] for c in COLS) + ' |')
    print('|---|---|' + '---|' * len(COLS))
    for r in rows[2:]:
        if len(r) != len(hdr):
            continue
        vals = []
        for name, _, scale in COLS:
            i = idx[name]
            if i is None or r[i] in ('', 'no data', 'n/a'):
                vals.append('-')
                continue
            v = float(r[i].replace(',', ''))
            u = units[i]
            if name == 'gpu__time_duration.sum':
                v = v * {'ns': 1e-3, 'us': 1.0, 'ms': 1e3, 's': 1e6}.get(u, 1e-3)
            elif name.startswith('dram__bytes'):
                v = v * {'byte': 1e-6, 'Kbyte': 1e-3, 'Mbyte': 1.0, 'Gbyte': 1e3}.get(u, 1e-6)
            vals.append('%.2f' % v)
        print('| %s | %s | ' % (short(r[ki]), r[gi]) + ' | '.join(vals) + ' |')


def traffic(path):
    """DRAM bytes per kernel CLASS of the step (read + written), as JSON: what bench.py reports as roofline.traffic."""
    import json
    rows = list(csv.reader(open(path)))
    hdr, units = rows[0], rows[1]
    ki = hdr.index('Kernel Name')
    ri, wi = hdr.index('dram__bytes_read.sum'), hdr.index('dram__bytes_write.sum')
    scale = {'byte': 1.0, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9}
    out, phase = {}, 'gat'
    for r in rows[2:]:
        if len(r) != len(hdr):
            continue
        k = short(r[ki])
        b = float(r[ri].replace(',', '')) * scale.get(units[ri], 1.0) + float(r[wi].replace(',', '')) * scale.get(units[wi], 1.0)
        if k.startswith('cluster_kernel'):
            phase = 'mlp'
        cls = ('gat_projection_gemm' if phase == 'gat' else 'mlp_gemm') if k.startswith('gemm_') else \
              'edge_softmax_aggregate' if k.startswith('gat_aggregate') else 'encode_dlt' if k.startswith('lift_person') else \
              'cluster' if k.startswith(('cluster_kernel', 'exclusive_scan', 'gather_persons')) else \
              'node_features' if k.startswith('head_features') else 'graph_build' if k.startswith('build_graph') else k
        out[cls] = out.get(cls, 0.0) + b
    print(json.dumps({k: round(v) for k, v in out.items()}, indent=1))


if __name__ == '__main__':
    {'launches': launches, 'full': full, 'traffic': traffic}[sys.argv[1]](sys.argv[2])
