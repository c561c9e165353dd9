"""Attributes the per-instruction columns of `ncu --page source --csv` (SASS view) to CUDA source lines, using the line
table of the same kernel in `nvdisasm -g <cubin>` (instruction order is the same in both listings).

    cuobjdump -xelf all libb200pose.so; nvdisasm -g gat.sm_100a.cubin > gat.sass
    ncu -i rep --page source --csv --kernel-name regex:K --launch-skip S --launch-count 1 > src.csv
    python scripts/ncu_by_line.py src.csv gat.sass <mangled-kernel-substring> [top]
"""
import csv
import re
import sys


def sass_lines(path, kernel):
    """[(line_no, inlined_chain)] per instruction of the kernel's .text section."""
    out, cur, active = [], None, False
    pat_file = re.compile(r'//## File "([^"]+)", line (\d+)(.*)')
    for l in open(path):
        if l.startswith('.text.'):
            active = kernel in l
            continue
        if not active:
            continue
        m = pat_file.search(l)
        if m:
            if 'inlined at' in l:
                # the outermost location of an inlined chain is the one in the kernel body
                chain = re.findall(r'line (\d+)', l)
                cur = int(chain[-1])
            else:
                cur = int(m.group(2))
            continue
        s = l.strip()
        if s.startswith('/*') and ';' in s:          # an instruction line: /*0000*/ OP ... ;
            out.append(cur)
    return out


def main():
    src_csv, sass, kernel = sys.argv[1], sys.argv[2], sys.argv[3]
    top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
    rows = list(csv.reader(open(src_csv)))
    hdr = [i for i, r in enumerate(rows) if r and r[0] == 'Address'][0]
    H = rows[hdr]
    data, seen = [], set()
    for r in rows[hdr + 1:]:
        if len(r) == len(H) and r[0] != 'Address' and r[0] not in seen:
            seen.add(r[0])
            data.append(r)
    lines = sass_lines(sass, kernel)
    print('instructions: ncu %d, nvdisasm %d' % (len(data), len(lines)))
    n = min(len(data), len(lines))
    si, ie = H.index('# Samples'), H.index('Instructions Executed')
    agg = {}
    for r, ln in zip(data[:n], lines[:n]):
        a = agg.setdefault(ln, [0, 0, 0])
        a[0] += int(r[si]); a[1] += int(r[ie]); a[2] += 1
    ts, ti = sum(a[0] for a in agg.values()), sum(a[1] for a in agg.values())
    src = open('/root/repo/3d_multi_pose_estimator_b200/csrc/' + ('gat.cu' if 'gat' in sass else 'gemm.cu')).read().splitlines()
    print('total samples %d, warp-instructions executed %d' % (ts, ti))
    print('%6s %7s %7s %5s  %s' % ('line', 'samp%', 'inst%', 'sass', 'source'))
    for ln, a in sorted(agg.items(), key=lambda kv: -kv[1][1])[:top]:
        text = src[ln - 1].strip()[:110] if ln and ln <= len(src) else '?'
        print('%6s %6.1f%% %6.1f%% %5d  %s' % (ln, 100.0 * a[0] / max(ts, 1), 100.0 * a[1] / max(ti, 1), a[2], text))


if __name__ == '__main__':
    main()
