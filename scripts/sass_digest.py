"""Mnemonic counts per kernel of the in-tree library (`cuobjdump -sass`), as the markdown table kept under profiles/.

    python scripts/sass_digest.py > profiles/r02_sass_digest.md
"""
import collections
import os
import re
import subprocess
import sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(REPO, '3d_multi_pose_estimator_b200', 'libb200pose.so')
COLS = [('UTCHMMA', r'\bUTCHMMA(?!\.2CTA)'), ('UTCHMMA.2CTA', r'\bUTCHMMA\.2CTA'), ('UTMALDG (TMA load)', r'\bUTMALDG'), ('UTMASTG (TMA store)', r'\bUTMASTG'),
        ('LDTM (tcgen05.ld)', r'\bLDTM'), ('UBLKCP (bulk copy)', r'\bUBLKCP'), ('UTCBAR', r'\bUTCBAR'), ('SYNCS (mbarrier)', r'\bSYNCS'),
        ('LDGSTS (cp.async)', r'\bLDGSTS'), ('HMMA (mma.sync)', r'\bHMMA'), ('DFMA', r'\bDFMA'), ('FFMA', r'\bFFMA')]


def main():
    sass = subprocess.run(['cuobjdump', '-sass', LIB], stdout=subprocess.PIPE, check=True).stdout.decode(errors='replace')
    names = subprocess.run(['cu++filt'], input='\n'.join(re.findall(r'Function : (\S+)', sass)).encode(), stdout=subprocess.PIPE).stdout.decode().splitlines()
    parts = re.split(r'\n\s*Function : \S+\n', sass)[1:]
    head = subprocess.run(['git', 'log', '--oneline', '-1'], cwd=REPO, stdout=subprocess.PIPE).stdout.decode().split()[0]
    print('# SASS digest of 3d_multi_pose_estimator_b200/libb200pose.so (sources of commit %s)\n' % head)
    print('`cuobjdump -sass libb200pose.so` (scripts/sass_digest.py), mnemonic counts per kernel (B200_PROFILING.md: UTCHMMA = tcgen05.mma, .2CTA = cta_group::2, '
          'UTMALDG/UTMASTG = TMA tensor load/store,\nLDTM = tcgen05.ld, UBLKCP = cp.async.bulk, SYNCS = mbarrier operations). '
          'The tensor-core work is tcgen05 only: the HMMA column (mma.sync) is zero everywhere.\n')
    print('| kernel | SASS instr | ' + ' | '.join(c for c, _ in COLS) + ' |')
    print('|---|---|' + '---|' * len(COLS))
    tot = collections.Counter()
    for name, body in zip(names, parts):
        k = re.sub(r'\(.*', '', re.sub(r'\((?:int|bool)\)', '', name)).replace('b200pose::', '').replace('void ', '')
        n_instr = len(re.findall(r'/\*[0-9a-f]{4,6}\*/', body))
        counts = [len(re.findall(rx, body)) for _, rx in COLS]
        for (c, _), v in zip(COLS, counts):
            tot[c] += v
        print('| %s | %d | ' % (k, n_instr) + ' | '.join(str(v) for v in counts) + ' |')
    print('\nTotals: ' + ', '.join('%s %d' % (c, tot[c]) for c, _ in COLS))


if __name__ == '__main__':
    main()
