"""SASS mnemonic counts per kernel of libb200ir.so (cuobjdump -sass): tcgen05.mma = UTCHMMA (.2CTA = cta_group::2),
tcgen05.ld = LDTM, TMA loads = UTMALDG (.2CTA = pair loads), tcgen05.commit = UTCBAR (.2CTA.MULTICAST), cluster barrier =
UCGABAR, mbarrier = SYNCS; HMMA (legacy mma.sync) must be 0.  Usage: python tools/sass_summary.py > profiles/<name>.txt"""
import collections
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
so = os.path.join(ROOT, 'image_restoration_b200', 'libb200ir.so')
sass = subprocess.run(['cuobjdump', '-sass', so], capture_output=True, text=True).stdout
COLS = ['UTCHMMA', 'UTCHMMA.2CTA', 'LDTM', 'UTMALDG', 'UTMALDG.2CTA', 'UTCBAR', 'UTCBAR.2CTA', 'UCGABAR', 'UTMASTG', 'HMMA', 'SYNCS']
per = collections.OrderedDict()
cur = None
for line in sass.splitlines():
    m = re.match(r'\s*Function : (\S+)', line)
    if m:
        name = subprocess.run(['c++filt', m.group(1)], capture_output=True, text=True).stdout.strip()
        name = re.sub(r'\(.*', '', name).replace('void ', '').replace('b200ir::', '')
        cur = per.setdefault(name, collections.Counter())
        continue
    m = re.search(r'^\s*/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_.]*)', line)
    if m and cur is not None:
        op = m.group(1)
        cur['instructions'] += 1
        base = op.split('.')[0]
        if base in ('UTCHMMA', 'UTMALDG', 'UTCBAR'):
            cur[base + ('.2CTA' if '.2CTA' in op else '')] += 1
        elif base in ('LDTM', 'UTMASTG', 'HMMA', 'SYNCS'):
            cur[base] += 1
        elif base.startswith('UCGABAR'):
            cur['UCGABAR'] += 1
print('# ' + __doc__.split('Usage')[0].strip().replace('\n', ' '))
print('kernel | instructions | ' + ' | '.join(COLS))
tot = collections.Counter()
for name, c in per.items():
    if not any(c[k] for k in COLS):
        continue
    print(f'{name} | {c["instructions"]} | ' + ' | '.join(str(c[k]) for k in COLS))
    tot.update(c)
print('TOTAL | ' + str(tot['instructions']) + ' | ' + ' | '.join(str(tot[k]) for k in COLS))
