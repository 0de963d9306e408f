"""Counts of Blackwell-specific and transfer mnemonics per kernel of the built library (no GPU needed):
    cuobjdump -sass x-transformers-rl_b200/lib/libxtrl.so | python scripts/sass_mnemonics.py > profiles/rNN_sass_blackwell_mnemonics.txt"""
import collections, re, subprocess, sys
cur, cnt = None, collections.Counter()
pat = re.compile(r'(UTCHMMA|UTCQMMA|UTCOMMA|LDTM|STTM|UBLKCP|UTCBAR|UTCATOMSWS|SYNCS\.ARRIVE\.TRANS64|LDGSTS|FFMA2|FMUL2|FADD2|REDG|RED\.|ATOMG|'
                 r'LD\.E\.\S*SYS|ST\.E\.\S*SYS|LDG\.E\.\S*SYS|STG\.E\.\S*SYS)')
for line in sys.stdin:
    m = re.search(r'Function : (\S+)', line)
    if m:
        cur = m.group(1)
        continue
    mm = re.search(r'^\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Za-z0-9_.]+)', line)
    if cur and mm and pat.match(mm.group(1)):
        cnt[(cur, mm.group(1))] += 1
names = sorted({k for k, _ in cnt})
dem = dict(zip(names, subprocess.run(['c++filt'] + names, capture_output=True, text=True).stdout.strip().split('\n'))) if names else {}
print('cuobjdump -sass x-transformers-rl_b200/lib/libxtrl.so (sm_100a): Blackwell-specific and transfer mnemonics per kernel')
print('UTCHMMA = tcgen05.mma, LDTM / STTM = tcgen05.ld / st, UBLKCP = cp.async.bulk (TMA 1-D), LDGSTS = cp.async, FFMA2 / FMUL2 / FADD2 = packed')
print('fp32 pairs; the .SYS loads / stores of opt_fused_kernel are the peer-memory gradient exchange\n')
for (k, op), c in sorted(cnt.items()):
    print(f'{c:5d} {dem.get(k, k)[:120]} :: {op}')
