"""Static evidence for profiles/: per kernel of libsessrec_b200.so the tcgen05 / TMEM / TMA SASS instruction counts
(cuobjdump -sass) and registers / spills / static shared memory (cuobjdump -res-usage).  No GPU needed."""
import collections
import re
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
LIB = ROOT / 'sessionrec-pytorch_b200' / 'libsessrec_b200.so'
PAT = {'UTC*MMA (tcgen05.mma)': r'\bUTC\w*MMA\b', 'LDTM (tcgen05.ld)': r'\bLDTM\b', 'STTM (tcgen05.st)': r'\bSTTM\b',
       'UTMALDG (TMA load)': r'\bUTMALDG\b', 'UTMASTG (TMA store)': r'\bUTMASTG\b', 'UTMAREDG (TMA reduce-add)': r'\bUTMAREDG\b',
       'UBLKCP (bulk copy)': r'\bUBLKCP\b', 'ACQBULK (griddepcontrol.wait)': r'\bACQBULK\b', 'PREEXIT (launch_dependents)': r'\bPREEXIT\b', 'ELECT': r'\bELECT\b', 'BRA.U.ANY (vote loop)': r'BRA\.U\.ANY', 'HMMA (legacy mma.sync)': r'\bHMMA\b', 'SYNCS (mbarrier)': r'\bSYNCS\b'}


def demangle(names):
    out = subprocess.run(['c++filt'], input='\n'.join(names), capture_output=True, text=True).stdout.splitlines()
    return {n: re.sub(r'\(.*', '', d.replace('(anonymous namespace)::', '')).replace('void ', '') for n, d in zip(names, out)}


def main():
    sass = subprocess.run(['cuobjdump', '-sass', str(LIB)], capture_output=True, text=True).stdout
    res = subprocess.run(['cuobjdump', '-res-usage', str(LIB)], capture_output=True, text=True).stdout
    counts, cur = collections.OrderedDict(), None
    for line in sass.splitlines():
        m = re.match(r'\s*Function : (\S+)', line)
        if m:
            cur = m.group(1)
            counts[cur] = collections.Counter()
            continue
        if cur:
            for k, p in PAT.items():
                if re.search(p, line):
                    counts[cur][k] += 1
    usage = {}
    fn = None
    for line in res.splitlines():
        m = re.match(r'\s*Function (\S+):', line)
        if m:
            fn = m.group(1)
            continue
        m = re.search(r'REG:(\d+).*?SHARED:(\d+).*?LOCAL:(\d+)', line)
        if m and fn:
            usage[fn] = tuple(map(int, m.groups()))
    names = demangle(list(counts))
    print('# Static SASS evidence, libsessrec_b200.so (sm_100a)\n')
    print('Kernels that use the Blackwell tensor / TMA paths (instruction counts in the SASS of each kernel):\n')
    cols = list(PAT)
    print('| kernel | ' + ' | '.join(cols) + ' | regs | static smem B | local (spill) B |')
    print('|---|' + '---|' * (len(cols) + 3))
    for fn, c in counts.items():
        if not any(c[k] for k in cols[:7]):
            continue
        r = usage.get(fn, ('?', '?', '?'))
        print(f'| `{names[fn]}` | ' + ' | '.join(str(c[k]) for k in cols) + f' | {r[0]} | {r[1]} | {r[2]} |')
    print('\nAll kernels: registers / static shared memory / local memory (non-zero local = spills or local arrays):\n')
    print('| kernel | regs | static smem B | local B |')
    print('|---|---|---|---|')
    for fn in counts:
        r = usage.get(fn, ('?', '?', '?'))
        print(f'| `{names[fn]}` | {r[0]} | {r[1]} | {r[2]} |')
    hm = sum(c['HMMA (legacy mma.sync)'] for c in counts.values())
    print(f'\nLegacy `HMMA` instructions in the whole library: {hm}.')


if __name__ == '__main__':
    sys.exit(main())
