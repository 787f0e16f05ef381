"""Summarise an `ncu --page source --csv` dump: stall samples by reason and by opcode, and the hottest SASS lines.
    python tools/ncu_src_top.py gpurun_out/x.src.csv [ntop]"""
import csv, sys, collections
rows = list(csv.reader(open(sys.argv[1])))
ntop = int(sys.argv[2]) if len(sys.argv) > 2 else 25
hdr = rows[1]; ix = {k: i for i, k in enumerate(hdr)}
body = rows[2:]
S = [int(r[ix['# Samples']]) for r in body]
print(rows[0][1][:120]); print('instructions', len(body), 'samples', sum(S))
st = [k for k in hdr if k.startswith('stall_')]
tot = collections.Counter()
for r in body:
    for k in st:
        try: tot[k] += int(r[ix[k]])
        except Exception: pass
print('stalls:', ', '.join(f'{k[6:]}={v}' for k, v in tot.most_common(12)))
byop = collections.Counter(); cnt = collections.Counter()
for r, s_ in zip(body, S):
    op = r[ix['Source']].split()
    if not op: continue
    o = op[1] if op[0].startswith('@') else op[0]
    byop[o] += s_; cnt[o] += int(r[ix['Instructions Executed']])
print('by opcode:', ', '.join(f'{o}={v}({cnt[o]//1000}k)' for o, v in byop.most_common(12)))
for i in sorted(sorted(range(len(body)), key=lambda i: -S[i])[:ntop]):
    r = body[i]
    top = sorted(((int(r[ix[k]] or 0), k[6:]) for k in st), reverse=True)[:2]
    print(i, S[i], r[ix['Source']].strip()[:80], r[ix['Instructions Executed']], top)
