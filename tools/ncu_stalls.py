"""summarise an `ncu --page source --csv` export: per kernel stall mix, opcode mix and the hottest instructions"""
import csv, collections, sys
rows = list(csv.reader(open(sys.argv[1])))
topn = int(sys.argv[2]) if len(sys.argv) > 2 else 14
# the export concatenates kernels: a row with "Kernel Name" starts each
blocks, cur = [], None
for r in rows:
    if r and r[0] == 'Kernel Name':
        cur = {'name': r[1], 'hdr': None, 'rows': []}
        blocks.append(cur)
    elif cur is not None and cur['hdr'] is None:
        cur['hdr'] = r
    elif cur is not None:
        cur['rows'].append(r)
for b in blocks:
    hdr = b['hdr']; ix = {h: i for i, h in enumerate(hdr)}
    stalls = [h for h in hdr if h.startswith('stall_') and 'Not Issued' not in h]
    tot = collections.Counter(); nsamp = 0; opc = collections.Counter(); hot = []
    for r in b['rows']:
        try:
            ex = int(r[ix['Instructions Executed']]); s = int(r[ix['# Samples']])
        except Exception:
            continue
        nsamp += s
        for h in stalls:
            try: tot[h] += int(r[ix[h]])
            except Exception: pass
        src = r[ix['Source']]
        t = src.split()
        op = (t[1] if t and t[0].startswith('@') else (t[0] if t else '?'))
        opc[op] += ex
        hot.append((s, r[ix['Address']], src, {h: int(r[ix[h]] or 0) for h in stalls if (r[ix[h]] or '0') != '0'}))
    tex = sum(opc.values())
    print("==", b['name'][:60], "samples", nsamp, "instr", tex)
    print("  stalls:", ", ".join("%s %.1f%%" % (h.replace('stall_', ''), 100 * v / max(nsamp, 1)) for h, v in tot.most_common(8)))
    print("  opcodes:", ", ".join("%s %.1f%%" % (op, 100 * v / max(tex, 1)) for op, v in opc.most_common(14)))
    hot.sort(reverse=True)
    for s, a, src, st in hot[:topn]:
        print("   %5.2f%% %s %-50s %s" % (100 * s / max(nsamp, 1), a[-5:], src[:50], " ".join("%s=%d" % (k.replace('stall_', ''), v) for k, v in sorted(st.items(), key=lambda kv: -kv[1])[:3])))
