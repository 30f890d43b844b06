"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel launches, total and share."""
import collections, csv, sys


def main():
    src, dst, note = sys.argv[1], sys.argv[2], (sys.argv[3] if len(sys.argv) > 3 else '')
    rows = list(csv.reader(open(src)))
    for i, r in enumerate(rows):
        if 'Kernel Name' in r:
            hdr, data = r, rows[i + 1:]
            break
    ix = {h: i for i, h in enumerate(hdr)}
    agg = collections.OrderedDict()
    for r in data:
        if len(r) < len(hdr) or r[ix['Metric Name']] != 'gpu__time_duration.sum':
            continue
        name = r[ix['Kernel Name']].split('(')[0][:72]
        v = float(r[ix['Metric Value']].replace(',', ''))
        unit = r[ix['Metric Unit']]
        v = v / 1000 if unit in ('ns', 'nsecond') else (v * 1000 if unit in ('ms', 'msecond') else v)
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += v
    spin = sum(a[1] for k, a in agg.items() if 'spin_kernel' in k)
    tot = sum(a[1] for a in agg.values()) - spin
    out = [note, '# cold-cache, serialised launches: compare SHARES, not absolutes.  (torch.cuda._sleep spin kernel of the'
           ' instrumented pass excluded from the shares.)',
           f"{'kernel':72s} {'launches':>8s} {'total_us':>10s} {'avg_us':>9s} {'share':>6s}"]
    for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        sh = '   -  ' if 'spin_kernel' in k else f'{100 * a[1] / tot:5.1f}%'
        out.append(f'{k:72s} {a[0]:8d} {a[1]:10.1f} {a[1] / a[0]:9.2f} {sh}')
    out.append(f'total {tot:.1f} us (without spin) over {sum(a[0] for a in agg.values())} launches')
    open(dst, 'w').write('\n'.join(out) + '\n')
    print('\n'.join(out))


main()
