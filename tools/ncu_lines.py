"""Per-source-line instruction / stall-sample shares from `ncu -i X.ncu-rep --page source --csv --print-source cuda,sass`."""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
cur, hdr, agg = None, None, {}
def num(x):
    try:
        return int(float(x))
    except ValueError:
        return 0
for r in rows:
    if len(r) >= 2 and r[0] == "File Path":
        cur = r[1]
        continue
    if len(r) > 2 and r[0] == "Line No":
        hdr = r
        continue
    if hdr and len(r) > 10 and r[0] != "":
        try:
            ln = int(r[0])
        except ValueError:
            continue
        iE, iS = hdr.index("Instructions Executed"), hdr.index("# Samples")
        agg[(cur.split('/')[-1], ln)] = (num(r[iE]), num(r[iS]), r[1][:110])
tot = sum(v[0] for v in agg.values()) or 1
tots = sum(v[1] for v in agg.values()) or 1
print('total inst', tot, 'samples', tots)
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:top]:
    print('%-16s %4d %6.2f%% inst %6.2f%% smp  %s' % (k[0], k[1], 100 * v[0] / tot, 100 * v[1] / tots, v[2]))
