"""Summarise an `ncu --page source --csv` export: dynamic opcode mix, stall reasons, hot code regions."""
import collections
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
hdr, data = rows[1], rows[2:]
ix = {h: i for i, h in enumerate(hdr)}
ex = [int(r[ix['Instructions Executed']] or 0) for r in data]
sm = [int(r[ix['# Samples']] or 0) for r in data]
tot, ts = sum(ex), sum(sm)


def opcode(r):
    parts = r[ix['Source']].split()
    op = parts[1] if parts[0].startswith('@') else parts[0]
    return op.split('.')[0].rstrip(';')


mix = collections.Counter()
for r, n in zip(data, ex):
    mix[opcode(r)] += n
print("executed warp-instructions: %d   samples: %d" % (tot, ts))
print("opcode mix: " + "  ".join("%s %.1f%%" % (k, 100 * v / tot) for k, v in mix.most_common(14)))
keys = [k for k in hdr if k.startswith('stall_') and 'Not Issued' not in k]
st = {k: sum(int(r[ix[k]] or 0) for r in data) for k in keys}
print("stalls: " + "  ".join("%s %.1f%%" % (k[6:], 100 * v / max(1, sum(st.values()))) for k, v in sorted(st.items(), key=lambda x: -x[1])[:7]))
start = 0
for i in range(1, len(ex) + 1):
    if i == len(ex) or abs(ex[i] - ex[start]) > 0.02 * max(ex[start], 1):
        seg = sum(ex[start:i])
        if seg > 0.01 * tot:
            ops = collections.Counter(opcode(r) for r in data[start:i])
            print("  instr %5d-%5d (n=%4d) x %10d = %5.1f%% instr, %5.1f%% samples  %s" % (
                start, i - 1, i - start, ex[start], 100 * seg / tot, 100 * sum(sm[start:i]) / ts,
                " ".join("%s:%d" % kv for kv in ops.most_common(6))))
        start = i
