"""Instruction mix and stall-sample share per SASS opcode from an ncu report's source page.

    ncu -i report.ncu-rep --page source --csv | python tools/ncu_mix.py [top]

Reads the CSV on stdin; prints, per opcode, the share of warp-level instructions executed and of stall samples.
"""
import collections
import csv
import sys


def main():
    top = int(sys.argv[1]) if len(sys.argv) > 1 else 30
    rows = list(csv.reader(sys.stdin))
    hdr = next(r for r in rows if "Source" in r and "Instructions Executed" in r)
    i_src, i_inst, i_smp = hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("# Samples")
    inst, smp = collections.Counter(), collections.Counter()
    for r in rows[rows.index(hdr) + 1:]:
        if len(r) <= max(i_inst, i_smp):
            continue
        src = r[i_src].strip()
        if src.startswith('@'):
            src = src.split(None, 1)[1] if ' ' in src else src
        op = src.split()[0].rstrip(';') if src else '?'
        op = '.'.join(op.split('.')[:2]) if op.startswith(('LDS', 'STS', 'LDG', 'STG', 'LDTM', 'UTC', 'MUFU')) else op.split('.')[0]
        try:
            n, s = int(r[i_inst]), int(r[i_smp])
        except ValueError:
            continue
        inst[op] += n
        smp[op] += s
    tot, ts = sum(inst.values()), max(sum(smp.values()), 1)
    print("warp instructions %d, stall samples %d" % (tot, ts))
    for op, n in inst.most_common(top):
        print("%-14s %6.2f%% inst  %6.2f%% samples" % (op, 100.0 * n / tot, 100.0 * smp[op] / ts))


if __name__ == "__main__":
    main()
