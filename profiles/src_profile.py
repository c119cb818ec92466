#!/usr/bin/env python3
"""Aggregate `ncu --page source --csv --print-source cuda,sass` output by source line and by kernel phase.

usage: ncu -i prof.ncu-rep --page source --csv --print-source cuda,sass > src.csv ; python profiles/src_profile.py src.csv
"""
import collections
import csv
import sys


def load(path):
    rows = list(csv.reader(open(path)))
    agg = collections.OrderedDict()
    cur_file = None
    cur = None
    for r in rows:
        if len(r) >= 2 and r[0] == "File Path":
            cur_file = r[1].split("/")[-1]
            continue
        if len(r) < 9 or r[0] in ("Line No", "Function Name"):
            continue
        if r[0].isdigit():
            cur = (cur_file, int(r[0]), r[1].strip()[:100])
        if r[2] and r[2] not in ("-", "...") and cur is not None:
            try:
                ie, te, sm = int(r[7] or 0), int(r[8] or 0), int(r[4] or 0)
            except ValueError:
                continue
            a = agg.setdefault(cur, [0, 0, 0])
            a[0] += ie; a[1] += te; a[2] += sm
    return agg


def phase_of(f, line, src, marks):
    if f == "amp_kernels.cuh":
        best = "kernels:other"
        for name, lo in marks:
            if line >= lo:
                best = name
        return best
    return f


def main():
    agg = load(sys.argv[1])
    tot = sum(v[0] for v in agg.values()) or 1
    tots = sum(v[2] for v in agg.values()) or 1
    print("total warp instructions %d, stall samples %d" % (tot, tots))
    # phases of amp_kernels.cuh located by their marker comments
    marks = []
    try:
        src = open("amplipy_b200/csrc/amp_kernels.cuh").read().splitlines()
        for i, l in enumerate(src, 1):
            for key, name in (("---- S:", "S stage"), ("---- T:", "T trim+plan"), ("---- W:", "W window"), ("---- C:", "C count"),
                              ("struct TileSink", "T sink"), ("AMP_HD void cta_trim_pileup", "prologue")):
                if key in l:
                    marks.append((name, i))
        marks.sort(key=lambda x: x[1])
    except OSError:
        pass
    ph = collections.Counter(); phs = collections.Counter()
    for (f, l, s), (ie, te, sm) in agg.items():
        k = phase_of(f, l, s, marks)
        ph[k] += ie; phs[k] += sm
    print("\nby phase / file:")
    for k, v in ph.most_common():
        print("  %-22s %5.1f%% inst  %5.1f%% stall samples" % (k, 100 * v / tot, 100 * phs[k] / tots))
    print("\ntop lines by instructions:")
    for (f, l, s), (ie, te, sm) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:int(sys.argv[2]) if len(sys.argv) > 2 else 30]:
        print("  %5.1f%% inst %5.1f%% samp  thr %4.1f  %s:%d  %s" % (100 * ie / tot, 100 * sm / tots, te / max(ie, 1), f, l, s))


if __name__ == "__main__":
    main()
