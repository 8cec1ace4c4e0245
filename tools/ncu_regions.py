#!/usr/bin/env python
"""Where the warp instructions of a traversal kernel go: splits the SASS of one captured launch (ncu --set full
--import-source on) at the warp votes that open the phases of trace_engine.cuh's loop — queue fetch, bookkeeping
(promote / pop / finish + sink), node phase, triangle phase — and prints per phase the share of executed warp
instructions, the average active threads per instruction and the share of stall samples.
usage: ncu_regions.py <file.ncu-rep> [launch index = 0] [out.txt]"""
import csv
import io
import subprocess
import sys


def main():
    rep = sys.argv[1]
    launch = int(sys.argv[2]) if len(sys.argv) > 2 else 0
    raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass", "--launch-skip", str(launch), "--launch-count", "1"],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    name = rows[0][1] if rows and len(rows[0]) > 1 else "?"
    hdr = rows[1]
    ix = {h: i for i, h in enumerate(hdr)}
    body = []
    for r in rows[2:]:
        if r and r[0] == "Kernel Name":
            break
        if len(r) > 10:
            body.append(r)
    ie = [int(r[ix["Instructions Executed"]]) for r in body]
    te = [int(r[ix["Thread Instructions Executed"]]) for r in body]
    ss = [int(r[ix["# Samples"]]) for r in body]
    src = [r[ix["Source"]].strip() for r in body]
    votes = [i for i, s in enumerate(src) if s.startswith("VOTE.ANY") or " VOTE.ANY" in s]
    exit_i = next((i for i, s in enumerate(src) if s.startswith("EXIT")), len(body))
    names = ["prologue (mask tables)", "queue fetch + ray set-up", "bookkeeping (promote, pop, finish, sink)", "node phase (8 child boxes)"]
    cuts = [0] + votes[:4] + [exit_i, len(body)]
    labels = names[:len(votes[:4])] + ["triangle phase (Moeller-Trumbore, candidate, exact keys)", "out-of-line subroutines"]
    tot_i, tot_s = max(1, sum(ie)), max(1, sum(ss))
    out = ["# %s, launch %d: %s" % (rep, launch, name),
           "# %d SASS instructions, %d warp instructions executed, %d loop iterations (executions of the first vote)" % (len(body), sum(ie), ie[votes[0]] if votes else 0),
           "%-58s %8s %10s %12s %9s" % ("phase", "SASS", "warp inst", "threads/inst", "samples")]
    for k in range(len(cuts) - 1):
        a, b = cuts[k], cuts[k + 1]
        i_, t_, s_ = sum(ie[a:b]), sum(te[a:b]), sum(ss[a:b])
        out.append("%-58s %8d %9.1f%% %12.1f %8.1f%%" % (labels[k] if k < len(labels) else "?", b - a, 100.0 * i_ / tot_i, t_ / max(1, i_), 100.0 * s_ / tot_s))
    txt = "\n".join(out) + "\n"
    if len(sys.argv) > 3:
        open(sys.argv[3], "w").write(txt)
    print(txt)


if __name__ == "__main__":
    main()
