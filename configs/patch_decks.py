"""Patch the reference's stale-format decks so that the CURRENT parser accepts them.

The reference's read_input (driver_io.f90:142-153,159) needs on line 10
`beta  MoenchM  [alphas...]` and on line 11 seven fields ending `MNtype order`.
Several shipped decks predate that format: line 10 has only
beta and line 11 lacks MNtype and/or order, so a list-directed read would hit the
`::` comment token and abort.  This script rewrites only those two lines:
  line 10: `<beta>  0`                      (no Moench alphas)
  line 11: first five numbers kept, then ` 2 5` (MNtype=2, order=5: unused by models 0-5)
Run once on fresh copies of the reference decks; the result is what is committed.
"""
import re
import sys

NUM = r"[-+]?(?:\d+\.?\d*|\.\d+)(?:[DdEe][-+]?\d+)?"


def patch(path):
    lines = open(path).read().split("\n")
    changed = []
    # line 10
    head, sep, tail = lines[9].partition("::")
    toks = head.split()
    if len(toks) == 1:
        lines[9] = f"{toks[0]}  0                      {sep}{tail}  [patched: MoenchM=0 added]"
        changed.append(10)
    head, sep, tail = lines[10].partition("::")
    toks = head.split()
    if len(toks) < 7:
        lines[10] = "  ".join(toks[:5]) + "  2 5   " + sep + tail + "  [patched: MNtype, order added]"
        changed.append(11)
    if changed:
        open(path, "w").write("\n".join(lines))
    return changed


if __name__ == "__main__":
    for p in sys.argv[1:]:
        print(p, "patched lines", patch(p))
