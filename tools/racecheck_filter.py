#!/usr/bin/env python3
"""Splits a `compute-sanitizer --tool racecheck` log into hazards that involve a TMA bulk-copy write (cp.async.bulk issued by
bulk_g2s / bulk_g2s_e: racecheck does not model the mbarrier complete_tx -> try_wait ordering that protects them; that protocol is
exercised by tests/test_gpu_ring_stress.py instead) and ALL OTHER hazards, which must be zero.

    compute-sanitizer --tool racecheck --racecheck-report all python -m pytest ... 2>&1 | tee race.log
    python tools/racecheck_filter.py race.log > profiles/r2_compute_sanitizer.txt     (exit status 1 if other hazards exist)"""
import re
import sys


def main():
    text = open(sys.argv[1], errors="replace").read()
    recs = re.split(r"(?m)^=========\s+(?=(?:Error|Warning): (?:Race|Potential))", text)
    head, recs = recs[0], recs[1:]
    tma, other = {}, {}
    for r in recs:
        first = r.strip().splitlines()[0]
        key = re.sub(r"\s+", " ", " | ".join(l.strip("= ").strip() for l in r.splitlines()[:2]))
        key = re.sub(r"0x[0-9a-f]+", "0x..", key)
        (tma if "bulk_g2s" in r else other)[key] = (tma if "bulk_g2s" in r else other).get(key, 0) + 1
    summary = re.findall(r"RACECHECK SUMMARY:.*", text)
    print("racecheck log: %s" % sys.argv[1])
    for s in summary:
        print("  " + s)
    print("hazard records involving a TMA bulk-copy write (bulk_g2s / bulk_g2s_e; protocol covered by tests/test_gpu_ring_stress.py): %d"
          % sum(tma.values()))
    for k, n in sorted(tma.items(), key=lambda kv: -kv[1]):
        print("  %5d x %s" % (n, k[:260]))
    print("ALL OTHER hazard records: %d" % sum(other.values()))
    for k, n in sorted(other.items(), key=lambda kv: -kv[1]):
        print("  %5d x %s" % (n, k[:400]))
    passed = re.findall(r"\d+ passed[^\n]*", text)
    for p in passed:
        print("pytest: " + p)
    return 1 if other else 0


if __name__ == "__main__":
    sys.exit(main())
