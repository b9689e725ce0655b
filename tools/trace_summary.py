#!/usr/bin/env python
"""Summarise a chrome trace written by `bench.py --profile`: per kernel the mean duration, and the timeline of one
step (start offset, duration, stream) so that overlap and idle gaps can be read off.

    python tools/trace_summary.py gpurun_out/trace_n2.json [step_marker_kernel]
"""
import collections
import json
import sys


def main(path, marker="peer_gather_send"):
    ev = json.load(open(path))["traceEvents"]
    ks = [e for e in ev if e.get("cat") in ("kernel", "gpu_memset", "gpu_memcpy") and "dur" in e]
    ks.sort(key=lambda e: e["ts"])
    agg = collections.OrderedDict()
    for e in ks:
        agg.setdefault(e["name"][:60], []).append(e["dur"])
    print("%-62s %4s %9s %9s" % ("kernel", "n", "mean us", "total us"))
    for n, v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
        print("%-62s %4d %9.1f %9.1f" % (n, len(v), sum(v) / len(v), sum(v)))
    starts = [e["ts"] for e in ks if marker in e["name"]]
    if len(starts) >= 3:
        t0, t1 = starts[1], starts[2]
        print("\none step (%s -> next %s): %.1f us" % (marker, marker, t1 - t0))
        for e in ks:
            if t0 <= e["ts"] < t1:
                print("  +%7.1f  %7.1f us  stream %-4s %s" % (e["ts"] - t0, e["dur"], e.get("args", {}).get("stream", "?"),
                                                             e["name"][:70]))


if __name__ == "__main__":
    main(*sys.argv[1:])
