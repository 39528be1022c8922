#!/usr/bin/env python
"""Summarises the in-kernel timeline of decoder_step3_kernel.

    MOONSHINE_B200_PROF=<step> python bench.py ... 2> prof.txt ; python scripts/prof_v3.py prof.txt

model.cu prints, for a few CTAs, `PROF cta N: tag:ns tag:ns ...` (stamps of thread 0, ns since the CTA's first
stamp).  The time between two consecutive stamps is booked on the LATER tag.  Tags (decoder_step3.cu):
  40+k  dependency wait of a phase of kind k is over      50+k  the CTA's jobs of that phase are done
  SELF  1 prologue+LN  2 qkv MMA  3 rope/append/attention  4 per-head O partial
  CROSS 11 prologue+LN 12 q MMA 13 attention
  GEMM  20+k x planes built   30+k MMA + epilogue done     (k: 2 OC, 3 FC1, 4 FC2)
  LOGITS 35 x planes  36 MMAs issued  37 accumulator ready  38 epilogue
"""
import collections
import re
import sys

NAMES = {0: "start", 1: "S.prologue", 2: "S.qkv", 3: "S.attn", 4: "S.wo", 11: "X.prologue", 12: "X.q", 13: "X.attn",
         22: "OC.x", 23: "FC1.x", 24: "FC2.x", 32: "OC.mma", 33: "FC1.mma", 34: "FC2.mma",
         60: "s.flags", 61: "s.resolve", 62: "s.ln", 63: "g.sync", 64: "g.issue", 65: "g.acc", 66: "a.scores", 67: "a.softmax",
         68: "a.pv", 69: "o.acq", 70: "x.scores", 71: "x.softmax", 72: "x.pv", 73: "f.resolve", 74: "g.ring0",
         35: "L.x", 36: "L.issue", 37: "L.acc", 38: "L.epi",
         40: "wait>SELF", 41: "wait>CROSS", 42: "wait>OC", 43: "wait>FC1", 44: "wait>FC2", 45: "wait>FINAL", 46: "wait>LOGITS",
         50: "end.SELF", 51: "end.CROSS", 52: "end.OC", 53: "end.FC1", 54: "end.FC2", 55: "end.FINAL", 56: "end.LOGITS"}


NAMES4 = {0: "start", 1: "embed", 10: "qkv", 11: "o", 12: "qc", 13: "oc", 14: "fc1", 15: "fc2", 20: "selfattn+CB", 21: "h+CB", 22: "cross+CB",
          23: "h+CB'", 24: "silu", 25: "CB", 30: "final+handoff", 31: "logits", 40: "rope", 41: "a.scores", 43: "a.pv", 44: "a.bcast",
          45: "x.scores", 46: "x.softmax", 47: "x.pv", 48: "CB(part)", 49: "reduce", 50: "g.acq", 51: "g.rows", 52: "g.sync"}


def main(path, v4=False):
    global NAMES
    if v4:
        NAMES = NAMES4
    scale = 1.0 / 1.965 if v4 else 1.0   # v4 stamps are raw SM cycles
    for line in open(path):
        m = re.match(r"PROF cta (\d+):(.*)", line)
        if not m:
            continue
        stamps = [(int(a), int(b) * scale) for a, b in re.findall(r"(\d+):(\d+)", m.group(2))]
        if len(stamps) < 2:
            continue
        tot = collections.OrderedDict()
        cnt = collections.Counter()
        for (t0, ns0), (t1, ns1) in zip(stamps, stamps[1:]):
            tot[t1] = tot.get(t1, 0) + (ns1 - ns0)
            cnt[t1] += 1
        total = stamps[-1][1] - stamps[0][1]
        print(f"PROF cta {m.group(1)} total {total / 1000:.1f} us")
        print("   " + "  ".join(f"{NAMES.get(t, t)}={v / 1000 / cnt[t]:.2f}x{cnt[t]}" for t, v in tot.items()))


if __name__ == "__main__":
    main(sys.argv[1], v4="--v4" in sys.argv)
