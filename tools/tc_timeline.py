"""Decode the one-tile timeline written by DISN_TC_TRACE=1 DISN_TC_TIMELINE=<file> (point_tc.cu, TL stamps)."""
import sys
ev = {}
for ln in open(sys.argv[1]):
    i, t = ln.split()
    if int(t):
        ev[int(i)] = int(t)
t0 = min(ev.values())
g = lambda i: (ev[i] - t0) if i in ev else None
names = []
for sidx in range(2):
    for sl in range(21):
        layer = 0 if sl == 0 else 1 if sl < 5 else 2 if sl < 13 else 3
        names.append((g(sidx * 21 + sl), "MMA  s%d L%d slice %2d acquired" % (sidx, layer, sl)))
        names.append((g(64 + sidx * 21 + sl), "MMA  s%d L%d slice %2d issued+released" % (sidx, layer, sl)))
    for layer in range(4):
        names.append((g(128 + sidx * 4 + layer), "MMA  s%d L%d committed (acc_full)" % (sidx, layer)))
    for eg in range(2):
        names.append((g(400 + sidx * 8 + eg * 4), "EPI%d s%d final: acc_full[3] seen" % (eg, sidx)))
        names.append((g(400 + sidx * 8 + eg * 4 + 1), "EPI%d s%d final: done" % (eg, sidx)))
for q in range(40):
    sidx, r = divmod(q, 20)
    layer, t = (0, r) if r < 4 else (1, r - 4) if r < 12 else (2, r - 12)
    for k, nm in enumerate(("start", "gather ok", "tmem ld done", "slot free", "stored+arrived")):
        names.append((g(192 + q * 5 + k), "EPI%d s%d drain acc%d slice %d: %s" % (t & 1, sidx, layer + 2, t, nm)))
for t, nm in sorted((x for x in names if x[0] is not None)):
    print("%7d  %s" % (t, nm))
