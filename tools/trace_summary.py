#!/usr/bin/env python
"""Phase summary of a K2 timeline trace (tools/tc3_trace.py): python tools/trace_summary.py trace.bin [tile]"""
import sys
import numpy as np
r = np.fromfile(sys.argv[1], dtype=np.uint64); r = r[r != 0]
tag = (r >> np.uint64(56)).astype(int); a = ((r >> np.uint64(48)) & np.uint64(0xff)).astype(int); b = ((r >> np.uint64(40)) & np.uint64(0xff)).astype(int)
t = (r & np.uint64(0xffffffffff)).astype(np.int64)
o = np.argsort(t, kind='stable'); tag, a, b, t = tag[o], a[o], b[o], t[o]
names = {1: "mma acc_free ok", 2: "mma L1 chunk full", 3: "mma L2 chunk full", 4: "mma layer committed", 5: "epi acc_full ok", 6: "epi chunk published",
         7: "epi acc_free sent", 8: "prod loads issued", 9: "prod stage empty ok", 10: "prod stage published"}
starts = t[(tag == 1) & (a == 0)]
print('tiles', len(starts), 'median period', np.median(np.diff(starts)))
c4 = t[(tag == 4) & (a == 0)]; c41 = t[(tag == 4) & (a == 1)]
f20 = t[(tag == 2) & (a == 0)]; f30 = t[(tag == 3) & (a == 0)]
n = min(len(c4), len(c41), len(f20), len(f30), len(starts)) - 1
print('acc_free ok -> L1 first chunk       ', np.median(f20[:n] - starts[:n]))
print('L1 first chunk -> L1 committed      ', np.median(c4[:n] - f20[:n]))
print('L1 committed -> L2 first chunk      ', np.median(f30[:n] - c4[:n]))
print('L2 first chunk -> L2 committed      ', np.median(c41[:n] - f30[:n]))
print('L2 committed -> next tile acc_free ok', np.median(starts[1:n + 1] - c41[:n]))
for tg in (2, 3, 6, 10):
    tt = t[tag == tg]; print(names[tg], 'median gap', np.median(np.diff(tt)))
# producer: empty ok -> published (consume body), issued -> next issued per group
for g in (0, 1):
    e = t[(tag == 9) & (b == g)]; p = t[(tag == 10) & (b == g)]; i = t[(tag == 8) & (b == g)]
    m = min(len(e), len(p))
    print(f'group {g}: consume body (empty ok -> published) median', np.median(p[:m] - e[:m]), ' chunk period median', np.median(np.diff(p)), 'mean', np.mean(np.diff(p)))
if len(sys.argv) > 2:
    k0 = int(sys.argv[2]); lo, hi = starts[k0], starts[k0 + 1] + 300
    for k in np.nonzero((t >= lo) & (t < hi))[0]:
        print(f"{t[k]-lo:8d}  {names[tag[k]]:22s} chunk={a[k]} grp={b[k]}")
