#!/usr/bin/env python
"""Timeline trace of the pipelined LinkPredictor kernel (cluster 0).

  on the GPU box:  EPS_EXTRA_NVCC_FLAGS=-DEPS_TC3_TRACE python -m edge_proposal_sets_b200.build --force
                   EPS_TC3_TRACE_FILE=gpurun_out/trace.bin python tools/tc3_trace.py run
  anywhere:        python tools/tc3_trace.py show gpurun_out/trace.bin [first_tile n_tiles]
"""
import sys
import numpy as np

TAGS = {1: "mma acc_free ok", 2: "mma chunk full", 3: "mma a2 chunk full", 4: "mma layer committed", 5: "epi acc_full ok",
        6: "epi kb published", 7: "epi acc_free sent", 8: "prod loads issued", 9: "prod stage empty ok", 10: "prod stage published",
        11: "mma chunk MMAs issued", 12: "mma chunk committed", 13: "mma before full wait", 14: "mma after full wait"}


def run():
    import torch
    from edge_proposal_sets_b200 import ops
    n, H, L, M, runlen = 576289, 256, 3, 1 << 23, 14600
    g = torch.Generator().manual_seed(0)
    h = torch.randn(n, H, generator=g).cuda() * 0.3
    owners = M // runlen + 1
    u = torch.randint(0, n, (owners, runlen), generator=g, dtype=torch.int32).sort(dim=1)[0].reshape(-1)[:M].contiguous()
    v = (torch.arange(M) // runlen).to(torch.int32)
    e = torch.stack([u, v]).cuda()
    Ws = [torch.randn(H, H, generator=g).cuda() / 16 for _ in range(L - 1)] + [torch.randn(1, H, generator=g).cuda() / 16]
    bs = [torch.randn(H, generator=g).cuda() / 16 for _ in range(L - 1)] + [torch.zeros(1).cuda()]
    for _ in range(2):
        ops.linkpred_mlp(h, e, Ws, bs, "f16")
    torch.cuda.synchronize()
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0.record(); ops.linkpred_mlp(h, e, Ws, bs, "f16"); t1.record(); torch.cuda.synchronize()
    print("M", M, "ms", t0.elapsed_time(t1))


def show(path, first=20, count=3):
    r = np.fromfile(path, dtype=np.uint64)
    r = r[r != 0]
    tag, a, b = (r >> np.uint64(56)).astype(int), ((r >> np.uint64(48)) & np.uint64(0xff)).astype(int), ((r >> np.uint64(40)) & np.uint64(0xff)).astype(int)
    t = (r & np.uint64(0xffffffffff)).astype(np.int64)
    order = np.argsort(t, kind="stable")
    tag, a, b, t = tag[order], a[order], b[order], t[order]
    print("records", len(r), "span clks", t[-1] - t[0])
    # tile boundaries = 'mma acc_free ok' with layer 0
    starts = t[(tag == 1) & (a == 0)]
    print("first-layer starts:", len(starts), "median period", np.median(np.diff(starts)))
    lo, hi = starts[first], starts[first + count]
    if len(sys.argv) > 5 and sys.argv[5] == "mma":
        keep = np.isin(tag, [1, 2, 3, 4, 11, 12, 13, 14])
        tag, a, b, t = tag[keep], a[keep], b[keep], t[keep]
    for k in np.nonzero((t >= lo) & (t < hi))[0]:
        print(f"{t[k] - lo:8d}  {TAGS[tag[k]]:22s} a={a[k]} b={b[k]}")
    # aggregate gaps
    for name, tg in (("mma chunk full", 2), ("mma a2 kb full", 3), ("epi kb published", 6), ("prod stage published", 10)):
        tt = t[tag == tg]
        print(name, "median gap", np.median(np.diff(tt)), "mean", np.mean(np.diff(tt)))


if __name__ == "__main__":
    if sys.argv[1] == "run":
        run()
    else:
        show(sys.argv[2], *(int(x) for x in sys.argv[3:5]))
