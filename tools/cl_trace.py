#!/usr/bin/env python
"""Developer tool: in-kernel timeline of the cluster probe (library built with -DHJ3D_CL_TRACE, see probe_cluster.cuh).

    HJ3D_LIB=tools/libhj3d_trace.so python tools/cl_trace.py
"""
import ctypes as C
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    import hj3d_loader
    pkg = hj3d_loader.load()
    lib = pkg.capi.load()
    dev = torch.device("cuda", 0)
    nR, nS = 1 << 24, 1 << 27
    g = torch.Generator(device=dev); g.manual_seed(1)
    R = torch.zeros((nR, 3), dtype=torch.int32, device=dev)
    R[:, 0] = torch.randperm(nR, device=dev, generator=g, dtype=torch.int64).to(torch.int32)
    S = torch.zeros((nS, 3), dtype=torch.int32, device=dev)
    S[:, 0] = torch.arange(nS, device=dev, dtype=torch.int64).to(torch.int32)
    S[:, 1] = torch.randint(0, nR, (nS,), device=dev, generator=g, dtype=torch.int64).to(torch.int32)
    ctx = pkg.Context(0, stream=torch.cuda.current_stream().cuda_stream)
    t = ctx.table(pkg.CHAINING, nR).build(R, nR, pkg.KeySpec(12, 0))
    out = torch.empty((nS, 2), dtype=torch.int32, device=dev)
    rows = np.zeros((4096, 8), np.int64)
    n = C.c_uint(0)
    for rep in range(2):
        rc, c = t.probe_chaining(S, nS, pkg.KeySpec(12, 4), unique=True, out=out, out_cap=nS)
        lib.hj3d_debug_trace_read(rows.ctypes.data_as(C.c_void_p), C.byref(n))
    print("probe ms", ctx.timings()["probe_ms"], "rows", n.value)
    r = rows[: n.value]
    t0 = r[:, 3].min()
    rt = r[r[:, 0] == 0]; pr = r[(r[:, 0] > 0) & (r[:, 0] < 100)]
    r7 = r[r[:, 0] == 100]; r7 = r7[np.argsort(r7[:, 1])]
    print("router warp 7 (k: start, ->after wait, ->ranked, ->bar1, ->bar2)")
    for x in r7[:12]:
        print(int(x[1]), int(x[3] - t0), *(int(x[4 + i] - x[3 + i]) for i in range(4)))
    rt = rt[np.argsort(rt[:, 1])]
    print("router rounds (k: start, empty-wait, rank, place, publish+fetch) in cycles")
    for x in rt[:12]:
        print(int(x[1]), int(x[3] - t0), *(int(x[4 + i] - x[3 + i]) for i in range(4)))
    d = np.diff(rt[:, 3])
    print("router round period: median", np.median(d), "mean", d.mean())
    for i, name in enumerate(["empty-wait", "rank", "bar1", "place->bar2"]):
        print("  router", name, "median", np.median(rt[:, 4 + i] - rt[:, 3 + i]), "mean", (rt[:, 4 + i] - rt[:, 3 + i]).mean())
    print("prober tasks:", len(pr))
    for i, name in enumerate(["full-wait", "runs+pull+probe", "atomic", "write"]):
        v = pr[:, 4 + i] - pr[:, 3 + i]
        print("  prober", name, "median", np.median(v), "mean", v.mean(), "max", v.max())
    pr = pr[np.argsort(pr[:, 3])]
    for x in pr[:40]:
        print("warp", int(x[0]) - 1, "k", int(x[1]), "src", int(x[2]), "t0", int(x[3] - t0), *(int(x[4 + i] - x[3 + i]) for i in range(4)))


if __name__ == "__main__":
    main()
