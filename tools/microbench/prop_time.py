"""Time the standalone propagate kernel (thread per state) at several sizes."""
import sys, os, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np, torch
from clockpunch_b200 import _lib
from clockpunch_b200.simulation.catalog import synth_catalog
ctx = _lib.get_ctx(0)
dev = torch.device("cuda", 0)
for n in (10_000, 140_000, 1_400_000):
    cat = synth_catalog(n, 1, seed=1)
    x = torch.from_numpy(cat.targets).to(dev)
    y = torch.empty_like(x)
    st = torch.cuda.current_stream().cuda_stream
    for sub in (0, 1, 2):
        for _ in range(3):
            ctx.check(ctx.lib.pc_propagate_twobody(ctx.handle, x.data_ptr(), y.data_ptr(), n, n, 0.0, 100.0, sub, 0, st))
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(20):
            ctx.check(ctx.lib.pc_propagate_twobody(ctx.handle, x.data_ptr(), y.data_ptr(), n, n, 0.0, 100.0, sub, 0, st))
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 20
        print(f"n={n} substeps={sub}: {ms*1e3:.1f} us  -> {n/ms/1e3:.1f} M state-steps/s")
