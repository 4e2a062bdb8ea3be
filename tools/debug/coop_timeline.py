"""In-pipeline timeline of the per-target kernel (variant library built with -DPC_STAMPS_HEADER):
   tools/build_variant.sh stamps -DPC_STAMPS_HEADER
   PUNCH_B200_LIB=$PWD/variants/libpunch_stamps.so python tools/debug/coop_timeline.py"""
import sys, os, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np, torch
from clockpunch_b200.engine import CatalogEngine
from clockpunch_b200.simulation.catalog import synth_catalog
SEED = 20240627
N = int(sys.argv[1]) if len(sys.argv) > 1 else 10000
cat = synth_catalog(N, 100, seed=SEED, sensor_seed=SEED)
eng = CatalogEngine(cat.sensors, cat.sensor_kinds, cat.targets, cat.est_x, cat.est_p, cat.Q, cat.R, device=0, seed=SEED)
lib = eng.ctx.lib
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
g = eng.build_graph(100.0, policy_probes=64, policy_seed=SEED)
eng.reset()
for _ in range(5):
    eng.replay(g)
names = ["", "loads landed", "rng done", "transform, pred_p", "chol pred_p", "S, chol S (end part 1)", "gain, xe, est_p",
         "stores, summaries", "chol est_p, sigma writes", "vis row (end)"]
acc = np.zeros(10); gts = []; pts = []
for it in range(20):
    flush.zero_(); torch.cuda.synchronize()
    lib.pc_debug_stamps(None, None, 1)
    lib.pc_debug_prop_stamps(None, 1)
    eng.replay(g); torch.cuda.synchronize()
    cst = np.zeros((12, 256), np.int64); gt = np.zeros(8, np.uint64)
    lib.pc_debug_stamps(cst.ctypes.data_as(C.c_void_p), gt.ctypes.data_as(C.c_void_p), 0)
    pt = np.zeros(4, np.uint64)
    lib.pc_debug_prop_stamps(pt.ctypes.data_as(C.c_void_p), 0)
    pts.append([float(pt[1]) - float(pt[0]), float(gt[0]) - float(pt[0]), float(gt[2]) - float(pt[0])])
    tof = eng.task_of.cpu().numpy()[:100]
    own = [j for j in range(100) if tof[j] >= 0 and tof[j] not in tof[:j]]
    d = np.diff(cst[:10, own], axis=0).mean(axis=1)
    acc[1:] += d
    gts.append([float(gt[1]) - float(gt[0]), float(gt[3]) - float(gt[0]), float(gt[2]) - float(gt[0])])
for i in range(1, 10):
    print(f"  update warp {i} {names[i]:28s} {acc[i] / 20:8.0f} cycles")
gts = np.array(gts)
print("regular path: first entry -> last exit  %.2f us" % (gts[:, 0].mean() / 1e3))
print("update warps: first regular entry -> last update exit  %.2f us (first update entry at %.2f us)" % (gts[:, 1].mean() / 1e3, gts[:, 2].mean() / 1e3))
pts = np.array(pts)
print("propagation: first entry -> last exit %.2f us; first regular entry of the per-target kernel at %.2f us, first update-warp entry at %.2f us"
      % tuple(pts.mean(axis=0) / 1e3))
ct = np.zeros((2048, 4), np.uint64)
lib.pc_debug_prop_ctas(ct.ctypes.data_as(C.c_void_p))
nc = 14 * ((N + 255) // 256)        # kPropWindow targets per CTA
ct = ct[:nc].astype(np.float64)
t0 = ct[:, 0].min()
ent, ld, ex, ns = (ct[:, 0] - t0) / 1e3, (ct[:, 1] - ct[:, 0]) / 1e3, (ct[:, 2] - ct[:, 0]) / 1e3, ct[:, 3]
print("propagation CTAs (last replay): %d; entry times: %.2f us median, %d enter after 2 us (second wave), last entry %.2f us" % (nc, np.median(ent), (ent > 2).sum(), ent.max()))
for k in sorted(set(ns)):
    m = ns == k
    print("  max substeps %d: %4d CTAs, entry -> loads landed %.2f us, entry -> exit %.2f us (max %.2f)" % (k, m.sum(), ld[m].mean(), ex[m].mean(), ex[m].max()))
ps = np.zeros((10, 2048), np.int64)
lib.pc_debug_phase_stamps(ps.ctypes.data_as(C.c_void_p))
w0, w1 = 100 * 4, 100 * 4 + 4 * ((N + 31) // 32)         # warps of the regular CTAs (behind the M update CTAs)
pn = ["", "entry, loads + tile issued", "pair loads + F accumulate", "quad reduce", "m, pred_p", "tasked flag", "outputs (P stores)",
      "chol + sigma writes", "tile wait", "vis words"]
d = np.diff(ps[:, w0:min(w1, 2048)], axis=0)
print("regular warps (last replay), cycles per phase:")
for i in range(1, 10):
    print(f"  {i} {pn[i]:28s} mean {d[i - 1].mean():8.0f}  max {d[i - 1].max():8.0f}")
print("  total mean %.0f" % d.sum(axis=0).mean())
