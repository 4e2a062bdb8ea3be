"""Where the e2e step time goes at the bench size (host clock)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np, torch
from clockpunch_b200.engine import CatalogEngine
from clockpunch_b200.simulation.catalog import synth_catalog
SEED = 20240627
N, M = int(sys.argv[1]) if len(sys.argv) > 1 else 10000, 100
cat = synth_catalog(N, M, seed=SEED, sensor_seed=SEED)
eng = CatalogEngine(cat.sensors, cat.sensor_kinds, cat.targets, cat.est_x, cat.est_p, cat.Q, cat.R, device=0, seed=SEED)
hb = eng.host_buffers()
hb["actions"].fill_(N)
def loop(fn, n=200):
    for _ in range(20): fn()
    t = time.perf_counter()
    for _ in range(n): fn()
    return (time.perf_counter() - t) / n * 1e6
full = loop(lambda: eng.step_host(hb, 100.0))
print(f"step_host (H2D actions, graph, D2H obs+vis+nt, sync): {full:.1f} us")
hb2 = dict(hb)
import ctypes as C
# variants: drop outputs one at a time by passing smaller dicts is not supported; time raw copies instead
s = torch.cuda.current_stream()
def d2h_obs():
    hb["obs"].copy_(eng.obs_f32, non_blocking=True); s.synchronize()
def d2h_vis():
    hb["vis_est"].copy_(eng.vis_est, non_blocking=True); s.synchronize()
def d2h_nt():
    hb["num_tasked"].copy_(eng.num_tasked, non_blocking=True); s.synchronize()
def sync_only():
    s.synchronize()
print(f"D2H obs ({hb['obs'].numel()*4/1e6:.2f} MB) + sync: {loop(d2h_obs):.1f} us; vis: {loop(d2h_vis):.1f} us; num_tasked: {loop(d2h_nt):.1f} us; bare sync {loop(sync_only):.1f} us")
g = eng.build_graph(100.0, pack_obs=True)
def graph_only():
    eng.replay(g); s.synchronize()
print(f"graph replay (step + obs pack) + sync: {loop(graph_only):.1f} us")
rng = np.random.default_rng(0)
probe_tab = rng.integers(0, N, size=(256, M, 64))
jcol = np.arange(M)[:, None]; jw, jb = jcol >> 5, (jcol & 31).astype(np.uint32); rows = np.arange(M)
h_vis_np = hb["vis_est"].numpy().view(np.uint32); h_act_np = hb["actions"].numpy()
k = [0]
def host_policy():
    probes = probe_tab[k[0] % 256]; k[0] += 1
    hit = (h_vis_np[probes, jw] >> jb) & 1
    first = hit.argmax(axis=1)
    h_act_np[:M] = np.where(hit[rows, first] != 0, probes[rows, first], N)
print(f"host policy (numpy): {loop(host_policy):.1f} us")
