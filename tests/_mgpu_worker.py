"""Worker of tests/test_multigpu.py: one rank of a world-size-N run (torch.distributed.run, one process per GPU).

Every rank owns a shard of targets (its own catalog seed) and the replicated sensor network, sets up the
peer exchange through the C-ABI (pc_comm_init / pc_comm_attach over CUDA IPC; torch.distributed only carries
the 64-byte handles), steps, and checks ON EVERY RANK that what the peers pushed into its gathered buffer is
byte-identical to an NCCL all-gather of the same summary buffers -- after graph-replayed steps (the push of
step k rides in step k + 1's sensor kernel; the check first exchanges the current summaries), for the four-lane
and the thread-per-target kernel, a slot size that is not a multiple of 16 bytes, the host-facing step (zero-copy
output block together with the peer exchange) and the stand-alone pc_allgather_step; and that, WITHOUT an
explicit exchange, the gathered buffer after step k holds exactly the summaries of step k - 1."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from clockpunch_b200 import _lib  # noqa: E402
from clockpunch_b200.engine import CatalogEngine  # noqa: E402
from clockpunch_b200.simulation.catalog import synth_catalog  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    checks = 0

    def gathered_reference(eng):
        out = torch.empty((world, eng.summary.numel()), dtype=torch.uint8, device=dev)
        dist.all_gather_into_tensor(out, eng.summary.contiguous())
        return out

    def C_status(eng):
        """pc_peer_status without an exchange."""
        import ctypes as C
        err, n = C.c_uint32(0), C.c_uint64(0)
        eng.ctx.check(eng.ctx.lib.pc_peer_status(eng.ctx.handle, C.byref(err), C.byref(n)))
        return err.value, n.value

    def check(eng, gathered, what):
        nonlocal checks
        err, pushes = eng.peer_status()                   # waits for every peer's latest push
        assert err == 0, (what, rank, hex(err))
        ref = gathered_reference(eng)
        got = gathered[pushes & 1]
        for r in range(world):
            assert torch.equal(got[r], ref[r]), (what, f"rank {rank}: slot of rank {r} differs", pushes)
        f = eng.peer_fields(pushes & 1, (rank + 1) % world)
        assert f["num_tasked"].shape == (eng.N,)
        checks += 1
        return pushes

    for n, name in ((3000, "quad kernel"), (21, "tiny shard: slot size not a multiple of 16 bytes"),
                    (40_000, "thread-per-target kernel")):
        cat = synth_catalog(n, 70, seed=900 + 17 * rank + n, sensor_seed=5)
        eng = CatalogEngine(cat.sensors, cat.sensor_kinds, cat.targets, cat.est_x, cat.est_p, cat.Q, cat.R,
                            device=local, seed=40 + rank)
        gathered = eng.enable_peer_push()
        g = eng.build_graph(100.0, policy_probes=32, policy_seed=3)
        last = 0
        for s in range(6):
            if s == 3:
                eng.reset()                               # an episode boundary must not rewind the push counter
            eng.replay(g)
            if s in (0, 2, 5):                            # steps 1, 3..4 run back to back: the exchange is the next step's
                pushes = check(eng, gathered, f"{name}, step {s}")
                assert pushes > last                      # one push per step + one per explicit exchange
                last = pushes
        assert int(eng.num_tasked.sum().item()) > 0
        # without an explicit exchange: after one more step the gathered buffer holds the PREVIOUS step's summaries
        before = gathered_reference(eng)                  # everybody's summaries now (NCCL)
        eng.replay(g)
        torch.cuda.synchronize(dev)
        err, pushes = C_status(eng)
        assert err == 0
        for r in range(world):
            assert torch.equal(gathered[pushes & 1][r], before[r]), (name, "deferred push", rank, r)
        # host-facing step: zero-copy output block together with the peer exchange
        eng._graph = None
        hb = eng.host_buffers()
        for s in range(3):
            eng.step_host(hb, 100.0, policy_probes=32, policy_seed=3)
            torch.cuda.synchronize(dev)
            assert torch.equal(hb["block"][eng._obs_bytes:], eng.summary.cpu()), (name, "host block", s)
            np.testing.assert_array_equal(hb["obs"].numpy()[6 * (eng.M + eng.N): 6 * (eng.M + eng.N) + 36 * eng.N],
                                          eng.est_p.cpu().numpy().astype(np.float32).ravel())
        check(eng, gathered, f"{name}, step_host")
        # stand-alone exchange of an arbitrary summary buffer
        eng.summary[:8] = torch.arange(8, dtype=torch.uint8, device=dev) + rank
        c = eng.ctx
        c.check(c.lib.pc_allgather_step(c.handle, eng.summary.data_ptr(), torch.cuda.current_stream(dev).cuda_stream))
        torch.cuda.synchronize(dev)
        err, pushes = C_status(eng)
        ref = gathered_reference(eng)
        assert err == 0 and all(torch.equal(gathered[pushes & 1][r], ref[r]) for r in range(world)), (name, "pc_allgather_step")
        checks += 1
        dist.barrier()
    torch.cuda.synchronize(dev)
    dist.barrier()
    print(f"MGPU_OK rank {rank} of {world}: {checks} gathered-buffer checks", flush=True)
    os._exit(0)


if __name__ == "__main__":
    main()
