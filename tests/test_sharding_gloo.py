"""world_size = 2 test of the batch-sharded path on CPU (gloo), with the oracle standing in for the CUDA engine:
the gathered output of two ranks equals the single-process output bit for bit, including the noise stream."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import sf_oracle as so
from oracle._refimport import make_cfg
from streamingflow_b200.sharding import shard_bounds


def test_shard_bounds_cover_the_batch():
    for n in (1, 2, 7, 8, 16, 64):
        for w in (1, 2, 3, 4, 8):
            b = [shard_bounds(n, r, w) for r in range(w)]
            assert b[0][0] == 0 and b[-1][1] == n and all(b[i][1] == b[i + 1][0] for i in range(w - 1))
            assert max(h - l for l, h in b) - min(h - l for l, h in b) <= 1


def _build(C=8, H=16, B=3, seed=4):
    from streamingflow_b200.models.future_prediction_ode import FuturePredictionODE
    from tests._oracle_backend import OracleBackend

    m = FuturePredictionODE(C, C, 4, make_cfg(C)).eval().double()
    m.load_state_dict(so.recipe_state_dict({k: tuple(v.shape) for k, v in m.state_dict().items()}, seed, 1.0, torch.float64), strict=True)
    m.gru_ode.__dict__["_engine_factory"] = lambda sd, h, w, n, prec, dev: OracleBackend(sd, h, w, n, prec, dev, torch.float64)
    ct = torch.tensor([[-1.0, -0.5, 0.0], [-1.013, -0.492, -0.004], [-0.99, -0.51, 0.0]], dtype=torch.float64)
    lt = torch.tensor([[-0.8, -0.6, -0.4, -0.2, 0.0], [-0.81, -0.6, -0.418, -0.2, 0.011], [-0.8, -0.62, -0.4, -0.2, 0.0]], dtype=torch.float64)
    tt = torch.tensor([[-1.0, 0.0, 1.0, 2.0]] * B, dtype=torch.float64)
    cam = so.recipe_array("cam", (B, 3, C, H, H), seed, torch.float64)
    lid = so.recipe_array("lidar", (B, 5, C, H, H), seed, torch.float64)
    return m, (torch.zeros(B, 1, C, H, H, dtype=torch.float64), cam, lid, ct, lt, tt)


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from streamingflow_b200.sharding import sharded_forward

    torch.set_num_threads(2)
    m, args = _build()
    torch.manual_seed(99)
    with torch.no_grad():
        x, aux = sharded_forward(m, *args)
        x2, _ = sharded_forward(m, *args)         # second call: every rank must have consumed the WHOLE batch's noise stream
    if rank == 0:
        np.save(out, torch.stack([x, x2]).numpy())
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 4])
def test_sharded_forward_equals_single_process(tmp_path, world):
    """world = 2: shards of 2 + 1 samples; world = 4 > B = 3: the last rank's shard is empty.  Two consecutive calls."""
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    out = str(tmp_path / "x.npy")
    mp.spawn(_worker, args=(world, port, out), nprocs=world, join=True)
    m, args = _build()
    torch.manual_seed(99)
    with torch.no_grad():
        ref = torch.stack([m(*args)[0], m(*args)[0]])
    assert not torch.equal(ref[0], ref[1])           # the second call sees later noise
    got = torch.from_numpy(np.load(out))
    # batch-size dependent blocking in the CPU conv kernels moves the last bits of the torch encoder / decoder; the noise
    # stream and the schedule must match exactly, which a 1e-12 bound on energised weights demonstrates (a shifted noise
    # tape changes the output at the 1e-1 level)
    assert got.shape == ref.shape
    err = ((got - ref).abs().max() / ref.abs().max()).item()
    assert err < 1e-12, err


def _halo_worker(rank, world, port, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from streamingflow_b200.row_sharding import HALO, exchange_halo_rows, local_window

    h, w, B, C = 60, 5, 2, 3
    full = torch.arange(B * h * w * C, dtype=torch.float32).view(B, h, w, C)
    own_lo, own_hi, lo, hi = local_window(h, rank, world)
    t = torch.full((B, hi - lo, w, C), -1.0)
    t[:, own_lo - lo:own_hi - lo] = full[:, own_lo:own_hi]          # only the band is valid before the exchange
    exchange_halo_rows(t, own_lo, own_hi, lo, hi, rank, world)
    ok = torch.equal(t, full[:, lo:hi])
    open(os.path.join(out_dir, f"ok{rank}"), "w").write(str(int(ok)))
    dist.destroy_process_group()


def test_halo_exchange_fills_the_local_window(tmp_path):
    """Row sharding host logic on CPU (gloo, 3 ranks): after one exchange every rank's local image (band + 12-row halos)
    equals the corresponding rows of the global grid."""
    from streamingflow_b200.row_sharding import local_window

    assert local_window(400, 0, 8) == (0, 50, 0, 62) and local_window(400, 3, 8) == (150, 200, 138, 212)
    assert local_window(400, 7, 8) == (350, 400, 338, 400)
    with pytest.raises(ValueError):
        local_window(50, 1, 8)
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    mp.spawn(_halo_worker, args=(3, port, str(tmp_path)), nprocs=3, join=True)
    assert all(open(tmp_path / f"ok{r}").read() == "1" for r in range(3))


def test_row_sharded_event_program_splits_at_the_nccl_calls():
    """The op list of one row-sharded event: stage launches, SE reduce -> all-reduce -> apply, then pack -> send/recv -> unpack;
    the graph replay captures the runs between the NCCL calls ('allreduce', 'p2p') as segments."""
    from types import SimpleNamespace as NS
    from streamingflow_b200 import _lib as L
    from streamingflow_b200.row_sharding import RowShardedOde

    eng = NS(cell_slots=[[0, 1, 2, 3, 4, 5], [6, 7, 8, 9, 10, 11]], prior_items=[12, 13, L.SE_ITEM_BASE, 14, 15, L.SE_ITEM_BASE + 1, 16])
    ev = NS(kind=1, run_cell=1, run_prior=1)
    ops = RowShardedOde._event_ops(NS(eng=eng, world=4, transport="nccl"), ev)
    kinds = [k for k, _ in ops]
    assert kinds == ["stage"] * 8 + ["se_reduce", "allreduce", "se_apply"] + ["stage"] * 2 + ["se_reduce", "allreduce", "se_apply", "stage",
                                                                                      "pack", "p2p", "unpack"]
    assert [a for k, a in ops if k == "stage"][:6] == eng.cell_slots[1]
    segments, cur = [], []
    for k, _ in ops + [("flush", 0)]:
        if k in ("allreduce", "p2p", "flush"):
            if cur:
                segments.append(cur)
            cur = []
        else:
            cur.append(k)
    assert len(segments) == 4 and segments[-1] == ["unpack"] and segments[2][-1] == "pack"
    # peer transport: the same event with every exchange a kernel (nothing splits the graph)
    peer = [k for k, _ in RowShardedOde._event_ops(NS(eng=eng, world=4, transport="peer"), ev)]
    assert peer == ["stage"] * 8 + ["se_reduce", "peer_allreduce", "se_apply"] + ["stage"] * 2 + ["se_reduce", "peer_allreduce", "se_apply", "stage",
                                                                                         "push", "pull"]
    single = RowShardedOde._event_ops(NS(eng=eng, world=1, transport="none"), NS(kind=0, run_cell=1, run_prior=0))
    assert [k for k, _ in single] == ["stage"] * 6          # one rank: no collectives, no halo exchange
