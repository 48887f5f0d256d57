"""world_size-2 (and 3) gloo tests of the halo plan logic on the CPU: partition, remote-row
deduplication, column remap, send/recv lists, and that exchanged rows + remapped columns reproduce
the single-process oracle result.  The pack/unpack kernels themselves are covered by the GPU tests."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, n, e, F, seed, q):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from keras_geometric_b200.dist import HaloPlan, partition_bounds
        from oracle import reference_path as ref
        rng = np.random.default_rng(seed)
        ei = np.stack([rng.integers(0, n, e), rng.integers(0, n, e)]).astype(np.int64)
        x = rng.standard_normal((n, F)).astype(np.float32)
        b = partition_bounds(n, world)
        assert b[0] == 0 and b[-1] == n and all(b[i] <= b[i + 1] for i in range(world))
        lo, hi = b[rank], b[rank + 1]
        mine = (ei[1] >= lo) & (ei[1] < hi)
        plan = HaloPlan(torch.from_numpy(ei[0][mine]), torch.from_numpy(ei[1][mine]), n, rank, world)
        # structure of the plan
        assert plan.n_local == hi - lo and sum(plan.recv_counts) == plan.n_halo
        assert plan.recv_counts[rank] == 0 and plan.send_counts[rank] == 0
        hg = plan.halo_global.numpy()
        assert (np.diff(hg) > 0).all() and ((hg < lo) | (hg >= hi)).all()
        want_halo = np.unique(ei[0][mine][(ei[0][mine] < lo) | (ei[0][mine] >= hi)])
        np.testing.assert_array_equal(hg, want_halo)
        # exchange (host emulation of pack -> all_to_all -> concat) then aggregate with local ids
        x_local = torch.from_numpy(x[lo:hi])
        send = x_local[plan.send_idx.long()]
        recv = torch.empty((plan.n_halo, F))
        dist.all_to_all_single(recv, send.contiguous(), output_split_sizes=plan.recv_counts,
                               input_split_sizes=plan.send_counts)
        np.testing.assert_array_equal(recv.numpy(), x[hg])            # the right rows arrived, in order
        x_ext = torch.cat([x_local, recv])
        eil = plan.edge_index_local()
        for op in ("sum", "mean", "max"):
            got = ref.aggregate(op, x_ext[eil[0].long()], eil[1], plan.n_local).numpy()
            full = ref.aggregate(op, torch.from_numpy(x[ei[0]]), torch.from_numpy(ei[1]), n).numpy()
            np.testing.assert_allclose(got, full[lo:hi], rtol=1e-5, atol=1e-6)
        # the split used to overlap the exchange: local-source part + halo-source part == the whole aggregation,
        # with the halo part indexing the received buffer directly (no [local | halo] concatenation)
        ei_l, ei_h = plan.split_edges()
        assert ei_l.shape[1] + ei_h.shape[1] == eil.shape[1]
        assert ei_l.shape[1] == 0 or int(ei_l[0].max()) < plan.n_local
        assert ei_h.shape[1] == 0 or (int(ei_h[0].min()) >= 0 and int(ei_h[0].max()) < plan.n_halo)
        part_l = ref.aggregate("sum", x_local[ei_l[0].long()], ei_l[1], plan.n_local).numpy()
        part_h = ref.aggregate("sum", recv[ei_h[0].long()], ei_h[1], plan.n_local).numpy()
        full = ref.aggregate("sum", torch.from_numpy(x[ei[0]]), torch.from_numpy(ei[1]), n).numpy()
        np.testing.assert_allclose(part_l + part_h, full[lo:hi], rtol=1e-5, atol=1e-5)
        # reverse direction: halo gradients go back to their owners and are summed per local row
        g_ext = torch.from_numpy(rng.standard_normal((plan.n_local + plan.n_halo, F)).astype(np.float32))
        back = torch.empty((plan.n_send, F))
        dist.all_to_all_single(back, g_ext[plan.n_local:].contiguous(), output_split_sizes=plan.send_counts,
                               input_split_sizes=plan.recv_counts)
        g_local = g_ext[:plan.n_local].clone()
        g_local.index_add_(0, plan.send_idx.long(), back)
        # cross-check with a global computation: gather every rank's g_ext rows by global id
        ids = torch.cat([torch.arange(lo, hi), plan.halo_global])
        all_ids = [None] * world
        all_g = [None] * world
        dist.all_gather_object(all_ids, ids.numpy())
        dist.all_gather_object(all_g, g_ext.numpy())
        tot = np.zeros((n, F), np.float32)
        for i_, g_ in zip(all_ids, all_g):
            np.add.at(tot, i_, g_)
        np.testing.assert_allclose(g_local.numpy(), tot[lo:hi], rtol=1e-5, atol=1e-5)
        # peer-memory transport: the push tables place every rank's block at the right rows of the receiver's window
        # (host emulation: each rank publishes what kgb_halo_push would store where, the receiver assembles it)
        t = plan.push_tables()
        assert t["fwd_begin"][-1] == plan.n_send and t["bwd_begin"][-1] == plan.n_halo
        assert t["n_halo_all"][rank] == plan.n_halo and t["n_send_all"][rank] == plan.n_send
        sends = [(p, t["fwd_row0"][p], x_local[plan.send_idx[t["fwd_begin"][p]:t["fwd_begin"][p + 1]].long()].numpy())
                 for p in range(world)]
        g_halo = g_ext[plan.n_local:]
        backs = [(o, t["bwd_row0"][o], g_halo[t["bwd_begin"][o]:t["bwd_begin"][o + 1]].numpy()) for o in range(world)]
        all_sends, all_backs = [None] * world, [None] * world
        dist.all_gather_object(all_sends, sends)
        dist.all_gather_object(all_backs, backs)
        win_halo = np.full((plan.n_halo, F), np.nan, np.float32)
        win_back = np.full((plan.n_send, F), np.nan, np.float32)
        for blocks in all_sends:
            dest, row0, rows = blocks[rank]
            win_halo[row0:row0 + len(rows)] = rows
        for blocks in all_backs:
            dest, row0, rows = blocks[rank]
            win_back[row0:row0 + len(rows)] = rows
        np.testing.assert_array_equal(win_halo, recv.numpy())    # same rows, same order as the all_to_all
        np.testing.assert_array_equal(win_back, back.numpy())
        # user-supplied ranges: validated, identical on all ranks, and the plan follows them
        from keras_geometric_b200.dist import check_bounds, cost_balanced_bounds
        cb = cost_balanced_bounds(torch.from_numpy(ei[1]), n, world, node_weight=2.0)
        assert check_bounds(cb, n, world) == cb
        mine2 = (ei[1] >= cb[rank]) & (ei[1] < cb[rank + 1])
        plan2 = HaloPlan(torch.from_numpy(ei[0][mine2]), torch.from_numpy(ei[1][mine2]), n, rank, world, bounds=cb)
        assert plan2.n_local == cb[rank + 1] - cb[rank]
        q.put((rank, "ok"))
    except Exception as ex:  # noqa: BLE001
        import traceback
        q.put((rank, "FAIL: " + traceback.format_exc()))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,n,e", [(2, 40, 300), (2, 7, 30), (3, 50, 500)])
def test_halo_plan_gloo(world, n, e):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() + world * 7 + n) % 2000
    procs = [ctx.Process(target=_worker, args=(r, world, port, n, e, 5, 1234 + n, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=180) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
    for rank, msg in res:
        assert msg == "ok", f"rank {rank}: {msg}"


def test_bounds_validation():
    from keras_geometric_b200.dist import check_bounds, cost_balanced_bounds
    assert check_bounds([0, 3, 10], 10, 2) == [0, 3, 10]
    for bad in ([0, 11, 10], [1, 3, 10], [0, 3, 9], [0, 10]):
        with pytest.raises(ValueError, match="bounds must be"):
            check_bounds(bad, 10, 2)
    dst = torch.tensor([0] * 50 + list(range(1, 11)))          # node 0 is a hub
    b = cost_balanced_bounds(dst, 11, 2, node_weight=1.0)
    assert b[0] == 0 and b[-1] == 11 and b[1] <= 2              # the hub's range is short
    assert cost_balanced_bounds(dst, 11, 1) == [0, 11]


def test_partition_bounds():
    from keras_geometric_b200.dist import partition_bounds
    assert partition_bounds(10, 3) == [0, 4, 7, 10]
    assert partition_bounds(2, 4) == [0, 1, 2, 2, 2]
    assert partition_bounds(0, 2) == [0, 0, 0]


def test_scramble_ids_is_a_bijection_and_balances_rmat_ranges():
    """dist.scramble_ids (hash partitioning): a seeded bijection applied to both endpoint rows; after it the
    cost-balanced contiguous ranges of an RMAT graph hold about the same number of nodes AND edges."""
    import sys
    sys.path.insert(0, ROOT) if ROOT not in sys.path else None
    from bench import rmat_edge_index
    from keras_geometric_b200.dist import cost_balanced_bounds, scramble_ids
    n, e, world = 1 << 14, 400_000, 8
    ei = rmat_edge_index(n, e, 14, 3, torch.device("cpu"))
    ei2, perm = scramble_ids(ei, n, seed=7)
    assert sorted(perm.tolist()) == list(range(n))
    assert torch.equal(ei2, perm[ei.long()]) and ei2.dtype == ei.dtype
    ei3, perm3 = scramble_ids(ei, n, seed=7)
    assert torch.equal(perm, perm3) and torch.equal(ei2, ei3)          # reproducible

    def spread(edges):
        b = cost_balanced_bounds(edges[1], n, world, node_weight=28.0)
        nodes = [b[r + 1] - b[r] for r in range(world)]
        cnt = [int(((edges[1] >= b[r]) & (edges[1] < b[r + 1])).sum()) for r in range(world)]
        return max(nodes) / max(min(nodes), 1), max(cnt) / max(min(cnt), 1)

    raw_nodes, raw_edges = spread(ei)
    scr_nodes, scr_edges = spread(ei2)
    assert scr_nodes < 1.3 and scr_edges < 1.3, (scr_nodes, scr_edges)
    assert raw_edges > 2.0 or raw_nodes > 2.0, (raw_nodes, raw_edges)   # what the scrambling removes
