"""Receiver-block partition + halo exchange on CPU (gloo, world_size 2 and 3): index plan and both exchange
directions.  The fused kernels themselves need a GPU; the partitioned result vs the single-GPU result is checked
on the GPU box by scripts/check_partition.py under torchrun."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from aero_gnn_b200.meshes import airfoil_o_mesh
from aero_gnn_b200.partition import HaloExchanger, block_bounds, build_halo_plan


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


@pytest.mark.parametrize("reorder", [True, False])
def test_halo_plan_covers_every_edge_once(reorder):
    mesh = airfoil_o_mesh(16, 9, seed=0)
    ei, n = mesh.edge_index.numpy(), mesh.num_nodes
    for world in (1, 2, 3, 8, 200):          # 200 > number of nodes in some blocks: empty ranks
        seen = []
        for r in range(world):
            pl = build_halo_plan(ei, n, r, world, reorder=reorder)
            lo, hi = block_bounds(n, world, r)
            assert (pl.lo, pl.hi) == (lo, hi)
            g = ei[:, pl.edge_ids]
            assert np.all((g[1] >= lo) & (g[1] < hi))
            # local ids map back to the global sender / receiver; own rows are a permutation of the block
            assert np.array_equal(np.sort(pl.own_order), np.arange(hi - lo))
            table = np.concatenate([pl.own_order + lo, pl.halo_global])
            assert np.array_equal(table[pl.local_edge_index[0]], g[0])
            assert np.array_equal(table[pl.local_edge_index[1]], g[1])
            # interior receivers (local rows < n_interior) have owned senders only; boundary receivers at least one remote
            recv_l, send_l = pl.local_edge_index[1], pl.local_edge_index[0]
            assert np.all(send_l[recv_l < pl.n_interior] < pl.n_own)
            if reorder:
                has_remote = np.zeros(pl.n_own, dtype=bool)
                has_remote[recv_l[send_l >= pl.n_own]] = True
                assert np.array_equal(has_remote, np.arange(pl.n_own) >= pl.n_interior)
                assert np.all(np.diff(pl.own_order[:pl.n_interior]) > 0) and np.all(np.diff(pl.own_order[pl.n_interior:]) > 0)
            else:
                assert pl.n_interior == 0 and np.array_equal(pl.own_order, np.arange(pl.n_own))
            assert np.all((pl.halo_global < lo) | (pl.halo_global >= hi)) and np.all(np.diff(pl.halo_global) > 0)
            assert sum(pl.recv_counts) == pl.n_halo
            seen.append(pl.edge_ids)
        assert np.array_equal(np.sort(np.concatenate(seen)), np.arange(ei.shape[1]))
        # what r sends to p is exactly what p expects from r
        plans = [build_halo_plan(ei, n, r, world, reorder=reorder) for r in range(world)]
        for r in range(world):
            for p in range(world):
                if r != p:
                    plo, phi = block_bounds(n, world, r)
                    want = plans[p].halo_global[(plans[p].halo_global >= plo) & (plans[p].halo_global < phi)]
                    assert np.array_equal(plans[r].own_order[plans[r].send_idx[p]] + plo, want)


def _worker(rank, world, port, ei, n, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        pl = build_halo_plan(ei, n, rank, world)
        ex = HaloExchanger(pl, "cpu")
        xg = torch.arange(n, dtype=torch.float32)[:, None] * torch.ones(1, 4) + torch.arange(4) * 0.25
        own = torch.from_numpy(pl.own_order)                         # local row i = global own row own[i]
        x_loc = xg[pl.lo:pl.hi][own].contiguous()
        halo = ex.forward(x_loc)
        ok_f = torch.equal(halo, xg[torch.from_numpy(pl.halo_global)])
        # reverse: every rank returns (rank+1) * ones for each of its halo rows
        g_own = torch.zeros(pl.n_own, 4)
        ex.backward(torch.full((pl.n_halo, 4), float(rank + 1)), g_own)
        expect_g = torch.zeros(pl.n_own, 4)                          # in global own order
        for p in range(world):
            if p != rank:
                other = build_halo_plan(ei, n, p, world)
                ids = other.halo_global[(other.halo_global >= pl.lo) & (other.halo_global < pl.hi)] - pl.lo
                expect_g[torch.from_numpy(ids)] += float(p + 1)
        expect = expect_g[own]
        ok_b = torch.equal(g_own, expect)
        # split-phase form: post, do unrelated work, finish; receive straight into a slice of a larger buffer
        x_ext = torch.full((pl.n_local, 4), -1.0)
        x_ext[:pl.n_own] = x_loc
        tok = ex.forward_start(x_ext, out=x_ext[pl.n_own:])
        ex.forward_finish(tok)
        ok_f = ok_f and torch.equal(x_ext[pl.n_own:], xg[torch.from_numpy(pl.halo_global)])
        g2 = torch.zeros(pl.n_own, 4)
        tok = ex.backward_start(torch.full((pl.n_halo, 4), float(rank + 1)), g2)
        g2 += 0.0
        ex.backward_finish(tok, g2)
        ok_b = ok_b and torch.equal(g2, expect)
        # HaloExtendFn (GMP / WeightedEdgeConv variant): extended rows forward, halo gradients back to their owners
        import types
        from aero_gnn_b200.partition import HaloExtendFn
        part = types.SimpleNamespace(exchanger=ex, n_own=pl.n_own, plan=types.SimpleNamespace(N=pl.n_local))
        xl = x_loc.clone().requires_grad_(True)
        xe = HaloExtendFn.apply(part, xl)
        ok_f = ok_f and torch.equal(xe[pl.n_own:], xg[torch.from_numpy(pl.halo_global)]) and torch.equal(xe[:pl.n_own], x_loc)
        wgt = torch.cat([torch.full((pl.n_own, 4), 0.5), torch.full((pl.n_halo, 4), float(rank + 1))])
        (xe * wgt).sum().backward()
        ok_b = ok_b and torch.equal(xl.grad, expect + 0.5)
        out[rank] = (ok_f, ok_b, pl.n_halo)
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_halo_exchange_gloo(world):
    mesh = airfoil_o_mesh(16, 9, seed=1)
    ei, n = mesh.edge_index.numpy(), mesh.num_nodes
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), ei, n, out), nprocs=world, join=True)
    assert len(out) == world
    for r in range(world):
        ok_f, ok_b, nh = out[r]
        assert ok_f and ok_b and nh > 0
