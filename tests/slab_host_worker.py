"""CPU / gloo check of the slab decomposition's HOST logic (no GPU): every rank runs a plain numpy
restatement of the fused PDHG passes on its column slab of a ROF problem, exchanging exactly the
halo columns the CUDA passes exchange (pb_comm.cuh) with torch.distributed (gloo) point-to-point
messages and all-reducing the four residual sums; rank 0 compares the gathered result with the
OpenMP oracle run on the whole image.  What this pins: the partition, shard_description, which
column travels in which direction and when, the lagged residual definition across slab edges and
the global eps / size bookkeeping.  Launched by tests/test_slab_host.py through torchrun."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

F = np.float32


def adj_slab(p1, p2, left_halo, has_left, has_right):
    """K^T y on a slab (block_gradient2d.cu:80-139); p1, p2: (nx, ny); left_halo: column x0-1 of p1."""
    nx, ny = p1.shape
    divx = p1.copy()
    if not has_right:
        divx[nx - 1, :] = 0
    divx[1:, :] -= p1[:-1, :]
    if has_left:
        divx[0, :] -= left_halo
    divy = p2.copy()
    divy[:, ny - 1] = 0
    divy[:, 1:] -= p2[:, :-1]
    return -(divx + divy)


def fwd_slab(u, right_halo, has_right):
    """K u on a slab (block_gradient2d.cu:25-78); right_halo: column x1 of u."""
    gx = np.zeros_like(u)
    gy = np.zeros_like(u)
    gx[:-1, :] = u[1:, :] - u[:-1, :]
    if has_right:
        gx[-1, :] = right_halo - u[-1, :]
    gy[:, :-1] = u[:, 1:] - u[:, :-1]
    return gx, gy


def main():
    import torch
    import torch.distributed as dist
    from prost_b200 import distributed as pbd
    from prost_b200 import synthetic as syn

    dist.init_process_group("gloo")
    rank, world = dist.get_rank(), dist.get_world_size()
    nx, ny, lam, iters, residual_iter = 45, 28, 10.0, 40, 4
    desc = syn.rof(nx, ny, lam)
    part = pbd.SlabPartition(nx, world)
    local = pbd.shard_description(desc, part, rank)
    x0, x1 = part.range(rank)
    w = x1 - x0
    assert local["ncols"] == w * ny and local["nrows"] == 2 * w * ny
    assert local["blocks"][0][3][:3] == [w, ny, 1]
    f = np.asarray(local["prox_g"][0][4][3][1], F).reshape(w, ny)
    np.testing.assert_array_equal(f, desc["data"]["f"].reshape(nx, ny)[x0:x1])
    has_left, has_right = rank > 0, rank + 1 < world

    def exchange(send_left=None, send_right=None):
        """send my column to a neighbour, receive the matching one: returns (from_left, from_right)"""
        reqs, from_left, from_right = [], None, None
        if send_left is not None and has_left:
            reqs.append(dist.isend(torch.from_numpy(np.ascontiguousarray(send_left)), rank - 1))
        if send_right is not None and has_right:
            reqs.append(dist.isend(torch.from_numpy(np.ascontiguousarray(send_right)), rank + 1))
        if send_left is not None and has_right:          # the right neighbour sent me ITS left edge
            buf = torch.empty(ny, dtype=torch.float32)
            dist.recv(buf, rank + 1)
            from_right = buf.numpy()
        if send_right is not None and has_left:
            buf = torch.empty(ny, dtype=torch.float32)
            dist.recv(buf, rank - 1)
            from_left = buf.numpy()
        for r in reqs:
            r.wait()
        return from_left, from_right

    T, S, tau, sigma, theta = F(0.25), F(0.5), F(1), F(1), F(1)
    x = np.zeros((w, ny), F); x_prev = x.copy()
    y1 = np.zeros((w, ny), F); y2 = y1.copy(); y1_prev = y1.copy(); y2_prev = y2.copy()
    kty = np.zeros((w, ny), F); kty_prev = kty.copy()
    kx = (np.zeros((w, ny), F), np.zeros((w, ny), F)); kx_prev = kx
    sums = np.zeros(4)
    n_glob, m_glob = nx * ny, 2 * nx * ny
    for it in range(iters):
        # primal pass; the new column 0 goes to the LEFT neighbour
        arg = x - tau * T * kty
        x_prev, x = x, ((arg - f) / F(1 + lam * tau * T) + f).astype(F)
        _, x_halo = exchange(send_left=x[0])
        kx_prev, kx = kx, fwd_slab(x, x_halo, has_right)
        # dual pass; the new last column of the x-component goes to the RIGHT neighbour
        a1 = y1 + sigma * S * ((1 + theta) * kx[0] - theta * kx_prev[0])
        a2 = y2 + sigma * S * ((1 + theta) * kx[1] - theta * kx_prev[1])
        nrm = np.maximum(np.sqrt(a1 * a1 + a2 * a2), F(1))
        y1_prev, y2_prev, y1, y2 = y1, y2, (a1 / nrm).astype(F), (a2 / nrm).astype(F)
        if it == 0 or it % residual_iter == 0:
            sq, st = np.sqrt(S), np.sqrt(T)
            part_sums = np.zeros(4)
            for yp, yn, k1, k0 in ((y1_prev, y1, kx[0], kx_prev[0]), (y2_prev, y2, kx[1], kx_prev[1])):
                z_hat = (yp - yn) / (sigma * sq) + sq * ((1 + theta) * k1 - theta * k0)
                diff = z_hat - sq * k1
                part_sums[0] += float((diff.astype(np.float64) ** 2).sum())
                part_sums[1] += float((z_hat.astype(np.float64) ** 2).sum())
            w_hat = (x_prev - x) / (tau * st) - st * kty_prev
            diff = w_hat + st * kty
            part_sums[2] = float((diff.astype(np.float64) ** 2).sum())
            part_sums[3] = float((w_hat.astype(np.float64) ** 2).sum())
            t = torch.from_numpy(part_sums)
            dist.all_reduce(t)                      # the one collective of the path
            sums = t.numpy().copy()
        # adjoint for the NEXT primal pass, applied after the residuals (backend_pdhg.cu:372-380)
        y_halo, _ = exchange(send_right=y1[-1])
        kty_prev, kty = kty, adj_slab(y1, y2, y_halo, has_left, has_right)

    gathered = [None] * world
    dist.gather_object(dict(x=x.reshape(-1), y=np.concatenate([y1.reshape(-1), y2.reshape(-1)])),
                       gathered if rank == 0 else None, dst=0)
    if rank == 0:
        from pdhg_util import run_oracle
        want = run_oracle(desc, iters, stepsize="alg1", residual_iter=residual_iter)
        xg = pbd.gather_planar([g["x"] for g in gathered], part, ny)
        yg = pbd.gather_planar([g["y"] for g in gathered], part, ny)
        res = np.sqrt(sums.astype(np.float32))
        out = dict(world=world,
                   err_x=float(np.abs(xg - want["x"]).max() / np.abs(want["x"]).max()),
                   err_y=float(np.abs(yg - want["y"]).max() / np.abs(want["y"]).max()),
                   res=[float(v) for v in res],
                   res_oracle=[want["res"][k] for k in ("primal_residual", "primal_var_norm", "dual_residual",
                                                        "dual_var_norm")],
                   dims=[m_glob, n_glob])
        print("SLAB_HOST " + json.dumps(out), flush=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
