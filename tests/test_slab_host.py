"""CPU suite: host-side logic of the multi-GPU slab decomposition (prost_b200/distributed.py) and a
world-size-2 gloo run of the decomposed PDHG iteration against the single-image oracle."""
import json
import os
import subprocess
import sys

import numpy as np
import pytest

from prost_b200 import distributed as pbd
from prost_b200 import synthetic as syn

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("nx,world,align", [(4096, 8, 4), (37, 2, 1), (45, 4, 1), (100, 3, 4), (8, 8, 1)])
def test_partition_covers_the_grid(nx, world, align):
    part = pbd.SlabPartition(nx, world, align)
    assert part.range(0)[0] == 0 and part.range(world - 1)[1] == nx
    widths = [part.width(r) for r in range(world)]
    assert sum(widths) == nx and min(widths) >= 1
    for r in range(world - 1):
        assert part.range(r)[1] == part.range(r + 1)[0]
        assert part.range(r)[1] % align == 0
    assert max(widths[:-1]) - min(widths[:-1]) <= align


def test_partition_rejects_too_many_ranks():
    with pytest.raises(ValueError):
        pbd.SlabPartition(3, 4)


def test_slice_and_gather_are_inverse():
    nx, ny, planes = 11, 6, 3
    a = np.arange(nx * ny * planes, dtype=np.float32)
    part = pbd.SlabPartition(nx, 3)
    parts = [pbd.slice_planar(a, nx, ny, *part.range(r)) for r in range(3)]
    assert [p.size for p in parts] == [part.width(r) * ny * planes for r in range(3)]
    np.testing.assert_array_equal(pbd.gather_planar(parts, part, ny), a)
    # column-major planes: element (x, y) of plane l sits at y + x*ny + l*nx*ny (block_gradient2d.cu:59)
    x0, x1 = part.range(1)
    assert parts[1][0] == a[x0 * ny] and parts[1][part.width(1) * ny] == a[x0 * ny + nx * ny]


@pytest.mark.parametrize("make", [lambda: syn.rof(12, 8), lambda: syn.tvl1(12, 8, nc=3),
                                  lambda: syn.tv3d(12, 8, 5), lambda: syn.lifting(12, 8, 6)])
def test_shard_description_is_consistent(make):
    desc = make()
    nx, ny, L = pbd._grid_of(desc)
    world = 3
    part = pbd.SlabPartition(nx, world)
    shards = [pbd.shard_description(desc, part, r) for r in range(world)]
    assert sum(s["nrows"] for s in shards) == desc["nrows"]
    assert sum(s["ncols"] for s in shards) == desc["ncols"]
    for r, s in enumerate(shards):
        w = part.width(r)
        assert s["blocks"][0][3] == [w, ny, L, False]
        for key in ("prox_g", "prox_fstar"):
            ranges = sorted((p[1], p[1] + p[2]) for p in s[key])
            assert ranges[0][0] == 0 and ranges[-1][1] == (s["ncols"] if key == "prox_g" else s["nrows"])
            for (a0, a1), (b0, b1) in zip(ranges, ranges[1:]):
                assert a1 == b0                            # the proxes still tile the index range
            for p, pg in zip(s[key], desc[key]):
                count, dim = p[4][0], p[4][1]
                assert count * dim == p[2] and count * nx == pg[4][0] * w
    # per-element coefficients: gathering the shards' arrays gives back the global array
    g0 = desc["prox_g"][0]
    if g0[0].startswith("elem_operation:1d"):
        b_parts = [s["prox_g"][0][4][3][1] for s in shards]
        np.testing.assert_array_equal(pbd.gather_planar(b_parts, part, ny), np.asarray(g0[4][3][1]).ravel())


def test_shard_description_rejects_what_does_not_shard():
    desc = syn.rof(12, 8)
    part = pbd.SlabPartition(12, 2)
    bad = dict(desc, blocks=[("gradient2d", 0, 0, [12, 8, 1, True])])
    with pytest.raises(Exception):
        pbd.shard_description(bad, part, 0)
    bad = dict(desc, prox_g=[("permute", 0, 96, False, [desc["prox_g"][0], list(range(96))])])
    with pytest.raises(Exception):
        pbd.shard_description(bad, part, 0)


@pytest.mark.parametrize("world", [2, 3])
def test_gloo_slab_iteration_matches_the_oracle(world):
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
           "--master-addr", "127.0.0.1", "--master-port", str(29620 + world),
           os.path.join(ROOT, "tests", "slab_host_worker.py")]
    p = subprocess.run(cmd, capture_output=True, text=True, timeout=300,
                       env=dict(os.environ, OMP_NUM_THREADS="2"))
    lines = [ln for ln in p.stdout.splitlines() if ln.startswith("SLAB_HOST ")]
    assert p.returncode == 0 and lines, f"rc={p.returncode}\n{p.stdout[-2000:]}\n{p.stderr[-2000:]}"
    rep = json.loads(lines[-1][len("SLAB_HOST "):])
    assert rep["world"] == world
    assert rep["err_x"] <= 1e-5 and rep["err_y"] <= 1e-5, rep
    for got, want in zip(rep["res"], rep["res_oracle"]):
        assert abs(got - want) <= 1e-4 * max(abs(want), 1e-6), rep


def test_slab_lifting_generator_equals_sharded_global_description():
    """synthetic.lifting(..., x0, x1) builds one rank's column slab directly (bench_lifting.py at 2048^2 x 32
    never materialises the global problem); it must be exactly shard_description() of the global description,
    and for x0 = 0, x1 = nx (odd label counts included) the global description itself."""
    import numpy as np
    from prost_b200 import distributed as pbd
    from prost_b200 import synthetic as syn

    def same(a, b):
        assert a["nrows"] == b["nrows"] and a["ncols"] == b["ncols"]
        for ba, bb in zip(a["blocks"], b["blocks"]):
            assert ba[0] == bb[0] and int(ba[1]) == int(bb[1]) and int(ba[2]) == int(bb[2])
            if ba[0] == "gradient2d":
                assert [int(v) for v in ba[3][:3]] == [int(v) for v in bb[3][:3]]
            else:
                assert int(ba[3][0]) == int(bb[3][0]) and int(ba[3][1]) == int(bb[3][1])
        for key in ("prox_g", "prox_fstar"):
            assert len(a[key]) == len(b[key])
            for pa, pb_ in zip(a[key], b[key]):
                assert pa[0] == pb_[0] and [int(v) for v in pa[1:3]] == [int(v) for v in pb_[1:3]] and pa[3] == pb_[3]
                assert [int(v) for v in pa[4][:2]] == [int(v) for v in pb_[4][:2]]
                if pa[0] == "ind_epi_quad":
                    for u, v in zip(pa[4][3], pb_[4][3]):
                        assert np.array_equal(np.asarray(u, np.float32).ravel(), np.asarray(v, np.float32).ravel())

    nx, ny, L = 12, 8, 6
    glob = syn.lifting(nx, ny, L)
    for world in (1, 2, 3):
        part = pbd.SlabPartition(nx, world)
        for r in range(world):
            x0, x1 = part.range(r)
            same(pbd.shard_description(glob, part, r), syn.lifting(nx, ny, L, x0=x0, x1=x1))
    # odd L: the pair index runs over the first NL/2 entries of the label-planar vector
    d = syn.lifting(12, 9, 5)
    b = d["prox_fstar"][1][4][3][1]
    ref = (2.0 * syn.uniform(30, np.arange(12 * 9 * 5 // 2, dtype=np.uint64)) - 1.0).astype(np.float32)
    assert np.array_equal(b, ref)


def test_shard_rows_partitions_the_stacked_operator():
    """Host logic of the row-sharded ADMM (prost_b200.distributed.shard_rows): the ranks' row blocks of the sparse +
    dense LASSO operator stack back to the global matrix, f-side coefficients are sliced with their rows, the
    g side stays whole."""
    import scipy.sparse as sp
    from prost_b200 import distributed as pbd
    from prost_b200 import synthetic as syn
    import cases
    desc = syn.lasso(500, 120, nnz_per_row=4, dense=40)
    M = cases.linop_matrix(desc["blocks"]).toarray()
    for world in (1, 2, 3, 7):
        part = pbd.RowPartition(desc["nrows"], world)
        rows, bvec = [], []
        for r in range(world):
            d = pbd.shard_rows(desc, part, r)
            r0, r1 = part.range(r)
            assert d["nrows"] == r1 - r0 and d["ncols"] == desc["ncols"]
            Mr = np.zeros((d["nrows"], d["ncols"]))
            for (name, row, col, data) in d["blocks"]:
                A = data[0].toarray() if sp.issparse(data[0]) else np.asarray(data[0])
                Mr[row:row + A.shape[0], col:col + A.shape[1]] += A
            rows.append(Mr)
            assert d["prox_g"] == desc["prox_g"]
            covered = 0
            for (name, idx, size, ds, (count, dim, il, coeffs)) in d["prox_f"]:
                assert idx == covered and count == size and dim == 1
                covered += size
                bvec.append(coeffs[1])
            assert covered == d["nrows"]
        assert np.allclose(np.vstack(rows), M, rtol=0, atol=0)
        assert np.array_equal(np.concatenate(bvec), desc["prox_f"][0][4][3][1])
