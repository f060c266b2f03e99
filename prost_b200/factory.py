"""MATLAB-free factory: builds blocks / proxes / problems from descriptions that use the
reference's mex registry names and argument order (matlab/+prost/private/factory.cpp:18-134,
288-445, 575-656, 820-882).

  prox  description = (name, idx, size, diagsteps, data)      -- the 5-cell of factory.cpp:820-837
  block description = (name, row, col, data)                  -- the 4-cell of factory.cpp:869-882

``data`` per name (same order as the cell arrays the MATLAB front end sends):
  elem_operation:1d:<fun>, elem_operation:norm2:<fun> : [count, dim, interleaved, [a,b,c,d,e,alpha,beta]]
  elem_operation:ind_simplex, elem_operation:ind_sum   : [count, dim, interleaved]
  elem_operation:singular_nx2:{sum_1d:<fun>, ind_l1_ball, moreau:ind_l1_ball},
  elem_operation:eigen_2x2|eigen_3x3|eigen_nxn:<fun>   : [count, dim, interleaved, [a,b,c,d,e,alpha,beta]]
  elem_operation:mass4|ind_comass4_ball|mass5|ind_comass5_ball : [count, dim, interleaved(, [cost])]
  ind_epi_quad                                         : [count, dim, interleaved, [a, b, c]]
  ind_sum                                              : [dim, inds, sum(, dim2, inds2, sum2)]
  ind_epi_conjquad_1d                                  : [count, interleaved, [a, b, c, alpha, beta]]
  ind_range                                            : [A (sparse), AA (dense A^T A)]
  ind_halfspace                                        : [count, dim, interleaved, [a, b]]
  ind_soc                                              : [count, dim, interleaved, alpha]
  moreau                                               : [child description]
  permute                                              : [child description, perm]
  transform                                            : [a, b, c, d, e, child description]   (+function/transform.m)
  zero                                                 : []
  gradient2d / gradient3d : [nx, ny, L, label_first]     diags : [nrows, ncols, factors, offsets]
  dense : [A]     sparse : [A]     zero : [nrows, ncols]     dense_kron_id / id_kron_dense / sparse_kron_id / id_kron_sparse : [K, diaglength]
"""
from . import api


def create_prox(ctx, desc):
    name, idx, size, diagsteps, data = desc
    if name.startswith("elem_operation:1d:"):
        count, dim, interleaved, coeffs = data
        return api.ProxElemOperation1D(ctx, name.split(":")[2], idx, count, dim, interleaved, diagsteps, coeffs)
    if name.startswith("elem_operation:norm2:"):
        count, dim, interleaved, coeffs = data
        return api.ProxElemOperationNorm2(ctx, name.split(":")[2], idx, count, dim, interleaved, diagsteps, coeffs)
    for kind in ("singular_nx2", "eigen_2x2", "eigen_3x3", "eigen_nxn"):       # factory.cpp:49-102
        prefix = f"elem_operation:{kind}:"
        if name.startswith(prefix):
            count, dim, interleaved, coeffs = data
            return api.ProxElemOperationSpectral(ctx, kind, name[len(prefix):], idx, count, dim, interleaved, diagsteps,
                                                 coeffs)
    if name in ("elem_operation:mass4", "elem_operation:ind_comass4_ball", "elem_operation:mass5",
                "elem_operation:ind_comass5_ball"):          # +function/sum_mass_norm.m, sum_ind_comass_ball.m
        count, dim, interleaved = data[:3]
        return api.ProxElemOperationSpectral(ctx, name.split(":")[1], "zero", idx, count, dim, interleaved, diagsteps,
                                             data[3] if len(data) > 3 else None)
    if name == "elem_operation:ind_simplex":
        count, dim, interleaved = data[:3]
        return api.ProxElemOperationIndSimplex(ctx, idx, count, dim, interleaved, diagsteps)
    if name == "elem_operation:ind_sum":
        count, dim, interleaved = data[:3]
        return api.ProxElemOperationIndSum(ctx, idx, count, dim, interleaved, diagsteps)
    if name == "ind_sum":                  # +function/sum_ind_sum2.m: { dim, inds, s1 [, dim2, inds2, s2] }
        if len(data) == 3:
            return api.ProxIndSum(ctx, idx, size, data[0], data[1], data[2])
        return api.ProxIndSum(ctx, idx, size, data[0], data[1], data[2], data[3], data[4], data[5])
    if name == "ind_epi_conjquad_1d":      # [count, interleaved, [a, b, c, alpha, beta]]  (source external, see api)
        count, interleaved, (a, b, c, alpha, beta) = data
        return api.ProxIndEpiConjQuad1D(ctx, idx, count, interleaved, diagsteps, a, b, c, alpha, beta)
    if name == "ind_range":                # +function/ind_range.m: { A, AA }
        return api.ProxIndRange(ctx, idx, size, diagsteps, data[0], data[1] if len(data) > 1 else None)
    if name == "ind_halfspace":
        count, dim, interleaved, (a, b) = data
        return api.ProxIndHalfspace(ctx, idx, count, dim, interleaved, diagsteps, a, b)
    if name == "ind_soc":
        count, dim, interleaved = data[:3]
        return api.ProxIndSOC(ctx, idx, count, dim, interleaved, diagsteps, data[3] if len(data) > 3 else 1.0)
    if name == "ind_epi_quad":
        count, dim, interleaved, (a, b, c) = data
        return api.ProxIndEpiQuad(ctx, idx, count, dim, interleaved, diagsteps, a, b, c)
    if name == "moreau":
        return api.ProxMoreau(ctx, create_prox(ctx, data[0]))
    if name == "permute":
        return api.ProxPermute(ctx, create_prox(ctx, data[0]), data[1])
    if name == "transform":
        a, b, c, d, e, child = data
        return api.ProxTransform(ctx, create_prox(ctx, child), a, b, c, d, e)
    if name == "zero":
        return api.ProxZero(ctx, idx, size)
    raise api.ProstError(-4, f"Unknown prox '{name}'")


def create_block(ctx, desc):
    name, row, col, data = desc
    if name == "gradient2d":
        return api.BlockGradient2D(ctx, row, col, *data)
    if name == "gradient3d":
        return api.BlockGradient3D(ctx, row, col, *data)
    if name == "diags":
        nrows, ncols, factors, offsets = data
        return api.BlockDiags(ctx, row, col, nrows, ncols, offsets, factors)
    if name == "dense":
        return api.BlockDense(ctx, row, col, data[0])
    if name == "dense_kron_id":
        return api.BlockDenseKronId(ctx, row, col, data[0], data[1])
    if name == "id_kron_dense":
        return api.BlockIdKronDense(ctx, row, col, data[0], data[1])
    if name == "sparse_kron_id":
        return api.BlockSparseKronId(ctx, row, col, data[0], data[1])
    if name == "id_kron_sparse":
        return api.BlockIdKronSparse(ctx, row, col, data[0], data[1])
    if name == "sparse":
        return api.BlockSparse(ctx, row, col, data[0])
    if name == "zero":
        return api.BlockZero(ctx, row, col, *data)
    raise api.ProstError(-4, f"Unknown block '{name}'")


def create_linop(ctx, block_descs):
    op = api.LinearOperator(ctx)
    for d in block_descs:
        op.AddBlock(create_block(ctx, d))
    op.Initialize()
    return op


def create_problem(ctx, desc):
    """desc: dict(nrows, ncols, blocks=[...], prox_g/prox_f/prox_gstar/prox_fstar=[...],
    scaling=("alpha", a) | ("identity",) | ("custom", left, right))  (factory.cpp:950-1012)."""
    p = api.Problem(ctx)
    for b in desc.get("blocks", []):
        p.AddBlock(create_block(ctx, b))
    for key, add in (("prox_g", p.AddProx_g), ("prox_f", p.AddProx_f),
                     ("prox_gstar", p.AddProx_gstar), ("prox_fstar", p.AddProx_fstar)):
        for d in desc.get(key, []):
            add(create_prox(ctx, d))
    if "nrows" in desc:
        p.SetDimensions(desc["nrows"], desc["ncols"])
    sc = desc.get("scaling", ("alpha", 1.0))
    if sc[0] == "alpha":
        p.SetScalingAlpha(sc[1])
    elif sc[0] == "identity":
        p.SetScalingIdentity()
    elif sc[0] == "custom":
        p.SetScalingCustom(sc[1], sc[2])
    else:
        raise api.ProstError(-1, "Problem scaling variant not recognized.")
    return p
