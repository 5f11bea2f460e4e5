"""CPU oracle for the LSIGF graph-filter path.  TEST INFRASTRUCTURE — NOT PRODUCT CODE.

Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s cpu_baseline / `--impl reference` legs may
import this file.  The product path (`graph-neural-networks_b200/`) never does: it fails loudly when
the CUDA extension is missing.

What is restated here (citations are /root/reference/alegnn/utils/graphML.py):

* `lsigf_dense`      – `LSIGF(h, S, x, b)`            :83-176  (shift loop :158-161, tap contraction with
                        flatten order (e,k,g) :170-171, bias :174-175).
* `lsigf_sparse`     – the same arithmetic with each S_e held as a scipy CSR matrix, so the oracle
                        scales to N where the reference's dense E x N x N GSO cannot be allocated.
* `lsigf_grads_*`    – the gradients PyTorch autograd derives for that function.
* `graph_filter_forward` – `GraphFilter.forward`        :2125-2144 (zero-pad :2131-2135, truncate :2142-2143).
* `lsigf_dense_torch` – dense torch CPU port used as the *timed* CPU baseline (same op sequence as the
                        reference: K-1 broadcast batched GEMMs + one contraction GEMM).

A plain-C twin of the forward / backward restatement lives in oracle/lsigf_oracle.c (bound by oracle/c_oracle.py,
checked by tests/test_c_oracle.py).

Pinning: the reference holds no tests or golden vectors for this path, so the oracle is
pinned against outputs of the unmodified reference itself, generated in the authoring container by
`oracle/make_golden.py` and committed under `tests/golden/` (see tests/test_oracle_golden.py), and —
when /root/reference is present — live against the imported reference (tests/test_oracle_vs_reference.py).

Conventions (reference, graphML.py:104-112): h [F,E,K,G]; S [E,N,N]; x [B,G,N]; b [F,1] or [F,N] or None;
y [B,F,N].  The shift is the ROW-vector product x·S (graphML.py:159): (x S)[.., j] = sum_i x[.., i] S[i, j].
All arithmetic here is float64 unless a dtype is passed.
"""
import numpy as np

try:  # scipy is only needed by the sparse variants
    import scipy.sparse as sp
except Exception:  # pragma: no cover
    sp = None


# --------------------------------------------------------------------------------------------
# forward
# --------------------------------------------------------------------------------------------
def shifted_signals_dense(S, x, K):
    """z[b,e,k,g,n] = (x_g S_e^k)[n]   — graphML.py:152-161 (k = 0 is x itself for every e)."""
    S = np.asarray(S)
    x = np.asarray(x)
    E, N, _ = S.shape
    B, G, _ = x.shape
    z = np.empty((B, E, K, G, N), dtype=np.result_type(S, x))
    cur = np.broadcast_to(x[:, None, :, :], (B, E, G, N)).copy()
    z[:, :, 0] = cur
    for k in range(1, K):
        # row-vector shift: contract the node index of x with the FIRST node index of S_e
        cur = np.einsum("begi,eij->begj", cur, S)
        z[:, :, k] = cur
    return z


def lsigf_dense(h, S, x, b=None):
    """y[b,f,n] = sum_{e,k,g} h[f,e,k,g] (x_g S_e^k)[n] + bias  — graphML.py:83-176."""
    h = np.asarray(h)
    F, E, K, G = h.shape
    S = np.asarray(S)
    x = np.asarray(x)
    assert S.shape[0] == E and S.shape[1] == S.shape[2]
    assert x.shape[1] == G and x.shape[2] == S.shape[1]
    z = shifted_signals_dense(S, x, K)
    y = np.einsum("bekgn,fekg->bfn", z, h)
    if b is not None:
        y = y + np.asarray(b)[None, :, :]  # [F,1] or [F,N] broadcast over batch (and nodes)
    return y


def _csr_list(S_list):
    assert sp is not None, "scipy required for the sparse oracle"
    return [sp.csr_matrix(S_e) for S_e in S_list]


def shifted_signals_sparse(S_list, x, K):
    """Same as shifted_signals_dense with S_e as scipy sparse matrices (list of length E).

    (x S)[., j] = sum_i x[., i] S[i, j]  ==  (S^T x^T)^T, so with X node-major [N, B*G] the hop is
    X <- S_e^T X.
    """
    S_list = _csr_list(S_list)
    x = np.asarray(x)
    B, G, N = x.shape
    E = len(S_list)
    X0 = np.ascontiguousarray(x.reshape(B * G, N).T)  # [N, C]
    z = np.empty((E, K, N, B * G), dtype=X0.dtype)
    for e in range(E):
        St = S_list[e].T.tocsr()
        cur = X0
        z[e, 0] = cur
        for k in range(1, K):
            cur = St @ cur
            z[e, k] = cur
    return z  # [E,K,N,C]


def lsigf_sparse(h, S_list, x, b=None):
    """Sparse restatement of LSIGF (graphML.py:83-176); S_list = [S_e as scipy sparse], same semantics."""
    h = np.asarray(h)
    F, E, K, G = h.shape
    x = np.asarray(x)
    B, _, N = x.shape
    assert len(S_list) == E and x.shape[1] == G
    z = shifted_signals_sparse(S_list, x, K).reshape(E, K, N, B, G)
    y = np.einsum("eknbg,fekg->bfn", z, h)
    if b is not None:
        y = y + np.asarray(b)[None, :, :]
    return y


def graph_filter_forward(weight, bias, S, x):
    """GraphFilter.forward — graphML.py:2125-2144: right-zero-pad the node axis up to N, filter, keep first Nin."""
    S = np.asarray(S)
    x = np.asarray(x)
    N = S.shape[1]
    B, G, Nin = x.shape
    if Nin < N:
        x = np.concatenate([x, np.zeros((B, G, N - Nin), dtype=x.dtype)], axis=2)
    u = lsigf_dense(weight, S, x, bias)
    return u[:, :, :Nin]


# --------------------------------------------------------------------------------------------
# backward (what autograd computes for the function above)
# --------------------------------------------------------------------------------------------
def lsigf_grads_dense(h, S, x, dy, bias_shape=None):
    """Returns (dh, dx, db) for upstream gradient dy [B,F,N].

    dh[f,e,k,g] = sum_{b,n} dy[b,f,n] z[b,e,k,g,n]
    dx          = sum_{e,k} (dy^T-contracted taps) shifted back with S_e^T:  dx_g = sum_{e,k,f} h[f,e,k,g] dy_f (S_e^T)^k
    db          = sum over batch (and over nodes when the bias is [F,1])
    """
    h = np.asarray(h)
    S = np.asarray(S)
    x = np.asarray(x)
    dy = np.asarray(dy)
    F, E, K, G = h.shape
    z = shifted_signals_dense(S, x, K)
    dh = np.einsum("bfn,bekgn->fekg", dy, z)
    # v[b,e,k,f,n] = (dy_f (S_e^T)^k)[n]
    St = np.transpose(S, (0, 2, 1))
    v = shifted_signals_dense(St, dy, K)
    dx = np.einsum("bekfn,fekg->bgn", v, h)
    db = None
    if bias_shape is not None:
        db = dy.sum(axis=0)
        if bias_shape[1] == 1:
            db = db.sum(axis=1, keepdims=True)
    return dh, dx, db


def lsigf_grads_sparse(h, S_list, x, dy, bias_shape=None):
    h = np.asarray(h)
    x = np.asarray(x)
    dy = np.asarray(dy)
    F, E, K, G = h.shape
    B, _, N = x.shape
    S_list = _csr_list(S_list)
    z = shifted_signals_sparse(S_list, x, K).reshape(E, K, N, B, G)
    dh = np.einsum("bfn,eknbg->fekg", dy, z)
    v = shifted_signals_sparse([S_e.T for S_e in S_list], dy, K).reshape(E, K, N, B, F)
    dx = np.einsum("eknbf,fekg->bgn", v, h)
    db = None
    if bias_shape is not None:
        db = dy.sum(axis=0)
        if bias_shape[1] == 1:
            db = db.sum(axis=1, keepdims=True)
    return dh, dx, db


def threaded_spmm(M):
    """scipy CSR matrix -> callable X -> M @ X (float64) running on all host cores through torch.sparse (MKL); scipy's own
    CSR x dense product is single-threaded, which makes the full-size checks (N = 1M, C = 2048) take minutes."""
    import warnings
    import torch
    M = sp.csr_matrix(M)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")                      # "sparse CSR support is in beta"
        Mt = torch.sparse_csr_tensor(torch.from_numpy(M.indptr.astype(np.int64)),
                                     torch.from_numpy(M.indices.astype(np.int64)),
                                     torch.from_numpy(M.data.astype(np.float64)), size=M.shape)
    return lambda X: torch.sparse.mm(Mt, torch.from_numpy(np.ascontiguousarray(X, dtype=np.float64))).numpy()


def lsigf_sparse_stream(h, S_list, x, b=None, spmm=None):
    """lsigf_sparse without materialising z [E,K,N,C]: hop by hop (graphML.py:158-161), each shifted signal contracted
    with its taps as soon as it exists (graphML.py:170-171 distributes over k).  Same result; memory 3 node-major
    matrices instead of E*K, BLAS GEMMs instead of an einsum — this is what the full-size GPU parity tests
    (N = 1M, C = 2048, E = 4) run on the host.  Cross-checked against lsigf_sparse in tests/test_oracle_golden.py."""
    h = np.asarray(h, dtype=np.float64)
    F, E, K, G = h.shape
    x = np.asarray(x, dtype=np.float64)
    B, _, N = x.shape
    S_list = _csr_list(S_list)
    assert len(S_list) == E and x.shape[1] == G
    X0 = np.ascontiguousarray(x.reshape(B * G, N).T)                       # [N, C], column (b, g) at b*G + g
    y = np.zeros((N * B, F))
    for e in range(E):
        St = S_list[e].T.tocsr()
        shift = (lambda X, St=St: St @ X) if spmm is None else spmm(St)    # `spmm`: optional threaded M -> (X -> M @ X)
        cur = X0
        for k in range(K):
            if k > 0:
                cur = shift(cur)                                           # x <- x S_e  (row-vector shift)
            y += cur.reshape(N * B, G) @ h[:, e, k, :].T
    y = y.reshape(N, B, F).transpose(1, 2, 0)
    if b is not None:
        y = y + np.asarray(b, dtype=np.float64)[None, :, :]
    return y


def lsigf_grads_sparse_stream(h, S_list, x, dy, bias_shape=None, spmm=None):
    """Streaming form of lsigf_grads_sparse: dh[f,e,k,g] = <dy_f, z_{e,k,g}>, dx = sum_{e,k} (dy S_e^T^k) h."""
    h = np.asarray(h, dtype=np.float64)
    x = np.asarray(x, dtype=np.float64)
    dy = np.asarray(dy, dtype=np.float64)
    F, E, K, G = h.shape
    B, _, N = x.shape
    S_list = _csr_list(S_list)
    X0 = np.ascontiguousarray(x.reshape(B * G, N).T)                       # [N, B*G]
    D0 = np.ascontiguousarray(dy.reshape(B * F, N).T)                      # [N, B*F]
    dh = np.zeros((F, E, K, G))
    dxn = np.zeros((N * B, G))
    for e in range(E):
        St = S_list[e].T.tocsr()
        Sm = S_list[e].tocsr()
        fwd = (lambda X, St=St: St @ X) if spmm is None else spmm(St)
        bwd = (lambda X, Sm=Sm: Sm @ X) if spmm is None else spmm(Sm)
        z, v = X0, D0
        for k in range(K):
            if k > 0:
                z = fwd(z)                                                 # z_k = z_{k-1} S_e
                v = bwd(v)                                                 # v_k = v_{k-1} S_e^T
            dh[:, e, k, :] = D0.reshape(N * B, F).T @ z.reshape(N * B, G)
            dxn += v.reshape(N * B, F) @ h[:, e, k, :]
    dx = dxn.reshape(N, B, G).transpose(1, 2, 0)
    db = None
    if bias_shape is not None:
        db = dy.sum(axis=0)
        if bias_shape[1] == 1:
            db = db.sum(axis=1, keepdims=True)
    return dh, dx, db


# --------------------------------------------------------------------------------------------
# dense torch CPU port: the timed CPU baseline (bench.py cpu_baseline / --impl reference)
# --------------------------------------------------------------------------------------------
def lsigf_dense_torch(h, S, x, b=None):
    """Dense torch restatement with the reference's own op mix (graphML.py:152-175): K-1 broadcast
    batched GEMMs x <- x @ S over a dense E x N x N GSO, the growing stack of shifted signals, one
    [B,N,EKG] x [EKG,F] contraction GEMM, then the bias add.  Written from the description of that
    algorithm (SURVEY.md §2.2 a-g), not copied."""
    import torch
    F, E, K, G = h.shape
    B, _, N = x.shape
    cur = x.unsqueeze(1).expand(B, E, G, N)
    stack = [cur]
    Sb = S.unsqueeze(0)
    for _ in range(1, K):
        cur = torch.matmul(cur, Sb)  # [B,E,G,N] x [1,E,N,N]
        stack.append(cur)
    z = torch.stack(stack, dim=2)  # [B,E,K,G,N]
    zt = z.permute(0, 4, 1, 2, 3).reshape(B, N, E * K * G)
    y = torch.matmul(zt, h.reshape(F, E * K * G).t()).permute(0, 2, 1)
    if b is not None:
        y = y + b
    return y


# --------------------------------------------------------------------------------------------
# seeded inputs shared by tests / golden generation
# --------------------------------------------------------------------------------------------
def random_sparse_gso(rng, N, avg_deg, E=1, symmetric=False, dtype=np.float64):
    """E random sparse GSOs as dense [E,N,N] arrays (small N only). Non-symmetric unless asked, so
    the row-vector convention (x·S vs S·x) is actually exercised."""
    S = np.zeros((E, N, N), dtype=dtype)
    p = min(1.0, avg_deg / max(N, 1))
    for e in range(E):
        mask = rng.random((N, N)) < p
        w = rng.standard_normal((N, N))
        A = np.where(mask, w, 0.0)
        if symmetric:
            A = np.triu(A, 1)
            A = A + A.T
        # keep the spectral radius near 1 so K hops neither blow up nor vanish
        scale = np.abs(A).sum(axis=1).max()
        S[e] = A / (scale if scale > 0 else 1.0)
    return S


def random_case(seed, N, B, G, F, K, E=1, avg_deg=6, bias="F1", symmetric=False):
    rng = np.random.default_rng(seed)
    S = random_sparse_gso(rng, N, avg_deg, E, symmetric)
    x = rng.standard_normal((B, G, N))
    bound = 1.0 / np.sqrt(G * K)  # graphML.py:2109-2114
    h = rng.uniform(-bound, bound, (F, E, K, G))
    if bias == "F1":
        b = rng.uniform(-bound, bound, (F, 1))
    elif bias == "FN":
        b = rng.uniform(-bound, bound, (F, N))
    else:
        b = None
    dy = rng.standard_normal((B, F, N))
    return dict(h=h, S=S, x=x, b=b, dy=dy)


# --------------------------------------------------------------------------------------------
# edge-variant filter (variant row a-7)
# --------------------------------------------------------------------------------------------
def evgf_dense(Phi, x, b=None):
    """EVGF(S, x, b) — graphML.py:389-488.  Phi [F,E,K,G,N,N], x [B,G,N].  Column convention (graphML.py:464,475):
    u_0 = Phi^(0) x_g, u_k = Phi^(k) u_{k-1};  y_f = sum_{e,k,g} u_k + b_f."""
    Phi = np.asarray(Phi)
    x = np.asarray(x)
    F, E, K, G, N, _ = Phi.shape
    u = np.einsum("fegij,bgj->bfegi", Phi[:, :, 0], x)
    y = u.sum(axis=(2, 3))
    for k in range(1, K):
        u = np.einsum("fegij,bfegj->bfegi", Phi[:, :, k], u)
        y = y + u.sum(axis=(2, 3))
    if b is not None:
        y = y + np.asarray(b)[None]
    return y


def edge_variant_masks(S, M, K, tol=1e-9):
    """sparsityPatternFull of EdgeVariantGF.addGSO (graphML.py:2608-2668): [1,E,K,1,N,N]; k = 0 is the identity on the
    first M nodes, k >= 1 the pattern of |S|+I, both restricted to entries with i < M or j < M when M < N."""
    S = np.asarray(S)
    E, N, _ = S.shape
    eye = np.eye(N)[None].repeat(E, axis=0)
    pattern = ((np.abs(S) + eye) > tol).astype(S.dtype)
    if M < N:
        idx = np.arange(N)
        hybrid = ((idx[:, None] < M) | (idx[None, :] < M)).astype(S.dtype)
    else:
        hybrid = np.ones((N, N), dtype=S.dtype)
    pattern = pattern * hybrid[None]
    ident = eye * hybrid[None]
    full = np.concatenate([ident[:, None]] + [pattern[:, None]] * (K - 1), axis=1)  # [E,K,N,N]
    return full[None, :, :, None]


def edge_variant_gf_forward(weightEV, weightLSI, bias, S, M, x):
    """EdgeVariantGF.forward — graphML.py:2670-2698 (bias enters twice in the hybrid case, as in the reference)."""
    K = weightEV.shape[2]
    N = S.shape[1]
    Phi = np.asarray(weightEV) * edge_variant_masks(S, M, K)
    y = evgf_dense(Phi, x, bias)
    if M < N:
        y = y + lsigf_dense(weightLSI, S, x, bias)
    return y


def evgf_sparse_chains(rowptr, col, w, xA, k0_identity=False):
    """EVGF (graphML.py:389-488) for ONE output feature f and one edge feature e on a compact node set, with the filter
    matrices given per non-zero of a CSR pattern: w [K, G, nnz], xA [B, G, NA] -> sum_g sum_k u_k, [B, NA], where
    u_0 = Phi^(0) x_g, u_k = Phi^(k) u_{k-1} (column convention, :464,:475).  k0_identity: Phi^(0) keeps only its
    diagonal (the layer's k = 0 mask, :2653-2663).  fp64 scipy; what the full-size EdgeNet check (bench cfg4ev) runs."""
    w = np.asarray(w, dtype=np.float64)
    xA = np.asarray(xA, dtype=np.float64)
    K, G, nnz = w.shape
    B, _, NA = xA.shape
    rowptr = np.asarray(rowptr, dtype=np.int64)
    col = np.asarray(col, dtype=np.int64)
    rows = np.repeat(np.arange(NA), np.diff(rowptr))
    out = np.zeros((NA, B))
    for g in range(G):
        u = np.ascontiguousarray(xA[:, g, :].T)                      # [NA, B]
        for k in range(K):
            wk = w[k, g]
            if k == 0 and k0_identity:
                wk = np.where(rows == col, wk, 0.0)
            u = sp.csr_matrix((wk, col, rowptr), shape=(NA, NA)) @ u
            out += u
    return out.T


# --------------------------------------------------------------------------------------------
# sparse torch CPU restatement: second timed CPU baseline (NOT reference code — the reference has no sparse path)
# --------------------------------------------------------------------------------------------
def prepare_sparse_torch(csr_list, N, dtype):
    """torch sparse CSR of S_e^T for every e (the forward shift gathers along columns of S_e); done once, untimed."""
    import torch
    out = []
    for (rowptr, col, val) in csr_list:
        S = torch.sparse_csr_tensor(torch.from_numpy(np.asarray(rowptr, dtype=np.int64)),
                                    torch.from_numpy(np.asarray(col, dtype=np.int64)),
                                    torch.from_numpy(np.asarray(val)).to(dtype), size=(N, N))
        out.append(S.to_sparse_coo().t().to_sparse_csr())
    return out


def lsigf_sparse_torch(h, csr_list, x, b=None, prepared=None):
    """Same arithmetic as lsigf_sparse with torch.sparse CSR x dense products (multi-threaded on the host).
    h [F,E,K,G] torch; csr_list: per e (rowptr, col, val) numpy arrays of S_e; x [B,G,N] torch; returns [B,F,N]."""
    import torch
    F, E, K, G = h.shape
    B, _, N = x.shape
    if prepared is None:
        prepared = prepare_sparse_torch(csr_list, N, x.dtype)
    X0 = x.reshape(B * G, N).t().contiguous()                      # node-major [N, C]
    y = torch.zeros(N, B, F, dtype=x.dtype)
    y += torch.einsum("nbg,fg->nbf", X0.view(N, B, G), h[:, :, 0, :].sum(1))
    for e in range(E):
        St = prepared[e]
        cur = X0
        for k in range(1, K):
            cur = torch.sparse.mm(St, cur)
            y += torch.einsum("nbg,fg->nbf", cur.view(N, B, G), h[:, e, k, :])
    y = y.permute(1, 2, 0)
    if b is not None:
        y = y + b
    return y
