"""Sparse source-localization data set: the caller-side data format of BASELINE.json's config 1, at any graph size.

    SourceLocalization(G, nTrain, nValid, nTest, sourceNodes, tMax=None, dataType=np.float64, device='cpu')
                                                                   <- alegnn/utils/dataTools.py:472-592

The reference materialises all powers W^0 .. W^(tMax-1) of the dense normalised adjacency (tMax x N x N,
dataTools.py:563-568) and reads one column per sample.  A sample is x = (W / lambda_max)^t delta_source, so here every
distinct source is diffused once by sparse matrix-vector products and the requested times are read off that trajectory:
O(|sources| * tMax * nnz) work, O(N) memory per trajectory.  Drawing order of the random numbers (sources, then times,
from numpy's global generator) is the reference's, so `np.random.seed(s)` gives the same samples as the reference class.
Same methods as the reference's data classes: getSamples, expandDims, astype, to, evaluate (dataTools.py:172-341).
Host-side numpy / scipy; not part of the GPU path.
"""
import numpy as np
import scipy.sparse as sp
import torch

from .graphtools_sparse import largest_real_eigenvalue


def diffusion_signals(W, sources, times):
    """x_i = (W / lambda_max)^{times[i]} e_{sources[i]} for every i: [len(sources), N] float64."""
    W = sp.csr_matrix(W).astype(np.float64)
    N = W.shape[0]
    Wn = (W / largest_real_eigenvalue(W)).tocsr()                   # dataTools.py:549-553
    sources = np.asarray(sources)
    times = np.asarray(times)
    out = np.zeros((len(sources), N))
    for s in np.unique(sources):
        idx = np.nonzero(sources == s)[0]
        want = times[idx]
        x = np.zeros(N)
        x[s] = 1.0
        for t in range(int(want.max()) + 1):
            if t > 0:
                x = Wn @ x                                           # column s of Wn^t
            for i in idx[want == t]:
                out[i] = x
    return out


def _convert(a, dataType):
    if "torch" in repr(dataType):
        return torch.as_tensor(np.asarray(a) if not isinstance(a, torch.Tensor) else a).to(dataType)
    if isinstance(a, torch.Tensor):
        a = a.cpu().numpy()
    return np.asarray(a).astype(dataType)


class SourceLocalization:
    """Same constructor, attributes (`samples`, `nTrain`, `nValid`, `nTest`, `dataType`, `device`) and methods as the
    reference class; `G` needs `.N` and `.W` (scipy sparse or dense), or may be the adjacency matrix itself."""

    def __init__(self, G, nTrain, nValid, nTest, sourceNodes, tMax=None, dataType=np.float64, device="cpu"):
        W = G.W if hasattr(G, "W") else G
        N = W.shape[0]
        self.dataType = dataType
        self.device = device
        self.nTrain, self.nValid, self.nTest = nTrain, nValid, nTest
        if tMax is None:
            tMax = N
        nTotal = nTrain + nValid + nTest
        sampledSources = np.random.choice(sourceNodes, size=nTotal)      # dataTools.py:557
        sampledTimes = np.random.choice(tMax, size=nTotal)               # dataTools.py:559
        signals = diffusion_signals(W, sampledSources, sampledTimes)
        nodesToLabels = {node: it for it, node in enumerate(sourceNodes)}
        labels = np.array([nodesToLabels[s] for s in sampledSources])
        cuts = {"train": (0, nTrain), "valid": (nTrain, nTrain + nValid), "test": (nTrain + nValid, nTotal)}
        self.samples = {k: {"signals": signals[a:b], "targets": labels[a:b]} for k, (a, b) in cuts.items()}
        self.astype(self.dataType)
        self.to(self.device)

    def getSamples(self, samplesType, *args):
        assert samplesType in ("train", "valid", "test")
        assert len(args) <= 1
        x = self.samples[samplesType]["signals"]
        y = self.samples[samplesType]["targets"]
        if len(args) == 1:
            if type(args[0]) == int:
                assert args[0] <= x.shape[0]
                sel = np.random.choice(x.shape[0], size=args[0], replace=False)
            else:
                sel = args[0]
            xs, y = x[sel], y[sel]
            if len(xs.shape) < len(x.shape):                             # a single sample: keep the sample axis
                xs = xs.unsqueeze(0) if isinstance(xs, torch.Tensor) else np.expand_dims(xs, axis=0)
            x = xs
        return x, y

    def expandDims(self):
        for part in self.samples.values():
            s = part["signals"]
            if s is not None and len(s.shape) in (2, 3):
                axis = len(s.shape) - 1                                  # [n, N] -> [n, 1, N];  [n, T, N] -> [n, T, 1, N]
                part["signals"] = s.unsqueeze(axis) if isinstance(s, torch.Tensor) else np.expand_dims(s, axis=axis)

    def astype(self, dataType):
        is_torch = "torch" in repr(dataType)
        tgt = str(self.samples["train"]["targets"].dtype)
        if "int" in tgt:
            bits64 = "64" in tgt
            targetType = (torch.int64 if bits64 else torch.int32) if is_torch else (np.int64 if bits64 else np.int32)
        else:
            targetType = dataType
        for part in self.samples.values():
            part["signals"] = _convert(part["signals"], dataType)
            part["targets"] = _convert(part["targets"], targetType)
        self.dataType = dataType

    def to(self, device):
        if "torch" in repr(self.dataType):
            for part in self.samples.values():
                for k in part:
                    part[k] = part[k].to(device)
            self.device = device

    def evaluate(self, yHat, y, tol=1e-9):
        """Error rate: fraction of samples whose arg-max label differs from y (dataTools.py:321-341)."""
        n = len(y)
        if "torch" in repr(self.dataType):
            wrong = torch.sum(torch.abs(torch.argmax(yHat, dim=1) - y) > tol)
            return wrong.type(self.dataType) / n
        wrong = np.sum(np.abs(np.argmax(np.array(yHat), axis=1) - np.array(y)) > tol)
        return wrong.astype(self.dataType) / n
