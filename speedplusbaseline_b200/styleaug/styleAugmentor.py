"""StyleAugmentor -- same constructor / call contract as /root/reference/src/styleaug/styleAugmentor.py:12-68:
`StyleAugmentor(alpha, device)(x)` returns a detached, re-styled image batch in (0,1).  The style embedding is
sampled exactly like the reference (CPU `torch.randn(n,100)` -> same RNG stream under torch.manual_seed,
styleAugmentor.py:47), everything after that is libb200sp kernels (ghiasi.GhiasiEngine).

Checkpoints (Ghiasi weights, PBN embedding mean/covariance, SPEED+ mean embedding) are the reference's own
files src/styleaug/checkpoints/*; they are looked up in $SPEEDPLUS_STYLE_CKPT, the staged reference
<repo>/baseline/_ref/src/styleaug/checkpoints (tools/stage_reference.py), then $SPEEDPLUS_REFERENCE/src/styleaug/checkpoints."""
import os

import numpy as np
import torch

from .. import _lib as L
from .ghiasi import GhiasiEngine

_HERE = os.path.dirname(os.path.abspath(__file__))
_ROOT = os.path.dirname(os.path.dirname(_HERE))


def checkpoint_dir():
    cands = [os.environ.get('SPEEDPLUS_STYLE_CKPT', ''), os.path.join(_ROOT, 'baseline', '_ref', 'src', 'styleaug', 'checkpoints'),
             os.path.join(_ROOT, 'baseline', '_ref', 'styleaug_checkpoints'),
             os.path.join(os.environ.get('SPEEDPLUS_REFERENCE', ''), 'src', 'styleaug', 'checkpoints'),
             os.path.join(_HERE, 'checkpoints')]
    for c in cands:
        if c and os.path.exists(os.path.join(c, 'checkpoint_transformer.pth')):
            return c
    raise FileNotFoundError('style-augmentation checkpoints not found; set SPEEDPLUS_STYLE_CKPT to the reference\'s '
                            'src/styleaug/checkpoints directory')


def style_matrix(cov):
    """A = U sqrt(S) from the SVD of the PBN embedding covariance (styleAugmentor.py:39-42); host float64, once."""
    u, s, _ = np.linalg.svd(np.asarray(cov, dtype=np.float64))
    return torch.tensor(np.matmul(u, np.diag(s ** 0.5))).float()


class StyleAugmentor(torch.nn.Module):
    def __init__(self, alpha, device, state=None, use_graph=True):
        """state: optional dict(ghiasi=state_dict, mean=[1,100], cov=[100,100] | A=[100,100], base=[100]) to bypass the
        checkpoint files (synthetic weights in tests / benchmarks)."""
        super().__init__()
        self.alpha = alpha
        self.device = torch.device(device)
        if state is None:
            d = checkpoint_dir()
            ck = torch.load(os.path.join(d, 'checkpoint_transformer.pth'), map_location='cpu', weights_only=False)
            emb = torch.load(os.path.join(d, 'checkpoint_embeddings.pth'), map_location='cpu', weights_only=False)
            state = dict(ghiasi=ck['state_dict_ghiasi'], mean=emb['pbn_embedding_mean'], cov=emb['pbn_embedding_covariance'],
                         base=torch.from_numpy(np.load(os.path.join(d, 'embedding_mean_speedplus.npy'))).float())
        self.engine = GhiasiEngine(state['ghiasi'], self.device)
        A = state['A'] if 'A' in state else style_matrix(state['cov'].numpy() if torch.is_tensor(state['cov']) else state['cov'])
        self.A = A.float().contiguous().to(self.device)                       # 100 x 100
        self.mean = state['mean'].float().reshape(-1).contiguous().to(self.device)
        self.imagenet_embedding = state['base'].float().reshape(-1).contiguous().to(self.device)   # SPEED+ mean, despite the name
        self.use_graph = use_graph
        self._graph = None            # (shape, graph, static x, static noise, static out)

    def sample_noise(self, n):
        """the reference draws on the CPU generator (styleAugmentor.py:47): same stream under torch.manual_seed"""
        noise = torch.randn(n, 100)
        # a FRESH pinned buffer per call (torch's caching host allocator recycles it only after the copy that uses it has
        # completed on its stream): a single reused staging buffer could be overwritten by the next call's draw while the
        # previous asynchronous H2D copy is still queued
        host = torch.empty(n, 100, pin_memory=True)
        host.copy_(noise)
        return host.to(self.device, non_blocking=True)

    def embed(self, noise):
        B = noise.shape[0]
        e = torch.empty(B, 100, dtype=torch.float32, device=self.device)
        L.call('b200sp_style_embed', noise.data_ptr(), self.A.data_ptr(), self.mean.data_ptr(), self.imagenet_embedding.data_ptr(),
               float(self.alpha), e.data_ptr(), B, 100, L.stream_ptr())
        return e

    def _run(self, x, noise, out=None):
        return self.engine.forward(x, self.embed(noise), out=out)

    @torch.no_grad()
    def forward(self, x, noise=None):
        x = x.to(self.device).contiguous().float()
        if noise is None:
            noise = self.sample_noise(x.size(0))
        else:
            noise = noise.to(self.device).contiguous().float()
        if not self.use_graph:
            return self._run(x, noise).detach()
        # the ~50 launches of a call are captured once per input shape; inputs go through static buffers
        if self._graph is None or self._graph[0] != tuple(x.shape):
            sx, sn = torch.empty_like(x), torch.empty_like(noise)
            so = torch.empty_like(x)
            sx.copy_(x); sn.copy_(noise)
            s = torch.cuda.Stream()
            s.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(s):
                self._run(sx, sn, so)                       # warm-up: allocates every plane / buffer
            torch.cuda.current_stream().wait_stream(s)
            torch.cuda.synchronize()
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                self._run(sx, sn, so)
            self._graph = (tuple(x.shape), g, sx, sn, so)
        _, g, sx, sn, so = self._graph
        sx.copy_(x, non_blocking=True)
        sn.copy_(noise, non_blocking=True)
        g.replay()
        return so.clone().detach()
