"""Generate tests/golden/postproc.npz by running the UNMODIFIED evaluation loops of the reference
(/root/reference/src/core/inference.py: valid_spn :148-221, _keypts_to_pose :223-249) on seeded logits and capturing
what they hand to the CPU-side pose code (pnp / weighted_mean_quaternion are replaced by recorders; EPnP and the
SPEED metrics are out of scope).  Build container only:  python -m oracle.make_golden_postproc
"""
import os
import types

import numpy as np

from oracle.make_golden import OUT, _shims

B_SPN, N_CLS, K_NB = 6, 5000, 5
B_KRN, N_KPT = 5, 11


def synth_inputs():
    import torch
    g = torch.Generator().manual_seed(909)
    weights = torch.randn(B_SPN, N_CLS, generator=g) * 3.0
    weights[1, 17] = weights[1, 4000] = weights[1].max() + 1.0       # an exact tie at the top: lowest index first
    weights[2, :] = weights[2, :].round()                             # many ties
    x_pr, y_pr = torch.rand(B_KRN, N_KPT, generator=g) * 1.2 - 0.1, torch.rand(B_KRN, N_KPT, generator=g) * 1.2 - 0.1
    bbox = torch.stack([torch.tensor([a, a + w, c, c + h]) for a, w, c, h in
                        [(12., 300., 40., 311.), (0., 1920., 0., 1200.), (733., 97., 512., 101.), (1500.5, 400.25, 3., 900.),
                         (250., 640., 100., 480.)]]).float()
    return weights, x_pr, y_pr, bbox


def main():
    _shims()
    import torch
    from config import cfg
    from src.core import inference
    weights, x_pr, y_pr, bbox = synth_inputs()

    # ---- SPN: run valid_spn with a stub model; record what reaches weighted_mean_quaternion ----
    rec = {'qs': [], 'w': None}
    qClass = np.arange(N_CLS * 4, dtype=np.float64).reshape(N_CLS, 4)      # row i = (4i, 4i+1, ..): reveals the index

    def fake_wmq(qs_pr, w):
        rec['qs'].append(np.asarray(qs_pr)[:, 0] / 4)
        rec['w'] = w.numpy().copy()
        return np.array([1.0, 0, 0, 0])
    inference.weighted_mean_quaternion = fake_wmq
    inference.compute_position_spn = lambda *a, **k: np.zeros(3)
    inference.error_orientation = lambda *a, **k: 0.0
    inference.error_translation = lambda *a, **k: 0.0
    inference.speed_score = lambda *a, **k: (0.0, 0.0)
    inference.report_progress = lambda **k: None

    class Stub(torch.nn.Module):
        def forward(self, images):
            return None, weights
    cfg.num_neighbors = K_NB
    loader = [(torch.zeros(B_SPN, 3, 8, 8), torch.zeros(B_SPN, 4), torch.zeros(B_SPN, 4), torch.zeros(B_SPN, 3))]
    inference.valid_spn(0, cfg, Stub(), loader, None, None, None, None, torch.device('cpu'), qClass)
    top_idx = np.stack(rec['qs']).astype(np.int64)
    top_w = rec['w']

    # ---- KRN: _keypts_to_pose with pnp replaced by a recorder ----
    pix = []

    def fake_pnp(corners3D, corners2D, cameraMatrix, distCoeffs):
        pix.append(np.array(corners2D, dtype=np.float32))
        return np.array([1.0, 0, 0, 0]), np.zeros(3)
    inference.pnp = fake_pnp
    for b in range(B_KRN):
        inference._keypts_to_pose(x_pr[b], y_pr[b], bbox[b], None, None)
    np.savez(os.path.join(OUT, 'postproc.npz'), top_w=top_w, top_idx=top_idx, kpt_pix=np.stack(pix))
    print('written', os.path.join(OUT, 'postproc.npz'), top_w.shape, top_idx.shape, np.stack(pix).shape)


if __name__ == '__main__':
    main()
