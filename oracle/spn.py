"""Oracle (test infrastructure): SPN (AlexNet, two FC branches) restatement.

Follows /root/reference/src/nets/spn.py:37-48 (soft-target cross entropy),
:50-99 (layers) and :125-143 (forward); loss mix from
src/core/trainer.py:152-165.
"""
from collections import OrderedDict

import torch
import torch.nn.functional as F


def spn_shapes(num_classes=5000):
    d = OrderedDict()
    for name, shp in (('conv1', (96, 3, 11, 11)), ('conv2', (256, 48, 5, 5)),
                      ('conv3', (384, 256, 3, 3)), ('conv4', (384, 192, 3, 3)),
                      ('conv5', (256, 192, 3, 3)), ('fc6', (4096, 9216)),
                      ('fc7', (4096, 4096)), ('fc8', (num_classes, 4096)),
                      ('fc9', (4096, 9216)), ('fc10', (4096, 4096)),
                      ('fc11', (num_classes, 4096))):
        d[name + '.weight'] = shp
        d[name + '.bias'] = (shp[0],)
    return d


def _lrn(x):   # spn.py:62,67: LocalResponseNorm(2, alpha=2e-5, beta=0.75, k=1.0)
    return F.local_response_norm(x, 2, 2e-5, 0.75, 1.0)


def spn_features(sd, x, taps=None):
    def tap(n, t):
        if taps is not None:
            taps[n] = t
        return t
    x = tap('conv1', F.relu(F.conv2d(x, sd['conv1.weight'], sd['conv1.bias'], 4, 0)))
    x = tap('norm1', _lrn(F.max_pool2d(x, 3, 2)))
    x = tap('conv2', F.relu(F.conv2d(x, sd['conv2.weight'], sd['conv2.bias'], 1, 2, 1, 2)))
    x = tap('norm2', _lrn(F.max_pool2d(x, 3, 2)))
    x = tap('conv3', F.relu(F.conv2d(x, sd['conv3.weight'], sd['conv3.bias'], 1, 1)))
    x = tap('conv4', F.relu(F.conv2d(x, sd['conv4.weight'], sd['conv4.bias'], 1, 1, 1, 2)))
    x = tap('conv5', F.relu(F.conv2d(x, sd['conv5.weight'], sd['conv5.bias'], 1, 1, 1, 2)))
    x = tap('pool5', F.max_pool2d(x, 3, 2))
    return torch.flatten(x, 1)


def spn_forward(sd, x, train=False, drop_p=0.5, taps=None):
    """spn.py:125-143 -> (c, r) logits [B, num_classes] each."""
    f = spn_features(sd, x, taps)

    def branch(a, b, c):
        h = F.dropout(F.relu(F.linear(f, sd[a + '.weight'], sd[a + '.bias'])), drop_p, train)
        h = F.dropout(F.relu(F.linear(h, sd[b + '.weight'], sd[b + '.bias'])), drop_p, train)
        return F.linear(h, sd[c + '.weight'], sd[c + '.bias'])
    return branch('fc6', 'fc7', 'fc8'), branch('fc9', 'fc10', 'fc11')


def soft_ce(logits, target):
    """spn.py:37-48 with reduction='mean'."""
    return (-torch.sum(target.detach() * F.log_softmax(logits, dim=1), dim=1)).mean()


def spn_loss(sd, x, y_classes, y_weights, train=True, drop_p=0.5):
    """trainer.py:152-158: loss_class + 10 * loss_regress."""
    c, r = spn_forward(sd, x, train, drop_p)
    lc, lr = soft_ce(c, y_classes), soft_ce(r, y_weights)
    return lc + 10.0 * lr, lc, lr
