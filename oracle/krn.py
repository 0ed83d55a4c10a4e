"""Oracle (test infrastructure): functional CPU restatement of the KRN graph.

Follows /root/reference/src/nets/park2019.py:32-165 (ConvDw :32-58, RouterV2
:60-80, KeypointRegressionNet :100-165) and the torchvision MobileNetV2 body it
instantiates (torchvision/models/mobilenetv2.py:19-64 InvertedResidual,
:105-114 the (t, c, n, s) table; ``features[:-1]`` = stem + 17 blocks).
Everything is a composition of torch CPU ops on a flat ``{key: tensor}``
state dict using the reference's own state_dict key names.
"""
from collections import OrderedDict

import torch
import torch.nn.functional as F

# torchvision mobilenetv2.py:105-114
MBV2_SETTING = [(1, 16, 1, 1), (6, 24, 2, 2), (6, 32, 3, 2), (6, 64, 4, 2),
                (6, 96, 3, 1), (6, 160, 3, 2), (6, 320, 1, 1)]
BN_EPS = 1e-5
BN_MOMENTUM = 0.1


def mbv2_blocks():
    """[(index, cin, cout, stride, expand_ratio)] for base[1..17]."""
    out, cin, idx = [], 32, 1
    for t, c, n, s in MBV2_SETTING:
        for i in range(n):
            out.append((idx, cin, c, s if i == 0 else 1, t))
            cin, idx = c, idx + 1
    return out


def _bn_keys(d, p, c):
    d[p + '.weight'] = (c,)
    d[p + '.bias'] = (c,)
    d[p + '.running_mean'] = (c,)
    d[p + '.running_var'] = (c,)
    d[p + '.num_batches_tracked'] = ()


def krn_shapes(num_keypoints=11, prefix=''):
    """state_dict key -> shape, in the reference's state_dict order."""
    d = OrderedDict()
    d[prefix + 'base.0.0.weight'] = (32, 3, 3, 3)
    _bn_keys(d, prefix + 'base.0.1', 32)
    for idx, cin, cout, s, t in mbv2_blocks():
        p = prefix + 'base.%d.conv' % idx
        hid = cin * t
        j = 0
        if t != 1:
            d['%s.0.0.weight' % p] = (hid, cin, 1, 1)
            _bn_keys(d, '%s.0.1' % p, hid)
            j = 1
        d['%s.%d.0.weight' % (p, j)] = (hid, 1, 3, 3)
        _bn_keys(d, '%s.%d.1' % (p, j), hid)
        d['%s.%d.weight' % (p, j + 1)] = (cout, hid, 1, 1)
        _bn_keys(d, '%s.%d' % (p, j + 2), cout)
    for i, (inp, oup) in ((0, (320, 1024)), (1, (1024, 1024)), (2, (96, 64)), (3, (1280, 1024))):
        p = prefix + 'extras.%d.conv' % i
        if i == 2:  # RouterV2(96, 64), park2019.py:63-67
            d[p + '.0.weight'] = (64, 96, 1, 1)
            _bn_keys(d, p + '.1', 64)
        else:       # ConvDw, park2019.py:45-54
            d[p + '.0.weight'] = (inp, 1, 3, 3)
            _bn_keys(d, p + '.1', inp)
            d[p + '.3.weight'] = (oup, inp, 1, 1)
            _bn_keys(d, p + '.4', oup)
    d[prefix + 'head.0.weight'] = (2 * num_keypoints, 1024, 7, 7)
    d[prefix + 'head.0.bias'] = (2 * num_keypoints,)
    return d


def is_param(key):
    return not (key.endswith('running_mean') or key.endswith('running_var')
                or key.endswith('num_batches_tracked'))


def _bn(sd, p, x, train, taps=None):
    if train:
        sd[p + '.num_batches_tracked'] += 1
    y = F.batch_norm(x, sd[p + '.running_mean'], sd[p + '.running_var'],
                     sd[p + '.weight'], sd[p + '.bias'], train, BN_MOMENTUM, BN_EPS)
    return y


def _tap(taps, name, t):
    if taps is not None:
        taps[name] = t


def krn_body(sd, x, train, prefix='', taps=None):
    """x [B,3,H,W] -> (feature [B,320,H/32,W/32] = base[17] output, head input chain result
    [B,1024,7,7]).  park2019.py:129-136."""
    p = prefix
    y = F.conv2d(x, sd[p + 'base.0.0.weight'], None, 2, 1)
    _tap(taps, 'base.0.0', y)
    x = F.relu6(_bn(sd, p + 'base.0.1', y, train))
    temp = None
    for idx, cin, cout, s, t in mbv2_blocks():
        q = p + 'base.%d.conv' % idx
        h = x
        j = 0
        if t != 1:
            y = F.conv2d(h, sd['%s.0.0.weight' % q])
            _tap(taps, 'base.%d.conv.0.0' % idx, y)
            h = F.relu6(_bn(sd, '%s.0.1' % q, y, train))
            j = 1
        y = F.conv2d(h, sd['%s.%d.0.weight' % (q, j)], None, s, 1, 1, h.shape[1])
        _tap(taps, 'base.%d.conv.%d.0' % (idx, j), y)
        h = F.relu6(_bn(sd, '%s.%d.1' % (q, j), y, train))
        y = F.conv2d(h, sd['%s.%d.weight' % (q, j + 1)])
        _tap(taps, 'base.%d.conv.%d' % (idx, j + 1), y)
        h = _bn(sd, '%s.%d' % (q, j + 2), y, train)
        x = x + h if (s == 1 and cin == cout) else h   # mobilenetv2.py:32,61-62
        _tap(taps, 'base.%d' % idx, x)
        if idx == 13:
            temp = x                                   # park2019.py:132
    feature = x

    def conv_dw(i, x):                                 # park2019.py:45-54
        q = p + 'extras.%d.conv' % i
        y = F.conv2d(x, sd[q + '.0.weight'], None, 1, 1, 1, x.shape[1])
        _tap(taps, 'extras.%d.conv.0' % i, y)
        x = F.relu(_bn(sd, q + '.1', y, train))
        y = F.conv2d(x, sd[q + '.3.weight'])
        _tap(taps, 'extras.%d.conv.3' % i, y)
        return F.relu(_bn(sd, q + '.4', y, train))

    x = conv_dw(0, x)
    x = conv_dw(1, x)
    # RouterV2, park2019.py:70-80
    q = p + 'extras.2.conv'
    y = F.conv2d(temp, sd[q + '.0.weight'])
    _tap(taps, 'extras.2.conv.0', y)
    x2 = F.leaky_relu(_bn(sd, q + '.1', y, train), 0.2)
    B, C, H, W = x2.shape
    s = 2
    x2 = x2.view(B, C, H // s, s, W // s, s).transpose(3, 4).contiguous()
    x2 = x2.view(B, C, H // s * W // s, s * s).transpose(2, 3).contiguous()
    x2 = x2.view(B, C, s * s, H // s, W // s).transpose(1, 2).contiguous()
    x2 = x2.view(B, s * s * C, H // s, W // s)
    x = torch.cat((x2, x), dim=1)
    _tap(taps, 'extras.2', x)
    x = conv_dw(3, x)
    return feature, x


def krn_logits(sd, x, train, prefix='', taps=None):
    feature, h = krn_body(sd, x, train, prefix, taps)
    out = F.conv2d(h, sd[prefix + 'head.0.weight'], sd[prefix + 'head.0.bias'])
    return feature, out.view(x.shape[0], -1)


def krn_loss(logits, y):
    """park2019.py:142-156: sum over keypoints of batch-mean MSE, x and y separately."""
    nk = logits.shape[1] // 2
    xc, yc = logits[:, 0::2], logits[:, 1::2]
    loss_x = sum(F.mse_loss(xc[:, i], y[:, 0, i]) for i in range(nk))
    loss_y = sum(F.mse_loss(yc[:, i], y[:, 1, i]) for i in range(nk))
    return loss_x + loss_y, loss_x, loss_y


def krn_forward(sd, x, y=None, train=False, prefix='', taps=None):
    """Mirror of KeypointRegressionNet.forward (park2019.py:126-165).
    Train (y given): (loss, {'loss_x','loss_y'}); else (xc, yc)."""
    feature, logits = krn_logits(sd, x, train, prefix, taps)
    _tap(taps, 'feature', feature)
    _tap(taps, 'logits', logits)
    if y is not None:
        loss, lx, ly = krn_loss(logits, y)
        return loss, {'loss_x': float(lx.detach()), 'loss_y': float(ly.detach())}
    return logits[:, 0::2], logits[:, 1::2]
