"""Teacher-forced parity of a whole TRAIN step.

A deep ReLU + BatchNorm network at random init amplifies a one-ulp difference in an early
activation into O(1) differences of the late ones, so an end-to-end gradient comparison between a
16-bit-storage pipeline and fp32 autograd measures that amplification, not the kernels.  This
checker removes the amplification: after ``TrainNet.forward_backward`` (built with
``keep_dy=True``) it walks every recorded unit and recomputes it with torch fp32 ops FROM THE
TENSORS THE B200 PATH STORED (its inputs, its incoming gradient), comparing

  forward : conv output z, activation a (BN batch statistics, residual, ReLU), pool outputs, logits, loss
  backward: dz, dgamma, dbeta, dW of every conv / BN, fc gradients, and for every activation tensor
            the SUM of its consumers' data gradients (conv transposes, pool backward, residual routes,
            head) against the gradient buffer the path accumulated.

Every comparison is local, so the tolerance is a few 16-bit roundings.
"""
import torch
import torch.nn.functional as F


def _nchw(view):
    return view.interior().float().permute(0, 3, 1, 2).contiguous()


def _check(name, got, want, tol, stats, robust=False):
    scale = float(want.abs().max()) + 1e-30
    err = (got - want).abs()
    rel = float((got - want).norm() / (want.norm() + 1e-30))
    if robust:      # a pre-activation within fp32 noise of zero may sit on the other side of the ReLU
        bad = float((err > tol * scale).float().mean())
        ok = bad <= 2e-4 and rel <= 4 * tol
    else:
        bad = float(err.max()) / scale
        ok = bad <= tol
    stats.append((name, rel, bad, ok))
    return ok


def local_parity(net, eps_of, tol=2.0 ** -7):
    """Returns (all_ok, [(name, rel_l2, worst, ok), ...]).  ``eps_of(bn_name)`` -> BatchNorm eps."""
    tdt = net.tdtype
    q = lambda t: t.to(tdt).float()
    stats = []
    contrib = {}            # (id(tensor), c0, c1) -> (view, expected gradient sum NCHW fp32)

    def add(view, g):
        key = (id(view.t), view.c0, view.c1)
        if key in contrib:
            contrib[key] = (view, contrib[key][1] + g)
        else:
            contrib[key] = (view, g)

    for rec in net.records:
        kind = rec['kind']
        if kind == 'conv_bn':
            nm = rec['name']
            pw, pg, pb = rec['pw'], rec['pg'], rec['pb']
            Ci, kh, kw, Co = rec['Ci'], rec['kh'], rec['kw'], rec['Co']
            w = pw.w[:, :, :Ci].reshape(Co, kh, kw, Ci).permute(0, 3, 1, 2).contiguous()
            if rec['stem'] is not None:
                # the patch matrix the 1x1 GEMM reads: k = (r*kw + s)*3 + c of the 16-bit rounded input
                g = rec['stem']
                taps = g['kh'] * g['kw']
                u = F.unfold(q(net.inp), (g['kh'], g['kw']), padding=g['pad'], stride=g['stride'])
                u = u.view(net.batch, 3, taps, rec['H'], rec['W']).permute(0, 2, 1, 3, 4).reshape(net.batch, taps * 3, rec['H'], rec['W'])
                got = _nchw(rec['x'])
                _check(nm + ':patches', got[:, :taps * 3], u, 0.0, stats)
                _check(nm + ':patch_pad', got[:, taps * 3:].abs().sum().view(1), torch.zeros(1, device=got.device), 0.0, stats)
            # the conv reference runs in float64: cuDNN's "fp32" convolutions on this GPU carry ~1e-3 relative
            # error of the ABSOLUTE sum (tools/dbg_dgrad.py), which swamps weight gradients that are small
            # differences of large terms (BN backward makes dz zero-mean per channel; the branch_pool gradients
            # came out 50-90 % off in fp32 while agreeing with an explicit float sum to 1e-3: tools/dbg_poolafter.py)
            x_in, w_op = _nchw(rec['x']).double(), q(w).double()
            x_in.requires_grad_(True)
            w_op.requires_grad_(True)
            z_exp = F.conv2d(x_in, w_op, stride=rec['stride'], padding=rec['pad'])
            if rec.get('pool_after') is not None:                 # Inception branch_pool: pool and 1x1 conv commuted
                z_exp = F.avg_pool2d(z_exp, *rec['pool_after'])
            z_ours = _nchw(rec['z'])
            _check(nm + ':z', z_ours, z_exp.detach().float(), tol, stats)
            zz = z_ours.clone().requires_grad_(True)
            gamma, beta = pg.w.clone().requires_grad_(True), pb.w.clone().requires_grad_(True)
            y = F.batch_norm(zz, None, None, gamma, beta, training=True, eps=eps_of(nm))
            res = None
            if rec['residual'] is not None:
                res = _nchw(rec['residual']).requires_grad_(True)
                y = y + res
            if rec['relu']:
                # the ReLU mask is teacher-forced too: a pre-activation within fp32 noise of zero may sit on the other side
                # here, and ONE such element moves a dbeta / dgamma that is a cancelling sum by several per cent.  The
                # masks must agree on all but 2e-4 of the elements; the values are checked against relu(y) regardless
                m_ours = _nchw(rec['out']) > 0
                flips = float(((y.detach() > 0) != m_ours).float().mean())
                stats.append((nm + ':relu_mask', flips, flips, flips <= 2e-4))
                _check(nm + ':a', _nchw(rec['out']), F.relu(y.detach()), tol, stats)
                y = y * m_ours
            else:
                _check(nm + ':a', _nchw(rec['out']), y.detach(), tol, stats)
            da = _nchw(rec['dy'])
            y.backward(da)
            dz_ours = _nchw(rec['dz'])
            _check(nm + ':dz', dz_ours, zz.grad, tol, stats, robust=True)
            _check(nm + ':dgamma', pg.g, gamma.grad, 4 * tol, stats)
            _check(nm + ':dbeta', pb.g, beta.grad, 4 * tol, stats)
            if res is not None:
                add(rec['residual'], res.grad)
            # conv backward from OUR dz
            z_exp.backward(dz_ours.double())
            dW = pw.g[:, :, :Ci].reshape(Co, kh, kw, Ci).permute(0, 3, 1, 2)
            _check(nm + ':dW', dW, w_op.grad.float(), tol, stats)
            if rec['stem'] is None:
                add(rec['x'], x_in.grad.float())
        elif kind in ('maxpool', 'avgpool'):
            x_in = _nchw(rec['x']).requires_grad_(True)
            k, s, p = rec['k'], rec['stride'], rec['pad']
            y = F.max_pool2d(x_in, k, s, p) if kind == 'maxpool' else F.avg_pool2d(x_in, k, s, p)
            _check('%s%d:y' % (kind, len(stats)), _nchw(rec['out']), y.detach(), tol if kind == 'avgpool' else 0.0, stats)
            y.backward(_nchw(net.grad_of(rec['out'])))
            add(rec['x'], x_in.grad)
        elif kind == 'head':
            x_in = _nchw(rec['x']).requires_grad_(True)
            Wt, bias = rec['pw'].w.clone().requires_grad_(True), rec['pb'].w.clone().requires_grad_(True)
            pooled = x_in.mean((2, 3))
            use_drop = rec['drop'] is not None and (net.dropout or rec.get('fixed_mask'))
            if use_drop:
                pooled = pooled * rec['drop'].view(net.batch, -1)
            logits = F.linear(pooled, Wt, bias)
            weight = 0.4 if rec['which'] == 'aux' else 1.0
            loss = weight * F.cross_entropy(logits, net.labels)
            _check(rec['which'] + ':logits', rec['logits'].view(net.batch, -1), logits.detach(), 1e-4, stats)
            loss.backward()
            _check(rec['which'] + ':dfc.weight', rec['pw'].g, Wt.grad, 1e-3, stats)
            _check(rec['which'] + ':dfc.bias', rec['pb'].g, bias.grad, 1e-3, stats)
            add(rec['x'], x_in.grad)
            rec['_loss'] = float(loss)
    for key, (view, g) in contrib.items():
        _check('grad[%dx%dx%d c%d:%d]' % (view.H, view.W, view.t.shape[3], view.c0, view.c1), _nchw(net.grad_of(view)), g, 2 * tol, stats,
               robust=True)
    want_loss = sum(r['_loss'] for r in net.records if r['kind'] == 'head')
    got_loss = float(net.loss[0])
    stats.append(('loss', abs(got_loss - want_loss) / abs(want_loss), abs(got_loss - want_loss), abs(got_loss - want_loss) <= 1e-4 * abs(want_loss)))
    return all(s[3] for s in stats), stats
