"""Seeded synthetic head outputs for bench.py and the full-size parity tests.

Three logit distributions over [B, N, C] (torch tensors, generated on the device they are asked for):

  dense      N(0, 1): 99.8 % of the scores exceed the 0.05 threshold — worst case for selection.
  sparse     N(-4.595, 1.5^2): centred on the class head's prior bias -log(99) (head/builder.py:30), ~14 % > 0.05.
  clustered  what a TRAINED detector emits: a background floor N(-6, 1) and, per image, 10-40 objects (Zipf-like
             class mix, log-uniform sizes 16 px .. 0.6 * image side, aspect ratios 0.5 .. 2).  Every anchor whose IoU
             with an object exceeds 0.1 gets a logit bump for the object's class that grows with the IoU (IoU 0.8 ->
             logit +5), and its box deltas point at the object (encode(object, anchor) + N(0, 0.05^2)): the top
             scores of a class sit on neighbouring anchors of the same object and overlap heavily — the case NMS
             exists for, and the one where sampled candidate lists run dry.

Deltas are N(0, 0.5^2) clipped to +-4 for dense / sparse (random boxes that rarely overlap).
"""
import math

import torch


def _gen(device, seed):
    g = torch.Generator(device=device)
    g.manual_seed(int(seed))
    return g


def random_inputs(B, N, C, device, seed_logits=42, seed_deltas=1234, dist='dense'):
    g = _gen(device, seed_logits)
    logits = torch.randn((B, N, C), generator=g, device=device, dtype=torch.float32)
    if dist == 'sparse':
        logits.mul_(1.5).add_(-4.595)
    elif dist != 'dense':
        raise ValueError(dist)
    g = _gen(device, seed_deltas)
    deltas = (torch.randn((B, N, 4), generator=g, device=device, dtype=torch.float32) * 0.5).clamp_(-4, 4)
    return logits, deltas


def clustered_inputs(B, anchors, C, H, W, device, seed=42, chunk=8):
    """anchors: [N, 4] = [cx, cy, w, h] in pixels (AnchorBoxGenerator.boxes), any device."""
    anchors = anchors.to(device=device, dtype=torch.float32)
    N = anchors.shape[0]
    g = _gen(device, seed)
    logits = torch.randn((B, N, C), generator=g, device=device, dtype=torch.float32).add_(-6.0)
    deltas = (torch.randn((B, N, 4), generator=g, device=device, dtype=torch.float32) * 0.2)
    O = 40
    side = float(min(H, W))
    # objects: [B, O] centre, size, class, presence
    u = torch.rand((B, O, 6), generator=g, device=device, dtype=torch.float32)
    n_obj = 10 + (u[:, 0, 5] * 31).floor()                       # 10 .. 40 objects per image
    present = torch.arange(O, device=device)[None, :] < n_obj[:, None]
    size = 16.0 * torch.exp(u[..., 0] * math.log(0.6 * side / 16.0))
    ratio = torch.exp((u[..., 1] - 0.5) * 2.0 * math.log(2.0))   # h / w in 0.5 .. 2
    ow = size / ratio.sqrt()
    oh = size * ratio.sqrt()
    ox = u[..., 2] * W
    oy = u[..., 3] * H
    cls = (C * u[..., 4] ** 3).floor().clamp_(0, C - 1).long()    # Zipf-like: low class ids are hot
    ax1, ay1 = anchors[:, 0] - anchors[:, 2] / 2, anchors[:, 1] - anchors[:, 3] / 2
    ax2, ay2 = anchors[:, 0] + anchors[:, 2] / 2, anchors[:, 1] + anchors[:, 3] / 2
    a_area = anchors[:, 2] * anchors[:, 3]
    for b0 in range(0, B, chunk):
        b1 = min(B, b0 + chunk)
        x1 = (ox - ow / 2)[b0:b1, :, None]
        y1 = (oy - oh / 2)[b0:b1, :, None]
        x2 = (ox + ow / 2)[b0:b1, :, None]
        y2 = (oy + oh / 2)[b0:b1, :, None]
        iw = (torch.minimum(x2, ax2) - torch.maximum(x1, ax1)).clamp_(min=0)
        ih = (torch.minimum(y2, ay2) - torch.maximum(y1, ay1)).clamp_(min=0)
        inter = iw * ih
        iou = inter / ((ow * oh)[b0:b1, :, None] + a_area - inter)            # [b, O, N]
        iou = iou * present[b0:b1, :, None]
        bump = 11.0 * ((iou - 0.1) / 0.7).clamp_(0, 1)
        lg = logits[b0:b1]
        ar = torch.arange(b1 - b0, device=device)
        for o in range(O):
            c = cls[b0:b1, o]
            cur = lg[ar, :, c]
            lg[ar, :, c] = torch.maximum(cur, cur * 0.5 - 3.0 + bump[:, o])    # -6 + noise/2 + bump
        # deltas of anchors that see an object: encode(best object, anchor) + noise
        best_iou, best = iou.max(dim=1)                                        # [b, N]
        sel = best_iou > 0.2
        gx = torch.gather(ox[b0:b1], 1, best)
        gy = torch.gather(oy[b0:b1], 1, best)
        gw = torch.gather(ow[b0:b1], 1, best)
        gh = torch.gather(oh[b0:b1], 1, best)
        enc = torch.stack([(gx - anchors[:, 0]) / anchors[:, 2], (gy - anchors[:, 1]) / anchors[:, 3],
                           torch.log(gw / anchors[:, 2]), torch.log(gh / anchors[:, 3])], dim=-1)
        d = deltas[b0:b1]
        d[sel] = enc[sel] + d[sel] * 0.25
    deltas.clamp_(-4, 4)
    return logits, deltas


def make_inputs(dist, B, anchors, C, H, W, device, seed_logits=42, seed_deltas=1234):
    """logits [B, N, C], deltas [B, N, 4] of the named distribution."""
    if dist == 'clustered':
        return clustered_inputs(B, anchors, C, H, W, device, seed=seed_logits)
    return random_inputs(B, anchors.shape[0], C, device, seed_logits, seed_deltas, dist)
