"""Shared helpers for the parity tests (test infrastructure)."""
import json
import os

import numpy as np
import torch

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')


def load_golden(name):
    g = np.load(os.path.join(GOLDEN_DIR, name + '.npz'))
    d = {k: g[k] for k in g.files}
    d['meta'] = json.loads(str(d['meta']))
    return d


def golden_inputs(meta):
    from generativeimage2text_b200.synthetic import synthetic_state_dict, synthetic_images
    sd = synthetic_state_dict(meta['param'], meta['seed'], meta['variant'])
    image = synthetic_images(meta['batch'], meta['frames'], meta['img_seed'], meta.get('image_hw', 224))
    batch = {'image': image}
    if 'prefix' in meta:
        batch['prefix'] = torch.tensor([meta['prefix']], dtype=torch.long)
    return sd, batch


def greedy_margins(raw_logits, tokens_in, eos=102):
    """Decision margin (top1 - top2) per row of one greedy step after the reference's masking
    (no-repeat scatter, layers/decoder.py:330; skipped on the very first step :257-273)."""
    z = raw_logits.clone()
    if tokens_in is not None:
        z.scatter_(1, tokens_in[:, None], -10000.0)
    top = z.topk(2, dim=1).values
    return top[:, 0] - top[:, 1]
