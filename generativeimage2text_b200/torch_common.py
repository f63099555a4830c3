"""Checkpoint ingestion for the engine's parameter shell -- the host-side mirror of the reference's
`generativeimage2text/torch_common.py` (same function names, argument meaning and tolerance of mismatches), so that

    checkpoint = torch_load('output/GIT_BASE/snapshot/model.pt')['model']
    load_state_dict(model, checkpoint)            # reference inference.py:84-86, 147-149

drives `get_git_model(...)` of this package unchanged.  Host logic only: no tensor of the
hot path is touched here; the engine re-packs whatever ends up in the module's parameters on the next call
(`GitB200CaptioningModel._ensure_engine`).
"""
import io
import logging
from pprint import pformat

import torch


def resize_2d_pos_embed(origin_pos_embed, origin_input, patch_size, after_input):
    """CLIP positional embedding [1 + g*g, d] (or [1, 1 + g*g, d]) re-sampled to the grid of another square input
    resolution: CLS row kept, grid rows bicubic (reference torch_common.py:19-39; used by get_image_encoder,
    model.py:75-90, when `test_crop_size` != 224)."""
    squeeze = origin_pos_embed.dim() == 2
    pe = origin_pos_embed.unsqueeze(0) if squeeze else origin_pos_embed
    assert origin_input % patch_size == 0 and after_input % patch_size == 0
    g0, g1 = origin_input // patch_size, after_input // patch_size
    d = pe.shape[-1]
    assert pe.shape[1] == g0 * g0 + 1
    grid = pe[0, 1:, :].reshape(g0, g0, d).permute(2, 0, 1).unsqueeze(0)
    grid = torch.nn.functional.interpolate(grid, size=(g1, g1), mode='bicubic')
    grid = grid.squeeze(0).permute(1, 2, 0).reshape(-1, d)
    out = torch.cat((pe[0, 0:1, :], grid), dim=0).unsqueeze(0)
    return out.squeeze(0) if squeeze else out


def torch_load(filename):
    """reference torch_common.py:41-45 without the azfuse indirection (plain local files)."""
    with open(filename, 'rb') as fp:
        buf = io.BytesIO(fp.read())
    return torch.load(buf, map_location=lambda storage, loc: storage)


def remove_prefix(model, prefix):
    """Strip every leading repetition of `prefix` from the keys (reference torch_common.py:47-53)."""
    out = {}
    for k, v in model.items():
        while k.startswith(prefix):
            k = k[len(prefix):]
        out[k] = v
    return out


def strip_prefix_if_present(state_dict, prefix):
    return remove_prefix(state_dict, prefix)


def align_and_update_state_dicts(model_state_dict, loaded_state_dict):
    """Suffix matching of reference torch_common.py:100-145: every model key takes the value of the LONGEST loaded key
    that is a (plain string) suffix of it; model keys without a match are removed from `model_state_dict`.

    The reference materialises the full |model| x |loaded| match matrix; a suffix of `key` is `key[s:]`, so probing
    the loaded keys with every tail, longest first, finds the same winner in O(len(key)) set lookups."""
    loaded = set(loaded_state_dict.keys())
    updated, used = [], set()
    for key in sorted(model_state_dict.keys()):
        for s in range(len(key)):
            tail = key[s:]
            if tail in loaded:
                model_state_dict[key] = loaded_state_dict[tail]
                updated.append(key)
                used.add(tail)
                logging.info('%s will be loaded from %s of shape %s', key, tail, tuple(loaded_state_dict[tail].shape))
                break
    logging.info('target model param = %d; name matched = %d; loaded = %d', len(model_state_dict), len(updated),
                 len(loaded_state_dict))
    logging.info('from loaded; ignore = %s', pformat([k for k in loaded_state_dict if k not in used]))
    keep = set(updated)
    for k in [k for k in model_state_dict.keys() if k not in keep]:
        del model_state_dict[k]


def load_model_state_ignore_mismatch(model, init_dict):
    """reference torch_common.py:58-91: tensors whose name or shape does not fit are skipped (logged), the rest goes
    through `model.load_state_dict(strict=False)`."""
    name_to_param = dict(model.named_parameters())
    name_to_param.update(dict(model.named_buffers()))
    real, unknown, mismatched = {}, [], []
    for k, v in init_dict.items():
        if k not in name_to_param:
            unknown.append(k)
        elif tuple(v.shape) != tuple(name_to_param[k].shape):
            logging.info('%s shape is not consistent, expected: %s; got %s', k, name_to_param[k].shape, v.shape)
            mismatched.append(k)
        else:
            real[k] = v
    logging.info('unique keys in init dict = %s; total = %d', pformat(unknown), len(unknown))
    result = model.load_state_dict(real, strict=False)
    logging.info('unique key (not initialized) in current model = %s', pformat(result.missing_keys))
    return result


def load_state_dict(model, loaded_state_dict):
    """reference torch_common.py:93-98."""
    model_state_dict = model.state_dict()
    loaded_state_dict = strip_prefix_if_present(loaded_state_dict, prefix='module.')
    align_and_update_state_dicts(model_state_dict, loaded_state_dict)
    return load_model_state_ignore_mismatch(model, model_state_dict)
