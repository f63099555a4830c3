"""Which buffer / which rows differ between the one-kernel decode step and the kernel chain (debug_layers = 1, one step)."""
import os, sys, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from generativeimage2text_b200 import _lib
from generativeimage2text_b200.model import get_git_model, AutoRegressiveBeamSearch
from generativeimage2text_b200.synthetic import synthetic_state_dict, synthetic_images


class Tok:
    cls_token_id, sep_token_id = 101, 102


ROWS = int(os.environ.get('ROWS', 32))
m = get_git_model(Tok(), {})
m.load_state_dict(synthetic_state_dict({}, 1, 'perturbed'))
m = m.cuda().eval()
m.decoder = AutoRegressiveBeamSearch(102, max_steps=3, beam_size=1, per_node_beam_size=1, fix_missing_prefix=True)
lib = _lib.load()
img = synthetic_images(ROWS, 0, 900 + ROWS).cuda()
forced = torch.full((ROWS, 3), 1037, dtype=torch.long)
forced[:, 0] = 101


def read(name, dtype, cols):
    t = torch.empty((ROWS, cols), dtype=dtype)
    n = lib.gitb200_debug_read(m._engine, name.encode(), t.data_ptr(), t.numel() * t.element_size())
    assert n == t.numel() * t.element_size(), (name, n)
    return t.float()


def run(mega):
    m.set_engine_option('use_mega', mega)
    m.set_engine_option('debug_layers', 1)
    m({'image': img}, forced_tokens=forced)
    torch.cuda.synchronize()
    return {k: read(k, dt, c) for k, dt, c in (('ctx', torch.bfloat16, 768), ('ub', torch.bfloat16, 3072), ('hb', torch.bfloat16, 768),
                                                ('x', torch.float32, 768))}


ref = run(0)
for trial in range(8):
    cur = run(1)
    msg = []
    for k in ('ctx', 'ub', 'x'):
        d = (cur[k] - ref[k]).abs()
        bad_rows = (d.amax(dim=1) > 0.05 * ref[k].abs().max()).nonzero().flatten().tolist()
        msg.append('%s max %.4f bad rows %s' % (k, d.max().item(), bad_rows[:12]))
        if k == 'ctx' and bad_rows:
            r = bad_rows[0]
            heads = (d[r].view(12, 64).amax(dim=1) > 0.05 * ref[k].abs().max()).nonzero().flatten().tolist()
            msg.append('row %d bad heads %s' % (r, heads))
    print('trial %d: %s' % (trial, ' | '.join(msg)), flush=True)
