"""Localises differences between the one-kernel decode step and the kernel chain: one decode step (max_steps 3), step logits
of both paths for a sweep of row counts and of decoder layers run (engine option debug_layers), plus run-to-run equality."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from generativeimage2text_b200.model import get_git_model, AutoRegressiveBeamSearch
from generativeimage2text_b200.synthetic import synthetic_state_dict, synthetic_images


class Tok:
    cls_token_id, sep_token_id = 101, 102


STEPS = int(os.environ.get('STEPS', 4))
m = get_git_model(Tok(), {})
m.load_state_dict(synthetic_state_dict({}, 1, 'perturbed'))
m = m.cuda().eval()
m.decoder = AutoRegressiveBeamSearch(102, max_steps=STEPS, beam_size=1, per_node_beam_size=1, fix_missing_prefix=True)


def logits(img, mega, layers, forced):
    m.set_engine_option('use_mega', mega)
    m.set_engine_option('debug_layers', layers)
    out = m({'image': img}, forced_tokens=forced, return_step_logits=True)
    torch.cuda.synchronize()
    return out['step_logits'].clone()


for rows in (2, 8, 12, 13, 16, 17, 32, 64):
    img = synthetic_images(rows, 0, 900 + rows).cuda()
    forced = torch.full((rows, STEPS), 1037, dtype=torch.long)
    forced[:, 0] = 101
    for layers in (0, 1, 6):
        a = logits(img, 1, layers, forced)
        a2 = logits(img, 1, layers, forced)
        b = logits(img, 0, layers, forced)
        per_step = [(a[i] - b[i]).abs().max().item() for i in range(a.shape[0])]
        worst_row = (a - b).abs().amax(dim=(0, 2)).argmax().item()
        print('rows %2d layers %d: mega vs chain max |dlogit| per step %s (worst row %d) | mega run-to-run identical: %s' % (
            rows, layers, ['%.4f' % e for e in per_step], worst_row, bool(torch.equal(a, a2))), flush=True)
