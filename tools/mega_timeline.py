"""Phase timeline of the one-kernel decode step (needs a -DGITB200_TIMELINE build: GITB200_TIMELINE=1 python -m
generativeimage2text_b200.build): CTA 0's arrival at and release from each of the step's grid barriers."""
import os, sys, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from generativeimage2text_b200 import _lib
from generativeimage2text_b200.model import get_git_model, AutoRegressiveBeamSearch
from generativeimage2text_b200.synthetic import synthetic_state_dict, synthetic_images


class Tok:
    cls_token_id, sep_token_id = 101, 102


m = get_git_model(Tok(), {})
m.load_state_dict(synthetic_state_dict({}, 0, 'init'))
m = m.cuda().eval()
m.decoder = AutoRegressiveBeamSearch(102, max_steps=40, beam_size=1, per_node_beam_size=1, fix_missing_prefix=True)
img = synthetic_images(int(os.environ.get('ROWS', '64'))).cuda()
s = torch.cuda.Stream()
lib = _lib.load()
with torch.cuda.stream(s):
    for _ in range(3):
        m({'image': img})
    torch.cuda.synchronize()
    assert lib.gitb200_debug_timeline(1, None, 0) == 0, 'not a timeline build'
    m({'image': img})
    torch.cuda.synchronize()
    buf = (ctypes.c_ulonglong * (2 * 8192))()
    n = lib.gitb200_debug_timeline(0, buf, 8192)
ev = sorted((buf[2 * i], int(buf[2 * i + 1])) for i in range(n))
ev = [(t, k) for t, k in ev if k >= 500000]
# one step = 44 barriers: take the 10th step
starts = [i for i, (t, k) in enumerate(ev) if k == 500001]
a, b = starts[10], starts[11]
names = ['qkv', 'attn', 'oproj', 'ln1', 'fc1', 'fc2', 'ln2']
t0 = ev[a][0]
prev = t0
for t, k in ev[a:b]:
    n_bar = k % 100000
    what = 'arrive ' if k < 600000 else 'release'
    lbl = ('L%d %s' % ((n_bar - 1) // 7, names[(n_bar - 1) % 7])) if n_bar <= 42 else ('lm_head' if n_bar == 43 else 'bar%d' % n_bar)
    print('%9.2f us (+%6.2f)  %s barrier %2d after %s' % ((t - t0) * 1e-3, (t - prev) * 1e-3, what, n_bar, lbl))
    prev = t
print('step total: %.2f us' % ((ev[b][0] - t0) * 1e-3))
