"""Per-kernel (start, dependency satisfied, end) stamps of block 0 inside the replayed decode-step graph.

    [ROWS=256] python tools/step_timeline2.py      (ROWS = images per engine launch, default 64)"""
import os, sys, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from generativeimage2text_b200 import _lib
from generativeimage2text_b200.model import get_git_model, AutoRegressiveBeamSearch
from generativeimage2text_b200.synthetic import synthetic_state_dict, synthetic_images

class Tok: cls_token_id, sep_token_id = 101, 102
m = get_git_model(Tok(), {})
m.load_state_dict(synthetic_state_dict({}, 0, 'init'))
m = m.cuda().eval()
m.decoder = AutoRegressiveBeamSearch(102, max_steps=40, beam_size=1, per_node_beam_size=1, fix_missing_prefix=True)
img = synthetic_images(int(os.environ.get('ROWS', '64'))).cuda()
s = torch.cuda.Stream()
lib = _lib.load()
NAMES = {2: 'layernorm', 3: 'decode_attn', 4: 'embed_ln', 5: 'greedy_select'}
def nm(k):
    k = k % 100000
    return NAMES.get(k, 'gemm g=%d' % (k - 1000))
with torch.cuda.stream(s):
    for _ in range(3):
        m({'image': img})
    torch.cuda.synchronize()
    lib.gitb200_debug_timeline(1, None, 0)
    m({'image': img})
    torch.cuda.synchronize()
    buf = (ctypes.c_ulonglong * (2 * 8192))()
    n = lib.gitb200_debug_timeline(0, buf, 8192)
ev = sorted((buf[2 * i], int(buf[2 * i + 1])) for i in range(n))
# find the 12th embed 'dependency satisfied' stamp and print until the next one
idx = [i for i, (t, k) in enumerate(ev) if k == 4]
a, b = idx[12], idx[13]
t0 = ev[a][0]
for t, k in ev[a:b + 1]:
    ph = {0: 'ready ', 1: 'START ', 2: 'end   ', 3: ' data ', 4: ' epi  '}[k // 100000]
    print('%9.2f us  %s %s' % ((t - t0) * 1e-3, ph, nm(k)))
