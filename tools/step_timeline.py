"""In-situ timeline of the decode-step kernels inside the replayed CUDA graph (config 2: GIT_BASE, batch 64)."""
import os, sys, ctypes, collections
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from generativeimage2text_b200 import _lib
from generativeimage2text_b200.model import get_git_model, AutoRegressiveBeamSearch
from generativeimage2text_b200.synthetic import synthetic_state_dict, synthetic_images

class Tok: cls_token_id, sep_token_id = 101, 102
B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
m = get_git_model(Tok(), {})
m.load_state_dict(synthetic_state_dict({}, 0, 'init'))
m = m.cuda().eval()
m.decoder = AutoRegressiveBeamSearch(102, max_steps=40, beam_size=1, per_node_beam_size=1, fix_missing_prefix=True)
img = synthetic_images(B).cuda()
s = torch.cuda.Stream()
lib = _lib.load()
NAMES = {2: 'layernorm', 3: 'decode_attn', 4: 'embed_ln', 5: 'greedy_select'}
with torch.cuda.stream(s):
    for _ in range(3):
        m({'image': img})
    torch.cuda.synchronize()
    lib.gitb200_debug_timeline(1, None, 0)
    m({'image': img})
    torch.cuda.synchronize()
    buf = (ctypes.c_ulonglong * (2 * 8192))()
    n = lib.gitb200_debug_timeline(0, buf, 8192)
ev = [(buf[2 * i], int(buf[2 * i + 1])) for i in range(n)]
print('entries', n)
# durations: time from this kernel's post-wait start to the next kernel's post-wait start
starts = [i for i, (t, k) in enumerate(ev) if k == 4]
print('decode steps seen', len(starts))
agg = collections.OrderedDict()
for si in range(5, min(len(starts) - 1, 35)):
    a, b = starts[si], starts[si + 1]
    for i in range(a, b):
        name = NAMES.get(ev[i][1], 'gemm grid=%d' % (ev[i][1] - 1000))
        d = (ev[i + 1][0] - ev[i][0]) * 1e-3
        x = agg.setdefault(name, [0, 0.0]); x[0] += 1; x[1] += d
nst = min(len(starts) - 1, 35) - 5
tot = sum(v[1] for v in agg.values()) / nst
print('average over %d steps: %.1f us per step' % (nst, tot))
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print('  %-22s n/step=%5.1f  avg %6.2f us  per-step %7.1f us (%4.1f%%)' % (k, v[0] / nst, v[1] / v[0], v[1] / nst, 100 * v[1] / nst / tot))
a, b = starts[10], starts[11]
print('one step in order:')
for i in range(a, b):
    name = NAMES.get(ev[i][1], 'gemm grid=%d' % (ev[i][1] - 1000))
    print('   %-22s %6.2f us' % (name, (ev[i + 1][0] - ev[i][0]) * 1e-3))
