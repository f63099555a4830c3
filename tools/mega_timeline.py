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
raw = [(buf[2 * i], int(buf[2 * i + 1])) for i in range(n)]
# fine marks: per launch a 160-entry slice of (%clock64, id) pairs opened by id 700000 (see MegaTl in decode_mega.cuh)
fine, coarse, i = [], [], 0
while i < len(raw):
    if raw[i][1] == 700000:
        fine.append([e for e in raw[i:i + 160] if e[1] != 0])
        i += 160
    else:
        if raw[i][1] != 0:
            coarse.append(raw[i])
        i += 1
ev = sorted(e for e in coarse if e[1] >= 500000)
# one step = 44 barriers (the fine-marked layer 2 logs its barriers through the slice instead): take the 10th step
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
if len(fine) > 10:
    sl = dict((k, t) for t, k in fine[10])
    if all(k in sl for k in (700000, 700001, 700002, 700003)):
        ns_per_clk = (sl[700003] - sl[700001]) / float(sl[700002] - sl[700000])
        print('\nlayer 2 of that step, thread 0 of CTA 0 (clock64 marks, %.3f ns per clock):' % ns_per_clk)
        sub = {0: 'phase start', 1: 'A loads issued', 2: 'A loads landed', 3: 'weight tiles landed', 4: 'MMAs + K-half exchange done',
               5: 'epilogue issued', 6: 'all compute warps of the CTA at the barrier', 7: 'fence done', 8: 'red.release issued',
               9: 'barrier released', 10: 'CTA released'}
        marks = sorted((t, k) for t, k in fine[10] if 710000 <= k < 720000)
        prev = marks[0][0] if marks else 0
        for t, k in marks:
            ph, sb = (k - 710000) // 100, (k - 710000) % 100
            print('  %-6s %-46s +%6.2f us' % (names[ph - 1], sub.get(sb, str(sb)), (t - prev) * ns_per_clk * 1e-3))
            prev = t
        lm = sorted((t, k) for t, k in fine[10] if 720000 <= k < 730000)
        if lm:
            print('\nLM head of that step, thread 0 of CTA 0:')
            what = {0: 'A loads issued', 1: 'tiles landed', 2: 'MMAs + exchange done', 3: 'statistics done'}
            prev = lm[0][0]
            for t, k in lm:
                print('  pair at tile %2d  %-24s +%6.2f us' % ((k - 720000) % 100, what[(k - 720000) // 100], (t - prev) * ns_per_clk * 1e-3))
                prev = t
