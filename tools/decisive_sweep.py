"""Seed sweep for the decisive-margin checkpoints (SURVEY.md section 7 hard part 1b): smallest free-running greedy
top-1/top-2 margin of the oracle per (weight seed, image seed).   python tools/decisive_sweep.py [param-name] [n]"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'oracle'))
import torch, git_oracle
from generativeimage2text_b200.synthetic import synthetic_state_dict, synthetic_images

PARAMS = {'base': {}, 'large': {'visual_feature_size': 1024, 'image_encoder_type': 'CLIPViT_L_14'}}
name = sys.argv[1] if len(sys.argv) > 1 else 'base'
n = int(sys.argv[2]) if len(sys.argv) > 2 else 40
B = int(os.environ.get('B', 4)); STEPS = int(os.environ.get('STEPS', 20)); SEED = int(os.environ.get('SEED', 0))
torch.set_num_threads(os.cpu_count())
sd = synthetic_state_dict(PARAMS[name], SEED, 'decisive')
best = []
for img_seed in range(5000, 5000 + n):
    img = synthetic_images(B, 0, img_seed)
    trace = []
    t0 = time.time()
    out = git_oracle.generate(sd, PARAMS[name], {'image': img}, 'greedy', STEPS, cached=True, trace=trace)
    pred = out['predictions']
    mins = []
    for i, z in enumerate(trace):
        top = z.topk(2, dim=1).values
        m = top[:, 0] - top[:, 1]
        live = torch.isfinite(m)                       # EOS-forced rows: -inf elsewhere -> margin inf
        if i > 0:
            live &= pred[:, i] != 102                  # the row's input token at this step (pred has CLS at col 0)
        if live.any():
            mins.append(m[live].min().item())
    mm = min(mins)
    best.append((mm, img_seed))
    print('img_seed %d: min margin %.4f, len %d, %.1fs, ends %s' % (img_seed, mm, pred.shape[1], time.time() - t0,
          [(r == 102).nonzero()[:1].flatten().tolist() for r in pred]), flush=True)
best.sort(reverse=True)
print('best:', best[:5])
