"""How the CPU baseline (oracle, as-shipped mode) scales with torch threads on this host."""
import sys, os, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'oracle'))
import torch, git_oracle
from generativeimage2text_b200.synthetic import synthetic_state_dict, synthetic_images
sd = synthetic_state_dict({}, 0, 'init')
img = synthetic_images(2, 0, 1234)
print('cpu_count', os.cpu_count())
for t in (8, 16, 32, 64, os.cpu_count()):
    torch.set_num_threads(t)
    t0 = time.perf_counter()
    git_oracle.generate(sd, {}, {'image': img}, 'greedy', 10, cached=False)
    print('threads %3d: %.2f s for B=2, 9 steps (as shipped)' % (t, time.perf_counter() - t0), flush=True)
