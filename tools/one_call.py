"""One warm `model(batch)` of GIT_BASE with ROWS images, then one more inside a profiler range -- the target of
    ncu --profile-from-start off --metrics ... -k regex:<kernel> python tools/one_call.py <ROWS>"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
from generativeimage2text_b200.model import get_git_model, AutoRegressiveBeamSearch  # noqa: E402
from generativeimage2text_b200.synthetic import synthetic_state_dict, synthetic_images  # noqa: E402


class Tok:
    cls_token_id, sep_token_id = 101, 102


rows = int(sys.argv[1]) if len(sys.argv) > 1 else 64
m = get_git_model(Tok(), {})
m.load_state_dict(synthetic_state_dict({}, 0, 'init'))
m = m.cuda().eval()
m.decoder = AutoRegressiveBeamSearch(102, max_steps=40, beam_size=1, per_node_beam_size=1, fix_missing_prefix=True)
img = synthetic_images(rows).cuda()
s = torch.cuda.Stream()
with torch.cuda.stream(s):
    for _ in range(2):
        m({'image': img})
    torch.cuda.synchronize()
    torch.cuda.profiler.start()
    out = m({'image': img})
    torch.cuda.synchronize()
    torch.cuda.profiler.stop()
print(tuple(out['predictions'].shape))
