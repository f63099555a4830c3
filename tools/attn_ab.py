"""A/B of an engine option (default: attn_pipe; e.g. kv_head_major) on whole synchronous calls:
    python tools/attn_ab.py [option] > gpurun_out/attn_ab.txt"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
from generativeimage2text_b200.model import get_git_model, AutoRegressiveBeamSearch  # noqa: E402
from generativeimage2text_b200.synthetic import synthetic_state_dict, synthetic_images  # noqa: E402


class Tok:
    cls_token_id, sep_token_id = 101, 102


OPTION = sys.argv[1] if len(sys.argv) > 1 else 'attn_pipe'
m = get_git_model(Tok(), {})
m.load_state_dict(synthetic_state_dict({}, 0, 'init'))
m = m.cuda().eval()
m.decoder = AutoRegressiveBeamSearch(102, max_steps=40, beam_size=1, per_node_beam_size=1, fix_missing_prefix=True)
s = torch.cuda.Stream()
ref = {}
with torch.cuda.stream(s):
    for rows in (256, 64):
        img = synthetic_images(rows).cuda()
        for pipe in (0, 1, 0, 1):
            m.set_engine_option(OPTION, pipe)
            for _ in range(2):
                out = m({'image': img})
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(s)
            n = 4
            for _ in range(n):
                out = m({'image': img})
            e1.record(s)
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / n
            key = rows
            if key not in ref:
                ref[key] = out['predictions'].clone()
            agree = float((out['predictions'] == ref[key]).float().mean())
            print(json.dumps(dict(rows=rows, option=OPTION, value=pipe, ms_per_launch=round(ms, 3), captions_per_s=round(rows / ms * 1e3, 1),
                                  token_agreement_with_first_run=round(agree, 4))), flush=True)
