"""In-process sweep of the switches that govern how several batches in flight share one B200.

For each configuration: captions/s of the BASELINE workload (GIT_BASE, batch 64, greedy, max_len 40, device-resident
inputs) with `depth` batches in flight through model.submit().  Run on the GPU box:

    GITB200_SLOTS=8 python tools/overlap_sweep.py [--quick] > gpurun_out/overlap_sweep.txt
"""
import argparse
import itertools
import json
import time
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ.setdefault('GITB200_SLOTS', '8')

import torch  # noqa: E402

import __graft_entry__  # noqa: E402


class Tok:
    cls_token_id, sep_token_id, pad_token_id, mask_token_id = 101, 102, 0, 103


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--quick', action='store_true')
    ap.add_argument('--batches', type=int, default=16)
    ap.add_argument('--out', default='gpurun_out/overlap_sweep.json')
    args = ap.parse_args()
    __graft_entry__.build()
    from generativeimage2text_b200.model import get_git_model, AutoRegressiveBeamSearch
    from generativeimage2text_b200.synthetic import synthetic_state_dict, synthetic_images
    dev = torch.device('cuda', 0)
    torch.cuda.set_device(0)
    model = get_git_model(Tok(), {})
    model.load_state_dict(synthetic_state_dict({}, 0, 'init'), strict=True)
    model = model.to(dev).eval()
    model.decoder = AutoRegressiveBeamSearch(102, max_steps=40, beam_size=1, per_node_beam_size=1, fix_missing_prefix=True)
    B = 64
    img = synthetic_images(B, 0, 1234).contiguous().to(dev)
    stream = torch.cuda.Stream(device=dev)

    def run(k, depth):
        pend, out = [], None
        for _ in range(k):
            if depth == 1:
                out = model({'image': img})
                continue
            pend.append(model.submit({'image': img}, depth=depth))
            if len(pend) >= depth:
                out = pend.pop(0).result()
        while pend:
            out = pend.pop(0).result()
        return out

    def timed(depth):
        with torch.cuda.stream(stream):
            out = run(max(3, 2 * depth), depth)
            torch.cuda.synchronize(dev)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            out = run(args.batches, depth)
            for sl in model._slots:
                if sl['stream'] is not None:
                    stream.wait_stream(sl['stream'])
            e1.record(stream)
            torch.cuda.synchronize(dev)
        return e0.elapsed_time(e1) / args.batches, out['predictions']

    if args.quick:
        grid = [dict(pdl_late=l, prio_split=p, depth=d, sm_reserve=0, decode_ctas=0)
                for l, p, d in itertools.product((0, 1), (0, 1), (1, 4))]
    else:
        grid = [dict(pdl_late=0, prio_split=0, depth=1, sm_reserve=0, decode_ctas=0)]
        for depth in (2, 3, 4):
            grid.append(dict(pdl_late=0, prio_split=0, depth=depth, sm_reserve=0, decode_ctas=0))
        for ctas, reserve in ((32, 0), (64, 0), (48, 32), (64, 32), (96, 0)):   # thinner decode kernels: do the chains of 4 batches overlap?
            for late in (0, 1):
                grid.append(dict(pdl_late=late, prio_split=0, depth=4, sm_reserve=reserve, decode_ctas=ctas))
        for reserve in (0, 32):                     # most informative first: the run may be cut short
            for depth in (4, 6, 8):
                for late, prio in itertools.product((0, 1), (0, 1)):
                    grid.append(dict(pdl_late=late, prio_split=prio, depth=depth, sm_reserve=reserve, decode_ctas=0))
        for late, prio in itertools.product((0, 1), (0, 1)):
            grid.append(dict(pdl_late=late, prio_split=prio, depth=4, sm_reserve=16, decode_ctas=100))
    ref = None
    rows = []
    for cfg in grid:
        t0 = time.time()
        for k in ('pdl_late', 'prio_split', 'sm_reserve', 'decode_ctas'):
            model.set_engine_option(k, cfg[k])
        t1 = time.time()
        ms, toks = timed(cfg['depth'])
        t2 = time.time()
        if ref is None:
            ref = toks.clone()
        agree = float((toks == ref).all(dim=1).float().mean().item())
        row = dict(cfg, ms_per_batch=round(ms, 3), captions_per_s=round(B / ms * 1e3, 1), same_captions=round(agree, 3),
                   host_s=[round(t1 - t0, 2), round(t2 - t1, 2)])
        rows.append(row)
        print(json.dumps(row), flush=True)
    best = max(rows, key=lambda r: r['captions_per_s'])
    print('BEST', json.dumps(best))
    os.makedirs(os.path.dirname(args.out), exist_ok=True)
    with open(args.out, 'w') as f:
        json.dump(rows, f, indent=1)


if __name__ == '__main__':
    main()
