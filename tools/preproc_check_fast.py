"""Validates the second-generation transform kernels (`fast`) bit for bit against the CPU oracle and times both
generations.  Run on the GPU box:  python tools/preproc_check_fast.py > gpurun_out/preproc_fast.txt"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'oracle'))

import numpy as np  # noqa: E402
import torch  # noqa: E402

import __graft_entry__  # noqa: E402


def img(h, w, seed):
    g = np.random.Generator(np.random.PCG64(seed))
    return g.integers(0, 256, size=(h, w, 3), dtype=np.uint8)


def main():
    __graft_entry__.build()
    import preprocess_oracle as po
    from generativeimage2text_b200 import _lib, inference as inf
    shapes = [(480, 640), (640, 480), (224, 224), (225, 1000), (37, 41), (1000, 225), (300, 224), (612, 408), (97, 301),
              (420, 420), (1, 9), (1080, 1920), (223, 225), (333, 500), (501, 333)]
    bad = 0
    for param in ({}, {'test_crop_size': 160}, {'test_crop_size': 480, 'test_respect_ratio_max': 640},
                  {'test_crop_size': 420, 'test_respect_ratio_max': 560}):
        t = inf.get_image_transform(param, fast=True)
        imgs = [img(h, w, i) for i, (h, w) in enumerate(shapes)]
        for rep in range(2):                      # second pass: cached tables, reused buffers
            outs = t.batch(imgs)
            torch.cuda.synchronize()
            for i, im in enumerate(imgs):
                want = po.transform(im, param)
                got = (outs[i] if t.minmax is None else outs[i][0]).cpu().numpy()
                if got.shape != want.shape or not np.array_equal(got, want):
                    bad += 1
                    print('MISMATCH', param, im.shape, rep)
    print('fast kernels vs oracle: %s (%d mismatches)' % ('PASS' if bad == 0 else 'FAIL', bad))
    # timing, device-resident source
    lib = _lib.load()
    dev = torch.device('cuda', 0)
    rows = []
    for fast in (0, 1):
        t = inf.get_image_transform({}, fast=bool(fast))
        t._ensure()
        for (h, w, n) in [(480, 640, 64), (1080, 1920, 16)]:
            g = np.random.Generator(np.random.PCG64(1))
            src = torch.from_numpy(g.integers(0, 256, size=(n, h, w, 3), dtype=np.uint8)).to(dev)
            rh, rw, top, left, oh, ow = t.geometry(h, w)
            descs = (_lib.ImageDesc * n)()
            for i in range(n):
                descs[i] = _lib.ImageDesc(i * h * w * 3, h, w, rh, rw, top, left, oh, ow, i * 3 * oh * ow)
            dst = torch.empty((n, 3, oh, ow), dtype=torch.float32, device=dev)
            stream = torch.cuda.Stream(device=dev)
            with torch.cuda.stream(stream):
                def run():
                    rc = lib.gitb200_preproc_run(t._handle, src.data_ptr(), src.numel(), 0, descs, n, t._mean, t._std,
                                                 dst.data_ptr(), dst.numel(), stream.cuda_stream)
                    assert rc == 0, lib.gitb200_preproc_last_error(t._handle)
                for _ in range(3):
                    run()
                durs = []
                for _ in range(10):
                    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    e0.record(stream)
                    run()
                    e1.record(stream)
                    e1.synchronize()
                    durs.append(e0.elapsed_time(e1))
            ms = sorted(durs)[len(durs) // 2]
            alg = n * h * w * 3 + dst.numel() * 4
            rows.append(dict(fast=fast, images=n, src=[h, w], ms_per_call=round(ms, 4), images_per_s=round(n / ms * 1e3),
                             achieved_gbs=round(alg / ms / 1e6, 1)))
            print(json.dumps(rows[-1]), flush=True)
            ref = dst.clone()
            if fast:
                t0 = inf.get_image_transform({}, fast=False)
                t0._ensure()
                rc = lib.gitb200_preproc_run(t0._handle, src.data_ptr(), src.numel(), 0, descs, n, t0._mean, t0._std, dst.data_ptr(),
                                             dst.numel(), None)
                torch.cuda.synchronize()
                print('fast == first generation on the timing input:', bool(torch.equal(ref, dst)))


if __name__ == '__main__':
    main()
