"""Where do the ~17 us of a decode-step (64-row) GEMM go?  Back-to-back launches, CUDA-graph replays, and the
kernel's own %globaltimer stamps (GITB200_GEMM_DBG)."""
import os, sys, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from generativeimage2text_b200 import _lib
lib = _lib.load()
dev = torch.device('cuda', 0)


def mk(rows, feats, K):
    x = torch.randn(rows, K, device=dev).to(torch.bfloat16)
    w = (torch.randn(feats, K, device=dev) * 0.03).to(torch.bfloat16)
    b = torch.randn(feats, device=dev)
    o = torch.zeros(rows, feats, device=dev)
    return x, w, b, o


def gemm(t, st, splits=1):
    x, w, b, o = t
    rc = lib.gitb200_op_gemm(x.data_ptr(), w.data_ptr(), b.data_ptr() if splits == 1 else None, None, o.data_ptr(),
                             x.shape[0], w.shape[0], x.shape[1], 0, 0, 1, splits, 0, st)
    assert rc == 0, _lib.last_error(None)


def ln(xf, g, st):
    rc = lib.gitb200_op_layernorm(xf.data_ptr(), None, None, g.data_ptr(), g.data_ptr(), ctypes.c_float(1e-5), xf.data_ptr(), None,
                                  xf.shape[0], 768, st)
    assert rc == 0


def timed(fn, n, label):
    s = torch.cuda.current_stream()
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(s)
    for _ in range(n):
        fn()
    e1.record(s)
    e1.synchronize()
    print('%-64s %8.2f us per iteration' % (label, e0.elapsed_time(e1) * 1e3 / n), flush=True)


stream = torch.cuda.Stream()
with torch.cuda.stream(stream):
    st = stream.cuda_stream
    qkv, o_, fc1, fc2, lm = mk(64, 2304, 768), mk(64, 768, 768), mk(64, 3072, 768), mk(64, 768, 3072), mk(64, 30522, 768)
    xf = torch.randn(64, 768, device=dev)
    g1 = torch.ones(768, device=dev)
    for t in (qkv, o_, fc1, fc2, lm):
        gemm(t, st)
    ln(xf, g1, st)
    torch.cuda.synchronize()
    timed(lambda: gemm(qkv, st), 50, 'eager back-to-back QKV-shape skinny GEMM')
    timed(lambda: gemm(o_, st), 50, 'eager back-to-back O-shape skinny GEMM')
    timed(lambda: ln(xf, g1, st), 50, 'eager back-to-back LayerNorm 64x768')
    timed(lambda: (gemm(qkv, st), ln(xf, g1, st)), 50, 'eager alternating GEMM + LN (pair)')
    for name, fn in [('graph of 40 x QKV GEMM', lambda: gemm(qkv, st)),
                     ('graph of 40 x O GEMM', lambda: gemm(o_, st)),
                     ('graph of 40 x fc2 GEMM split 4', lambda: gemm(fc2, st, 4)),
                     ('graph of 40 x LM-head GEMM', lambda: gemm(lm, st)),
                     ('graph of 40 x LN', lambda: ln(xf, g1, st)),
                     ('graph of 40 x (GEMM + LN)', lambda: (gemm(qkv, st), ln(xf, g1, st)))]:
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=stream):
            for _ in range(40):
                fn()
        timed(g.replay, 5, name + ' -> per replay / 40 = divide by 40')
    os.environ['GITB200_GEMM_DBG'] = '1'
    for t in (qkv, qkv, o_, fc1, lm):
        gemm(t, st)
    big = mk(12608, 3072, 768)
    x, w, b, o = big
    ob = torch.zeros(12608, 3072, dtype=torch.bfloat16, device=dev)
    for _ in range(2):
        rc = lib.gitb200_op_gemm(x.data_ptr(), w.data_ptr(), b.data_ptr(), None, ob.data_ptr(), 12608, 3072, 768, 1, 1, 0, 1, 0, st)
        assert rc == 0
