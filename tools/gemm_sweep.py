"""Times gitb200_op_gemm on the hot-path GEMM shapes (CUDA events, L2 flushed between launches)."""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from generativeimage2text_b200 import _lib

lib = _lib.load()
dev = torch.device('cuda', 0)
stream = torch.cuda.Stream()
flush = torch.empty(160 * 1024 * 1024, dtype=torch.uint8, device=dev)


def bench(M, N, K, act=0, out_bf16=1, resid=False, transposed=0, splits=1, bn=0, iters=8, flush_l2=True):
    a = torch.randn(M, K, device=dev).to(torch.bfloat16)
    w = (torch.randn(N, K, device=dev) * 0.03).to(torch.bfloat16)
    bias = torch.randn(N, device=dev)
    r = torch.randn(M, N, device=dev) if resid else None
    out = torch.zeros(M, N, dtype=torch.bfloat16 if out_bf16 else torch.float32, device=dev)
    durs = []
    with torch.cuda.stream(stream):
        for i in range(iters + 2):
            if flush_l2:
                flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            rc = lib.gitb200_op_gemm(a.data_ptr(), w.data_ptr(), bias.data_ptr() if not (transposed and splits > 1) else None,
                                     r.data_ptr() if r is not None else None, out.data_ptr(), M, N, K, act, out_bf16,
                                     transposed, splits, bn, stream.cuda_stream)
            assert rc == 0, _lib.last_error(None)
            e1.record(stream)
            e1.synchronize()
            if i >= 2:
                durs.append(e0.elapsed_time(e1))
    ms = sum(durs) / len(durs)
    tf = 2.0 * M * N * K / ms / 1e9
    gb = (M * K * 2 + N * K * 2 + M * N * (2 if out_bf16 else 4)) / ms / 1e6
    print('M=%6d N=%5d K=%5d act=%d bf16=%d resid=%d T=%d splits=%d bn=%3d : %8.1f us  %7.1f TFLOP/s  %7.1f GB/s' % (
        M, N, K, act, out_bf16, int(resid), transposed, splits, bn, ms * 1e3, tf, gb), flush=True)


if __name__ == '__main__':
    M = 64 * 197
    print('--- encoder shapes: 1-CTA (bn) vs 2-CTA pairs (bn = 1000 + BN)')
    for bn in (256, 1256):
        bench(M, 3072, 768, act=1, bn=bn)
    for bn in (256, 1256):
        bench(M, 3072, 768, act=0, bn=bn)
    for bn in (192, 1192, 256, 1256):
        bench(M, 768, 3072, act=0, out_bf16=0, resid=True, bn=bn)
    for bn in (256, 1256):
        bench(M, 2304, 768, bn=bn)
    for bn in (192, 1192, 1256):
        bench(M, 768, 768, out_bf16=0, resid=True, bn=bn)
    for bn in (256, 1256):
        bench(M, 3072, 768, act=2, bn=bn)
    bench(8192, 8192, 8192, bn=256, iters=3)
    bench(8192, 8192, 8192, bn=1256, iters=3)
    if len(sys.argv) > 1 and sys.argv[1] == 'enc':
        sys.exit(0)
    print('--- decode-step shapes (swap-AB)')
    for fl in (True, False):
        bench(64, 2304, 768, out_bf16=0, transposed=1, flush_l2=fl)
        bench(64, 768, 768, out_bf16=0, transposed=1, splits=2, flush_l2=fl)
        bench(64, 768, 768, out_bf16=0, transposed=1, splits=1, flush_l2=fl)
        bench(64, 3072, 768, act=2, transposed=1, flush_l2=fl)
        bench(64, 768, 3072, out_bf16=0, transposed=1, splits=4, flush_l2=fl)
        bench(64, 768, 3072, out_bf16=0, transposed=1, splits=1, flush_l2=fl)
        bench(64, 30522, 768, out_bf16=0, transposed=1, flush_l2=fl)
