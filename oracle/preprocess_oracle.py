"""TEST INFRASTRUCTURE ONLY -- CPU restatement (numpy) of the reference's test-time image transform
(generativeimage2text/inference.py:111-132 `get_image_transform`, :29-64 `MinMaxResizeForTest`), i.e. of the
third-party code it calls:

  * torchvision.transforms `Resize(size, BICUBIC)` on a PIL image (torchvision 0.26: shorter edge -> size, longer edge
    -> int(size * long / short), then `img.resize((w, h), BICUBIC)`), `CenterCrop` (top/left = int(round((in - out)/2))),
    `ToTensor` (uint8 HWC -> fp32 CHW / 255), `Normalize` ((x - mean) / std, fp32, IEEE division);
  * Pillow (12.2 here) `Image.resize` = libImaging/Resample.c `ImagingResample`: separable, horizontal pass first then
    vertical, each with per-output-pixel windows [xmin, xmin+xmax) of a bicubic (a = -0.5) kernel stretched by
    max(scale, 1) ("antialias"), coefficients normalised in double, converted to 22-bit fixed point
    (`normalize_coeffs_8bpc`), accumulated in int32 starting from 1 << 21, and clipped to uint8 after EACH pass.

Pinned (tests/test_preprocess_oracle.py) against PIL / torchvision executing in this container, bit for bit.
Nothing in the product package may import this file; only tests/ and __graft_entry__.smoke() do.
"""
import math

import numpy as np

PRECISION_BITS = 32 - 8 - 2
CLIP_MEAN = (0.48145466, 0.4578275, 0.40821073)
CLIP_STD = (0.26862954, 0.26130258, 0.27577711)


def bicubic_filter(x):
    """Keys' cubic convolution kernel with a = -0.5 (Resample.c `bicubic_filter`), evaluated in double."""
    a = -0.5
    if x < 0.0:
        x = -x
    if x < 1.0:
        return ((a + 2.0) * x - (a + 3.0)) * x * x + 1
    if x < 2.0:
        return (((x - 5) * x + 8) * x - 4) * a
    return 0.0


def precompute_coeffs(in_size, out_size):
    """Resample.c `precompute_coeffs` (box = the whole axis) followed by `normalize_coeffs_8bpc`.
    -> ksize, bounds int32 [out, 2] (first tap, tap count), kk int32 [out, ksize] fixed-point weights."""
    scale = float(np.float32(in_size) - np.float32(0.0)) / out_size      # (double)(in1 - in0) / outSize, in0/in1 are C floats
    filterscale = max(scale, 1.0)
    support = 2.0 * filterscale
    ksize = int(math.ceil(support)) * 2 + 1
    bounds = np.zeros((out_size, 2), dtype=np.int32)
    kk = np.zeros((out_size, ksize), dtype=np.int32)
    ss = 1.0 / filterscale
    for xx in range(out_size):
        center = 0.0 + (xx + 0.5) * scale
        xmin = int(center - support + 0.5)
        if xmin < 0:
            xmin = 0
        xmax = int(center + support + 0.5)
        if xmax > in_size:
            xmax = in_size
        xmax -= xmin
        w = [bicubic_filter((x + xmin - center + 0.5) * ss) for x in range(xmax)]
        ww = 0.0
        for v in w:
            ww += v
        for x in range(xmax):
            v = w[x] / ww if ww != 0.0 else w[x]
            kk[xx, x] = int(-0.5 + v * (1 << PRECISION_BITS)) if v < 0 else int(0.5 + v * (1 << PRECISION_BITS))
        bounds[xx] = (xmin, xmax)
    return ksize, bounds, kk


def _pass(src, bounds, kk, axis):
    """One 8-bit resampling pass along `axis` (0 = vertical, 1 = horizontal) of an HxWxC uint8 image."""
    src = src.astype(np.int64)
    n_out = bounds.shape[0]
    shape = list(src.shape)
    shape[axis] = n_out
    out = np.empty(shape, dtype=np.uint8)
    for o in range(n_out):
        lo, cnt = int(bounds[o, 0]), int(bounds[o, 1])
        k = kk[o, :cnt].astype(np.int64)
        if axis == 1:
            acc = (src[:, lo:lo + cnt, :] * k[None, :, None]).sum(axis=1)
        else:
            acc = (src[lo:lo + cnt, :, :] * k[:, None, None]).sum(axis=0)
        acc = acc + (1 << (PRECISION_BITS - 1))
        # int32 accumulation cannot overflow: |sum| <= 255 * sum|k| < 2^31 (Resample.c keeps two guard bits)
        val = np.clip(acc >> PRECISION_BITS, 0, 255).astype(np.uint8)
        if axis == 1:
            out[:, o, :] = val
        else:
            out[o, :, :] = val
    return out


def pil_resize_bicubic(img, out_h, out_w):
    """`PIL.Image.fromarray(img).resize((out_w, out_h), BICUBIC)` for an HxWx3 uint8 array."""
    h, w = img.shape[:2]
    cur = img
    if out_w != w:                      # ImagingResample: the horizontal pass is skipped when the width is unchanged
        _, b, k = precompute_coeffs(w, out_w)
        cur = _pass(cur, b, k, axis=1)
    if out_h != h:
        _, b, k = precompute_coeffs(h, out_h)
        cur = _pass(cur, b, k, axis=0)
    return cur


def resize_shorter_edge(h, w, size):
    """torchvision `Resize(int)` output size (functional._compute_resized_output_size, no max_size)."""
    short, long = (w, h) if w <= h else (h, w)
    new_short, new_long = size, int(size * long / short)
    return (new_long, new_short) if w <= h else (new_short, new_long)      # (new_h, new_w)


def minmax_size(h, w, min_size, max_size):
    """`MinMaxResizeForTest.get_size` (reference inference.py:34-55); returns (oh, ow)."""
    size = min_size
    mn, mx = float(min(w, h)), float(max(w, h))
    if mx / mn * size > max_size:
        size = int(round(max_size * mn / mx))
    if (w <= h and w == size) or (h <= w and h == size):
        return h, w
    if w < h:
        ow = size
        oh = int(size * h / w)
    else:
        oh = size
        ow = int(size * w / h)
    return oh, ow


def center_crop_box(h, w, size):
    """torchvision `CenterCrop(size)` on an image at least `size` on both axes -> (top, left)."""
    return int(round((h - size) / 2.0)), int(round((w - size) / 2.0))


def to_tensor_normalize(img, mean=CLIP_MEAN, std=CLIP_STD):
    """ToTensor + Normalize: uint8 HWC -> fp32 CHW, ((x / 255) - mean) / std with every step rounded to fp32."""
    x = img.astype(np.float32) / np.float32(255.0)
    m = np.asarray(mean, dtype=np.float32)[None, None, :]
    s = np.asarray(std, dtype=np.float32)[None, None, :]
    x = (x - m) / s
    return np.ascontiguousarray(x.transpose(2, 0, 1))


def transform(img, param=None):
    """`get_image_transform(param)(pil_image)` for an HxWx3 uint8 RGB array -> fp32 [3, oh, ow]."""
    param = param or {}
    crop = param.get('test_crop_size', 224)
    h, w = img.shape[:2]
    if 'test_respect_ratio_max' in param:
        oh, ow = minmax_size(h, w, crop, param['test_respect_ratio_max'])
        out = img if (oh, ow) == (h, w) else pil_resize_bicubic(img, oh, ow)
    else:
        rh, rw = resize_shorter_edge(h, w, crop)
        r = pil_resize_bicubic(img, rh, rw)
        top, left = center_crop_box(rh, rw, crop)
        out = r[top:top + crop, left:left + crop]
    return to_tensor_normalize(out)
