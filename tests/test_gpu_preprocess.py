"""GPU: the image transform kernels (gitb200_preproc_run through generativeimage2text_b200.inference) against the CPU
oracle, BIT-EXACT (integer / byte work + IEEE fp32 division), and the batched TSV inference path end to end."""
import base64
import io
import json

import numpy as np
import pytest
import torch

import preprocess_oracle as po
from generativeimage2text_b200 import inference as inf
from generativeimage2text_b200 import tsv_io

pytestmark = pytest.mark.gpu


def _img(h, w, seed):
    g = np.random.Generator(np.random.PCG64(seed))
    if seed % 3 == 0:
        return g.integers(0, 256, size=(h, w, 3), dtype=np.uint8)
    yy, xx = np.mgrid[0:h, 0:w]
    base = np.stack([(yy * 3 + xx) % 256, (xx * 2 + seed) % 256, (yy + 2 * xx) % 256], axis=-1).astype(np.int64)
    noise = g.integers(-20, 21, size=(h, w, 3))
    return np.clip(base + noise, 0, 255).astype(np.uint8)


SHAPES = [(480, 640), (640, 480), (224, 224), (225, 1000), (37, 41), (1000, 225), (300, 224), (612, 408), (97, 301),
          (420, 420), (1, 9), (1080, 1920), (223, 225)]


@pytest.mark.parametrize('param', [{}, {'test_crop_size': 160}])
def test_fixed_crop_transform_bit_exact(param):
    t = inf.get_image_transform(param)
    imgs = [_img(h, w, i) for i, (h, w) in enumerate(SHAPES)]
    out = t.batch(imgs)                       # one call, every image a different size
    torch.cuda.synchronize()
    crop = param.get('test_crop_size', 224)
    assert out.shape == (len(imgs), 3, crop, crop) and out.dtype == torch.float32 and out.is_cuda
    got = out.cpu().numpy()
    for i, im in enumerate(imgs):
        want = po.transform(im, param)
        assert np.array_equal(got[i], want), 'image %d %s differs from the PIL/torchvision result' % (i, im.shape)
    # single-image call, PIL input, tensor input
    from PIL import Image
    one = t(Image.fromarray(imgs[0]))
    assert np.array_equal(one.cpu().numpy(), got[0])
    assert np.array_equal(t(torch.from_numpy(imgs[1])).cpu().numpy(), got[1])
    assert t.launch_count() >= 2


@pytest.mark.parametrize('param', [{'test_crop_size': 480, 'test_respect_ratio_max': 640},
                                   {'test_crop_size': 420, 'test_respect_ratio_max': 560}])
def test_minmax_transform_bit_exact(param):
    t = inf.get_image_transform(param)
    imgs = [_img(h, w, 10 + i) for i, (h, w) in enumerate(SHAPES[:10])] + [_img(480, 600, 5), _img(param['test_crop_size'], 500, 6)]
    outs = t.batch(imgs)
    torch.cuda.synchronize()
    assert isinstance(outs, list) and len(outs) == len(imgs)
    for im, o in zip(imgs, outs):
        want = po.transform(im, param)
        assert tuple(o.shape) == (1,) + want.shape
        assert np.array_equal(o[0].cpu().numpy(), want), im.shape


def test_repeated_calls_reuse_buffers_and_device_source():
    """Back-to-back calls on one handle (staging reuse) stay exact; a bad descriptor is refused, not executed."""
    import ctypes
    from generativeimage2text_b200 import _lib
    t = inf.get_image_transform({})
    for rep in range(4):
        imgs = [_img(200 + 17 * rep + 3 * i, 260 + 11 * i, 40 + rep * 8 + i) for i in range(6)]
        out = t.batch(imgs).cpu().numpy()
        for i, im in enumerate(imgs):
            assert np.array_equal(out[i], po.transform(im, {}))
    lib = _lib.load()
    d = (_lib.ImageDesc * 1)(_lib.ImageDesc(0, 10, 10, 224, 224, 0, 0, 300, 224, 0))   # crop window outside the image
    src = torch.zeros(300, dtype=torch.uint8, device='cuda')
    dst = torch.zeros(3 * 300 * 224, dtype=torch.float32, device='cuda')
    mean = (ctypes.c_float * 3)(*inf.CLIP_MEAN)
    std = (ctypes.c_float * 3)(*inf.CLIP_STD)
    rc = lib.gitb200_preproc_run(t._handle, src.data_ptr(), 300, 0, d, 1, mean, std, dst.data_ptr(), dst.numel(), None)
    assert rc != 0 and b'crop window' in lib.gitb200_preproc_last_error(t._handle)
    # device-resident source
    im = _img(120, 90, 2)
    rh, rw, top, left, oh, ow = t.geometry(120, 90)
    d = (_lib.ImageDesc * 1)(_lib.ImageDesc(0, 120, 90, rh, rw, top, left, oh, ow, 0))
    src = torch.from_numpy(im.reshape(-1)).cuda()
    dst = torch.empty(3 * oh * ow, dtype=torch.float32, device='cuda')
    rc = lib.gitb200_preproc_run(t._handle, src.data_ptr(), src.numel(), 0, d, 1, mean, std, dst.data_ptr(), dst.numel(),
                                 torch.cuda.current_stream().cuda_stream)
    assert rc == 0, lib.gitb200_preproc_last_error(t._handle)
    torch.cuda.synchronize()
    assert np.array_equal(dst.view(3, oh, ow).cpu().numpy(), po.transform(im, {}))


class StubTokenizer:
    """Stands in for BertTokenizer (its vocabulary file is not available offline): ids <-> decimal strings."""
    cls_token_id, sep_token_id = 101, 102

    def decode(self, ids, skip_special_tokens=True):
        return ' '.join(str(i) for i in ids if not (skip_special_tokens and i in (0, 101, 102)))

    def __call__(self, text, **kw):
        return {'input_ids': [int(x) for x in text.split()]}


def _png_b64(arr):
    from PIL import Image
    buf = io.BytesIO()
    Image.fromarray(arr).save(buf, format='PNG')       # lossless: every decoder returns the same bytes
    return base64.b64encode(buf.getvalue())


def test_tsv_inference_batched_end_to_end(tmp_path):
    from generativeimage2text_b200.model import get_git_model, AutoRegressiveBeamSearch
    from generativeimage2text_b200.synthetic import synthetic_state_dict
    tok = StubTokenizer()
    imgs = [_img(150 + 13 * i, 200 + 7 * i, 70 + i) for i in range(11)]
    in_tsv = str(tmp_path / 'images.tsv')
    tsv_io.tsv_writer(((('key%02d' % i), _png_b64(im)) for i, im in enumerate(imgs)), in_tsv)
    sd = synthetic_state_dict({}, 0, 'perturbed')
    model = get_git_model(tok, {})
    model.load_state_dict(sd, strict=True)
    model.decoder = AutoRegressiveBeamSearch(102, max_steps=8, beam_size=1, per_node_beam_size=1, fix_missing_prefix=True)
    out_tsv = str(tmp_path / 'pred.tsv')
    n = inf.test_git_inference_single_tsv(in_tsv, 'GIT_BASE', None, out_tsv, tokenizer=tok, param={}, batch_size=4, depth=2,
                                          model=model)
    assert n == len(imgs)
    rows = list(tsv_io.TSVFile(out_tsv))
    assert [r[0] for r in rows] == ['key%02d' % i for i in range(len(imgs))]
    # the same pixels through the oracle transform and a direct model call: same captions up to run-to-run noise
    x = torch.from_numpy(np.stack([po.transform(im, {}) for im in imgs])).cuda()
    direct = model({'image': x})['predictions'].tolist()
    agree = total = 0
    for r, d in zip(rows, direct):
        cap = json.loads(r[1])
        assert isinstance(cap, list) and set(cap[0]) == {'caption'}
        got = cap[0]['caption'].split()
        want = tok.decode(d).split()
        assert len(got) == len(want) == 6          # 8 columns minus the two [CLS]
        agree += sum(a == b for a, b in zip(got, want))
        total += len(want)
    assert agree / total >= 0.75
    # question path (one prefix per call, batch 1 like the reference) + VQA json conversion
    q_tsv = str(tmp_path / 'questions.tsv')
    tsv_io.tsv_writer(((('key%02d' % i), json.dumps([{'question': '2054 2003', 'question_id': 100 + i}])) for i in range(3)),
                      q_tsv)
    short_tsv = str(tmp_path / 'images3.tsv')
    tsv_io.tsv_writer(((('key%02d' % i), _png_b64(im)) for i, im in enumerate(imgs[:3])), short_tsv)
    ans_tsv = str(tmp_path / 'answers.tsv')
    assert inf.test_git_inference_single_tsv(short_tsv, 'GIT_BASE', q_tsv, ans_tsv, tokenizer=tok, param={}, model=model) == 3
    inf.convert_tsv_to_vqa_json(ans_tsv, str(tmp_path / 'answers.json'))
    answers = json.load(open(str(tmp_path / 'answers.json')))
    assert [a['question_id'] for a in answers] == [100, 101, 102] and all('answer' in a for a in answers)


def test_variable_resolution_inputs_through_the_model():
    """A respect-ratio model end to end: transform -> [1,3,oh,ow] of a non-square size -> run-time positional-embedding
    re-sampling in the engine; features against the oracle."""
    import git_oracle
    from generativeimage2text_b200.model import get_git_model
    from generativeimage2text_b200.synthetic import synthetic_state_dict
    param = {'test_crop_size': 160, 'test_respect_ratio_max': 224}
    sd = synthetic_state_dict(param, 0, 'perturbed')
    m = get_git_model(StubTokenizer(), param)
    m.load_state_dict(sd, strict=True)
    m = m.cuda().eval()
    t = inf.get_image_transform(param)
    for hw in [(300, 400), (200, 200), (500, 300)]:
        x = t(_img(hw[0], hw[1], 9)).unsqueeze(0)
        feats = m.encode_image(x)
        torch.cuda.synchronize()
        ref = git_oracle.encode_image(sd, param, x.cpu())
        err = (feats.cpu() - ref).abs()
        assert feats.shape == ref.shape
        assert err.mean().item() < 0.01 and err.max().item() < 0.15, (hw, err.max().item())
