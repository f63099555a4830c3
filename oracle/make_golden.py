"""TEST INFRASTRUCTURE ONLY -- generates tests/golden/*.npz by running the UNMODIFIED reference
(/root/reference, imported through oracle/ref_shim.py) on seeded synthetic checkpoints.

Run in the build container only:  python oracle/make_golden.py [case ...]

The reference ships no golden vectors of its own; these fixtures pin
oracle/git_oracle.py (tests/test_oracle_golden.py) and, through it and directly, the CUDA engine.
Per case we store: the config, `predictions`, `logprobs`, a strided sample of the image features
`CaptioningModel.forward_one` hands to the decoder, and for every `decoding_step` call the raw
last-position logits at 256 fixed vocabulary columns plus the top-2 values / indices per row.
The reference's source is not modified: `decoding_step` and `image_encoder.forward` are observed by
wrapping the bound methods on the instance.
"""
import json
import os
import sys
import time

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.dirname(HERE))

import ref_shim  # noqa: E402
from generativeimage2text_b200.synthetic import synthetic_state_dict, synthetic_images  # noqa: E402

GOLDEN_DIR = os.path.join(os.path.dirname(HERE), 'tests', 'golden')
LARGE = {'visual_feature_size': 1024, 'image_encoder_type': 'CLIPViT_L_14'}

CASES = {
    # name: param, variant, batch, frames, search, max_steps, prefix
    'base_greedy_init': dict(param={}, variant='init', batch=2, frames=0, search='greedy', max_steps=40),
    'base_greedy': dict(param={}, variant='perturbed', batch=2, frames=0, search='greedy', max_steps=40),
    'base_beam': dict(param={}, variant='init', batch=2, frames=0, search='beam', max_steps=40),
    'base_prefix': dict(param={}, variant='perturbed', batch=1, frames=1, search='greedy', max_steps=20,
                        prefix=[101, 2054, 2003, 2023]),
    'vatex_greedy': dict(param={'num_image_with_embedding': 6}, variant='perturbed', batch=1, frames=6,
                         search='greedy', max_steps=16),
    'large_greedy': dict(param=LARGE, variant='perturbed', batch=1, frames=0, search='greedy', max_steps=12),
    'large_beam': dict(param=LARGE, variant='init', batch=1, frames=0, search='beam', max_steps=12),
    # MinMaxResizeForTest-style input: a 160-crop model (10x10 grid embedding) fed 160x208 pixels (10x13 grid) ->
    # run-time positional-embedding interpolation (reference layers/CLIP/model.py:245-251)
    'base_ratio_greedy': dict(param={'test_crop_size': 160, 'test_respect_ratio_max': 224}, variant='perturbed', batch=2,
                              frames=0, search='greedy', max_steps=12, image_hw=[160, 208]),
    # the shipped GIT_BASE_VQAv2 / TEXTVQA geometry (aux_data/models/GIT_BASE_VQAv2/parameter.yaml): 480-crop model
    # (30x30 grid embedding), a 480x640 input (30x40 grid = 1201 image tokens) and a question prefix
    'base_vqa_ratio_greedy': dict(param={'test_crop_size': 480, 'test_respect_ratio_max': 640}, variant='perturbed', batch=1,
                                  frames=1, search='greedy', max_steps=10, image_hw=[480, 640], prefix=[101, 2054, 2003, 2023]),
    # square non-default crop: the embedding is built for the 10x10 grid, no run-time interpolation
    'base_crop160_greedy': dict(param={'test_crop_size': 160}, variant='perturbed', batch=2, frames=1, search='greedy',
                                max_steps=12, image_hw=[160, 160]),
}


def vocab_sample():
    g = np.random.Generator(np.random.PCG64(777))
    return np.sort(g.choice(30522, size=256, replace=False)).astype(np.int64)


def run_case(name, cfg, seed=0, img_seed=1234):
    sd = synthetic_state_dict(cfg['param'], seed, cfg['variant'])
    model = ref_shim.load_reference_model(cfg['param'], cfg['search'], cfg['max_steps'], state_dict=sd)
    image = synthetic_images(cfg['batch'], cfg['frames'], img_seed, cfg.get('image_hw', 224))
    batch = {'image': image}
    if 'prefix' in cfg:
        batch['prefix'] = torch.tensor([cfg['prefix']], dtype=torch.long)
    cols = torch.from_numpy(vocab_sample())
    steps = []
    orig = model.decoding_step

    def spy(*a, **kw):
        z = orig(*a, **kw)
        top = z.topk(2, dim=1)
        steps.append((z[:, cols].clone(), top.values.clone(), top.indices.clone()))
        return z
    model.decoding_step = spy
    t0 = time.time()
    with torch.no_grad():
        out = model(batch)
        # image features as the decoder sees them (reference layers/decoder.py:846-857)
        if isinstance(image, (list, tuple)):
            fs = [model.image_encoder(im) for im in image]
            if model.num_image_with_embedding:
                fs = [f + e for f, e in zip(fs, model.img_temperal_embedding)]
            vf = torch.cat(fs, dim=1)
        else:
            vf = model.image_encoder(image)
        vproj = model.textual.visual_projection(vf)
    dt = time.time() - t0
    meta = dict(cfg)
    meta.update(seed=seed, img_seed=img_seed, reference_commit='faae4fb9', torch=torch.__version__,
                generator='oracle/make_golden.py', seconds=round(dt, 2))
    np.savez_compressed(
        os.path.join(GOLDEN_DIR, name + '.npz'),
        meta=np.array(json.dumps(meta)),
        predictions=out['predictions'].numpy(),
        logprobs=out['logprobs'].numpy(),
        vocab_cols=cols.numpy(),
        step_logits=torch.stack([s[0] for s in steps]).numpy(),
        step_top2_val=torch.stack([s[1] for s in steps]).numpy(),
        step_top2_idx=torch.stack([s[2] for s in steps]).numpy(),
        feats_sample=vf[:, ::17, ::29].numpy(),
        vproj_sample=vproj[:, ::17, ::29].numpy(),
        feats_absmean=np.array(vf.abs().mean().item(), dtype=np.float64),
    )
    print('%-18s %5.1fs steps=%d pred=%s lp=%s' % (
        name, dt, len(steps), tuple(out['predictions'].shape),
        np.round(out['logprobs'].flatten().numpy(), 4).tolist()))


if __name__ == '__main__':
    os.makedirs(GOLDEN_DIR, exist_ok=True)
    torch.set_num_threads(os.cpu_count())
    names = sys.argv[1:] or list(CASES)
    for n in names:
        run_case(n, CASES[n])
